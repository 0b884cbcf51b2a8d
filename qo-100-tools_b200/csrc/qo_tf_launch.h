/* qo_tf_launch.h -- host-visible entry points of the transfer-function kernel (qo_tf.cu) */
#pragma once
#include <cuda_runtime.h>
struct TfParams;
struct DevProg;
/* Can this job run on the transfer-function kernel?  Structural test + numerical self-check of the polynomial
 * expansion against the per-element evaluation at the nominal network and its tolerance-box corners, on every
 * grid point.  Returns 1 and fills K (coefficients per chain), mode (QO_TF_*), wref, el0/n_el (lumped ops), cpl_op,
 * err (worst relative disagreement on |den|^2 seen); 0 when the job must stay on the chain kernels (why: reason). */
extern "C" int qo_tf_plan_check(const DevProg *hp, int mode_reduce_only, int precision, int generic, const double *f, int nf,
                                const unsigned char *mask, int *K, int *mode, double *wref, int *el0, int *n_el, int *cpl_op,
                                double *err, const char **reason);
/* pp = frequency pairs per thread per iteration the plan padded its tables for (4: |S21| modes, 2: coupler mode);
 * returns 0, -1 when no instantiation covers (K, mode, pp), else the cudaError_t of the launch */
extern "C" int qo_tf_launch(int K, int mode, int pp, int variant, int sm_count, const TfParams *P, cudaStream_t st);
