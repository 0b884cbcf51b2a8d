/*
 * qo_tf_front.cu -- instantiations of qo_mc_tf_kernel for a transmission line or a measured two-port in front of the lumped
 * cascade (CPLM = 4; the reference puts such blocks at the source end: util/pa-bias-simulation/pa-bias-simulation.sch:39) and
 * for |S11| specs behind any front block (two row vectors of the block).  A translation unit of its own so that the library
 * builds in parallel.
 */
#include <cuda_runtime.h>
#include "qo_tf.cuh"
#include "qo_tf_launch.h"

#define QO_TF_TPB 128
#define QO_TF_MINB 4
#define QO_TF_CPL_PP 2
#define QO_TF_CPL_TPB 128
#define QO_TF_CPL_MINB 4
typedef void (*tf_fn)(const TfParams);

template <int CPL, bool S11, int NS, int PP, int TPB, int MINB> static tf_fn pick_ns(int den)
{
    switch (den) {
    case QO_TF_DEN_NONE: return qo_mc_tf_kernel<4, QO_TF_DEN_NONE, CPL, S11, false, NS, PP, TPB, MINB>;
    case QO_TF_DEN_E: return qo_mc_tf_kernel<4, QO_TF_DEN_E, CPL, S11, false, NS, PP, TPB, MINB>;
    case QO_TF_DEN_D: return qo_mc_tf_kernel<4, QO_TF_DEN_D, CPL, S11, false, NS, PP, TPB, MINB>;
    default: return nullptr;
    }
}
template <int CPL, bool S11, int PP, int TPB, int MINB> static tf_fn pick(int den, int nspec)
{
    return nspec > 4 ? pick_ns<CPL, S11, 8, PP, TPB, MINB>(den) : pick_ns<CPL, S11, 4, PP, TPB, MINB>(den);
}

tf_fn qo_tf_pick_front(int front, int s11, int pp, int den, int nspec, int *tpb, int *minb)
{
    if (pp == 1) {
        *tpb = QO_TF_TPB; *minb = QO_TF_MINB;
        if (front) return s11 ? pick<4, true, 1, QO_TF_TPB, QO_TF_MINB>(den, nspec) : pick<4, false, 1, QO_TF_TPB, QO_TF_MINB>(den, nspec);
        return s11 ? pick<1, true, 1, QO_TF_TPB, QO_TF_MINB>(den, nspec) : nullptr;
    }
    if (pp == QO_TF_CPL_PP) {
        *tpb = QO_TF_CPL_TPB; *minb = QO_TF_CPL_MINB;
        if (front) return s11 ? pick<4, true, QO_TF_CPL_PP, QO_TF_CPL_TPB, QO_TF_CPL_MINB>(den, nspec) : pick<4, false, QO_TF_CPL_PP, QO_TF_CPL_TPB, QO_TF_CPL_MINB>(den, nspec);
        return s11 ? pick<1, true, QO_TF_CPL_PP, QO_TF_CPL_TPB, QO_TF_CPL_MINB>(den, nspec) : nullptr;
    }
    return nullptr;
}
