/*
 * qo_ts.cu -- instantiations and launcher of the thread-per-sample transfer-function kernel (qo_ts.cuh).
 * Its own translation unit: every kernel carries one loop body per kept-length pair (KN, KE), which is most of the
 * library's compile time.  -DQO_TS_DEV_KN=9 -DQO_TS_DEV_KE=14 (development) compiles that one body only.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "qo_ts.cuh"
#include "qo_tf_launch.h"

/* ---- thread-per-sample flavour (qo_ts.cuh) ------------------------------------------------------------------- */
/* Can the bulk of this plan's launches run thread-per-sample?  Plain |S21| jobs with at most four specs whose kept polynomial
 * lengths fit the register capacities; behind a coupler only the matched / uniform-grid / equal-angle form (config 5). */
extern "C" int qo_ts_eligible(const TfPlan *tp, int nspec, int cpl_rot_same)
{
    const char *force = getenv("QO100NET_KERNEL");
    if (force && strcmp(force, "tf") == 0) return 0;          /* keep everything on the warp-per-sample kernel (A/B runs) */
    if (tp->s11 || tp->gd || nspec > 4) return 0;
    if (tp->den != QO_TF_DEN_NONE && tp->den != QO_TF_DEN_E) return 0;
    if (tp->kn < 2 || (tp->den == QO_TF_DEN_E && (tp->kd < 2 || (tp->kd & 1)))) return 0;
    if (tp->cpl_op < 0) return tp->nn == 2 && tp->kn <= QO_TS_CAPN2 && (tp->den == QO_TF_DEN_NONE || tp->kd <= QO_TS_CAPE2);
    return cpl_rot_same && tp->nn == 4 && tp->kn <= QO_TS_CAPN4 && (tp->den == QO_TF_DEN_NONE || tp->kd <= QO_TS_CAPE4);
}

typedef void (*ts_fn)(const TsParams);
static ts_fn ts_pick_kernel(const TfPlan *tp)
{
    if (tp->cpl_op < 0) return tp->den == QO_TF_DEN_E ? qo_mc_ts_kernel<2, QO_TF_DEN_E, false> : qo_mc_ts_kernel<2, QO_TF_DEN_NONE, false>;
    return tp->den == QO_TF_DEN_E ? qo_mc_ts_kernel<4, QO_TF_DEN_E, true> : qo_mc_ts_kernel<4, QO_TF_DEN_NONE, true>;
}

/* blocks of this plan's kernel that are resident on one SM at the same time (what the hardware grants, not what
 * __launch_bounds__ asked for): the grid is exactly one wave of them, so that the static deal of samples is balanced */
static int ts_blocks_per_sm(const TfPlan *tp)
{
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ts_pick_kernel(tp), QO_TS_TPB, 0) != cudaSuccess) { cudaGetLastError(); occ = 0; }
    if (getenv("QO100NET_TS_DEBUG")) fprintf(stderr, "qo_ts: %d blocks of %d threads resident per SM\n", occ, QO_TS_TPB);
    return occ;
}

/* threads of one full wave of the thread-per-sample kernel: the launcher deals whole rounds of one sample per thread */
extern "C" int qo_ts_wave_threads(const TfPlan *tp, int sm_count) { return sm_count * ts_blocks_per_sm(tp) * QO_TS_TPB; }

extern "C" int qo_ts_launch(const TfPlan *tp, int sm_count, const TsParams *Q, cudaStream_t st)
{
    const int bps = ts_blocks_per_sm(tp);
    if (bps <= 0) return (int)cudaErrorLaunchOutOfResources;
    ts_pick_kernel(tp)<<<(unsigned)(sm_count * bps), QO_TS_TPB, 0, st>>>(*Q);
    return (int)cudaGetLastError();
}

