/*
 * qo_ts4.cu -- instantiations of the thread-per-sample transfer-function kernel (qo_ts.cuh) for ladders behind a coupled-line block.
 * The kernels are split over several translation units (one kernel per numerator length, each carrying one loop body per
 * denominator length) so that they compile in parallel.
 */
#include <cuda_runtime.h>
#include "qo_ts.cuh"
#include "qo_ts_launch.h"

extern "C" ts_fn qo_ts_kernel_4(int kn)
{
    switch (kn) {
    case 2: return qo_mc_ts_kernel<4, true, QO_TS_PT4, QO_TS_MINB4, 2>;
    case 3: return qo_mc_ts_kernel<4, true, QO_TS_PT4, QO_TS_MINB4, 3>;
    case 4: return qo_mc_ts_kernel<4, true, QO_TS_PT4, QO_TS_MINB4, 4>;
    case 5: return qo_mc_ts_kernel<4, true, QO_TS_PT4, QO_TS_MINB4, 5>;
    case 6: return qo_mc_ts_kernel<4, true, QO_TS_PT4, QO_TS_MINB4, 6>;
    case 7: return qo_mc_ts_kernel<4, true, QO_TS_PT4, QO_TS_MINB4, 7>;
    case 8: return qo_mc_ts_kernel<4, true, QO_TS_PT4, QO_TS_MINB4, 8>;
    default: return nullptr;
    }
}
