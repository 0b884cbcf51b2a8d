/*
 * qo_ts2b.cu -- instantiations of the thread-per-sample transfer-function kernel (qo_ts.cuh) for plain ladders, 6-8 numerator pairs.
 * The kernels are split over several translation units (one kernel per numerator length, each carrying one loop body per
 * denominator length) so that they compile in parallel.
 */
#include <cuda_runtime.h>
#include "qo_ts.cuh"
#include "qo_ts_launch.h"

extern "C" ts_fn qo_ts_kernel_2b(int kn)
{
    switch (kn) {
    case 6: return qo_mc_ts_kernel<2, false, QO_TS_PT2, QO_TS_MINB2, 6>;
    case 7: return qo_mc_ts_kernel<2, false, QO_TS_PT2, QO_TS_MINB2, 7>;
    case 8: return qo_mc_ts_kernel<2, false, QO_TS_PT2, QO_TS_MINB2, 8>;
    default: return nullptr;
    }
}
