/*
 * qo_ts2c.cu -- instantiations of the thread-per-sample transfer-function kernel (qo_ts.cuh) for plain ladders, 9-10 numerator pairs.
 * The kernels are split over several translation units (one kernel per numerator length, each carrying one loop body per
 * denominator length) so that they compile in parallel.
 */
#include <cuda_runtime.h>
#include "qo_ts.cuh"
#include "qo_ts_launch.h"

extern "C" ts_fn qo_ts_kernel_2c(int kn)
{
    switch (kn) {
    case 9: return qo_mc_ts_kernel<2, false, QO_TS_PT2, QO_TS_MINB2, 9>;
    case 10: return qo_mc_ts_kernel<2, false, QO_TS_PT2, QO_TS_MINB2, 10>;
    default: return nullptr;
    }
}
