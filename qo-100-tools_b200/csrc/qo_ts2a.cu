/*
 * qo_ts2a.cu -- instantiations of the thread-per-sample transfer-function kernel (qo_ts.cuh) for plain ladders, 2-5 numerator pairs.
 * The kernels are split over several translation units (one kernel per numerator length, each carrying one loop body per
 * denominator length) so that they compile in parallel.
 */
#include <cuda_runtime.h>
#include "qo_ts.cuh"
#include "qo_ts_launch.h"

extern "C" ts_fn qo_ts_kernel_2a(int kn)
{
    switch (kn) {
    case 2: return qo_mc_ts_kernel<2, false, QO_TS_PT2, QO_TS_MINB2, 2>;
    case 3: return qo_mc_ts_kernel<2, false, QO_TS_PT2, QO_TS_MINB2, 3>;
    case 4: return qo_mc_ts_kernel<2, false, QO_TS_PT2, QO_TS_MINB2, 4>;
    case 5: return qo_mc_ts_kernel<2, false, QO_TS_PT2, QO_TS_MINB2, 5>;
    default: return nullptr;
    }
}
