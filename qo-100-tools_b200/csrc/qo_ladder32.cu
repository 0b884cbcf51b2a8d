/*
 * qo_ladder32.cu -- FP32 instantiations of the straight-line ladder kernel (qo_ladder.cuh), the optional
 * "1e-3 dB" mode of north_star.  A separate translation unit so that it builds in parallel with the FP64 ones.
 */
#include <cuda_runtime.h>
#include "qo_ladder.cuh"

/* four points per thread, three blocks per SM (77-80 registers) */
#define QO_LAD32_PP 2
#define QO_LAD32_TPB 256
#define QO_LAD32_MINB 3

typedef void (*lad_fn)(const LadParams);

template <typename T, int N, int FIRST, bool CPL, int NROWS, int PP, int TPB, int MINB> static lad_fn lad_get()
{
    return qo_mc_ladder_kernel<T, N, FIRST, CPL, NROWS, PP, TPB, MINB>;
}

template <typename T, int NROWS, int PP, int TPB, int MINB> static lad_fn lad_pick(int n, int first, int cpl)
{
#define QO_LAD_ROW(NN)                                                                       \
    case NN:                                                                                 \
        return cpl ? (first ? lad_get<T, NN, 1, true, NROWS, PP, TPB, MINB>() : lad_get<T, NN, 0, true, NROWS, PP, TPB, MINB>())   \
                   : (first ? lad_get<T, NN, 1, false, NROWS, PP, TPB, MINB>() : lad_get<T, NN, 0, false, NROWS, PP, TPB, MINB>());
    switch (n) {
        QO_LAD_ROW(1) QO_LAD_ROW(2) QO_LAD_ROW(3) QO_LAD_ROW(4) QO_LAD_ROW(5) QO_LAD_ROW(6)
        QO_LAD_ROW(7) QO_LAD_ROW(8) QO_LAD_ROW(9) QO_LAD_ROW(10) QO_LAD_ROW(11)
    default: return nullptr;
    }
#undef QO_LAD_ROW
}


extern "C" lad_fn qo_ladder_pick32(int n, int first, int cpl, int *tpb, int *minb)
{
    *tpb = QO_LAD32_TPB; *minb = QO_LAD32_MINB;
    return lad_pick<float, 1, QO_LAD32_PP, QO_LAD32_TPB, QO_LAD32_MINB>(n, first, cpl);
}
