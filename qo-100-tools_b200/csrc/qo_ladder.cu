/*
 * qo_ladder.cu -- instantiations and launcher of the straight-line ladder kernel
 * (qo_ladder.cuh).  Its own translation unit so that the 44 instantiations
 * (N = 1..11, series/shunt first, with/without the coupled-line block) compile
 * in parallel with qo_cuda.cu.
 */
#include <cuda_runtime.h>
#include "qo_ladder.cuh"
#include "qo_ladder_launch.h"

/* launch shape chosen from the r01 ncu sweeps (DESIGN.md "ladder kernel"): */
#ifndef QO_LAD_PP
#define QO_LAD_PP 2
#define QO_LAD_TPB 256
#define QO_LAD_MINB 2
#endif

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

#ifdef QO_LAD_EXPERIMENT
/* development builds: extra launch shapes of the 11-element series-first ladder, chosen with QO100NET_LAD_VARIANT */
template <int PP, int TPB, int MINB> static lad_fn lad_pick11(int n, int first, int cpl)
{
    if (n != 11 || first) return nullptr;
    return cpl ? lad_get<double, 11, 0, true, 1, PP, TPB, MINB>() : lad_get<double, 11, 0, false, 1, PP, TPB, MINB>();
}
#endif

/* kernels with the |S11| row: two points per thread (the second row vector doubles the chain state) */
#define QO_LAD2_PP 1
#define QO_LAD2_TPB 256
#define QO_LAD2_MINB 2

typedef void (*lad_fn_t)(const LadParams);
extern "C" lad_fn_t qo_ladder_pick32(int n, int first, int cpl, int *tpb, int *minb);

extern "C" int qo_ladder_launch(int n, int first, int cpl, int nrows, int precision, int variant, int sm_count, const LadParams *P,
                                cudaStream_t st, const char **shape)
{
    lad_fn fn = nullptr;
    int tpb = QO_LAD_TPB, minb = QO_LAD_MINB;
    const char *name = "default";
    (void)variant;
#ifdef QO_LAD_EXPERIMENT
    switch (variant) {
    case 1: fn = lad_pick11<1, 256, 2>(n, first, cpl); tpb = 256; minb = 2; name = "pp1-256x2"; break;
    case 2: fn = lad_pick11<1, 128, 5>(n, first, cpl); tpb = 128; minb = 5; name = "pp1-128x5"; break;
    case 3: fn = lad_pick11<2, 256, 2>(n, first, cpl); tpb = 256; minb = 2; name = "pp2-256x2"; break;
    case 4: fn = lad_pick11<1, 128, 6>(n, first, cpl); tpb = 128; minb = 6; name = "pp1-128x6"; break;
    case 5: fn = lad_pick11<1, 256, 4>(n, first, cpl); tpb = 256; minb = 4; name = "pp1-256x4"; break;
    case 6: fn = lad_pick11<2, 128, 4>(n, first, cpl); tpb = 128; minb = 4; name = "pp2-128x4"; break;
    case 7: fn = lad_pick11<2, 128, 5>(n, first, cpl); tpb = 128; minb = 5; name = "pp2-128x5"; break;
    case 8: fn = lad_pick11<2, 256, 3>(n, first, cpl); tpb = 256; minb = 3; name = "pp2-256x3"; break;
    case 9: fn = lad_pick11<2, 64, 8>(n, first, cpl); tpb = 64; minb = 8; name = "pp2-64x8"; break;
    default: break;
    }
#endif
    if (precision == 32) {
        if (nrows != 1) return -1;
        fn = qo_ladder_pick32(n, first, cpl, &tpb, &minb);      /* qo_ladder32.cu */
        name = "fp32";
    } else if (nrows == 2) {
        fn = lad_pick<double, 2, QO_LAD2_PP, QO_LAD2_TPB, QO_LAD2_MINB>(n, first, cpl);
        tpb = QO_LAD2_TPB; minb = QO_LAD2_MINB; name = "s11";
    } else if (!fn) {
        fn = lad_pick<double, 1, QO_LAD_PP, QO_LAD_TPB, QO_LAD_MINB>(n, first, cpl);
        tpb = QO_LAD_TPB; minb = QO_LAD_MINB; name = "default";
    }
    if (!fn) return -1;
    if (shape) *shape = name;
    const unsigned long long warps = (unsigned long long)(tpb / 32);
    unsigned long long blocks = (P->nsamples + warps - 1) / warps;
    const unsigned long long resident = (unsigned long long)sm_count * (unsigned long long)minb;
    if (blocks > resident) blocks = resident;
    if (blocks < 1) blocks = 1;
    fn<<<(unsigned)blocks, tpb, 0, st>>>(*P);
    return (int)cudaGetLastError();
}
