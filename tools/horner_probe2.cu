// horner_probe2.cu -- development microbenchmark: the frequency loop of qo_tf.cuh with the per-sample coefficients held in
// REGISTERS (compile-time lengths KN pairs / KE coefficients) instead of being re-read from shared memory at every Horner step.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o qo-100-tools_b200/lib/horner_probe2 tools/horner_probe2.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int KN, int KE, int PTS, int TPB, int MINB, bool INREG>
__global__ void __launch_bounds__(TPB, MINB) probe(double *out, int nsamp, int npts_per_lane, const double2 *yt, double thr)
{
    __shared__ __align__(16) double s_num[TPB / 32][2 * KN], s_e[TPB / 32][KE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int fails = 0;
    for (int s = 0; s < nsamp; s++) {
        // "per-sample stage": new coefficients (cheap stand-in), written by the lanes, read back by everyone
        if (lane < 2 * KN) s_num[warp][lane] = 1.0 / (1.0 + lane + 1e-3 * s);
        if (lane < KE) s_e[warp][lane] = 1.0 / (2.0 + lane + 1e-3 * s);
        __syncwarp();
        double cn[2 * KN], ce[KE];
        if (INREG) {
#pragma unroll
            for (int k = 0; k < 2 * KN; k += 2) { const double2 v = *(const double2 *)&s_num[warp][k]; cn[k] = v.x; cn[k + 1] = v.y; }
#pragma unroll
            for (int k = 0; k < KE; k += 2) { const double2 v = *(const double2 *)&s_e[warp][k]; ce[k] = v.x; ce[k + 1] = v.y; }
        }
        const unsigned nb = (unsigned)__cvta_generic_to_shared(&s_num[warp][0]), eb = (unsigned)__cvta_generic_to_shared(&s_e[warp][0]);
        unsigned int acc = 0;
        for (int it = 0; it < npts_per_lane / PTS; it++) {
            double y[PTS];
#pragma unroll
            for (int q = 0; q < PTS / 2; q++) { const double2 a = yt[(it * (PTS / 2) + q) * 32 + lane]; y[2 * q] = a.x; y[2 * q + 1] = a.y; }
            double re[PTS], ro[PTS], dd[PTS];
            if (INREG) {
#pragma unroll
                for (int p = 0; p < PTS; p++) { re[p] = fma(cn[2 * KN - 2], y[p], cn[2 * KN - 4]); ro[p] = fma(cn[2 * KN - 1], y[p], cn[2 * KN - 3]); }
#pragma unroll
                for (int k = KN - 3; k >= 0; k--)
#pragma unroll
                    for (int p = 0; p < PTS; p++) { re[p] = fma(re[p], y[p], cn[2 * k]); ro[p] = fma(ro[p], y[p], cn[2 * k + 1]); }
#pragma unroll
                for (int p = 0; p < PTS; p++) dd[p] = fma(ce[KE - 1], y[p], ce[KE - 2]);
#pragma unroll
                for (int k = KE - 3; k >= 0; k--)
#pragma unroll
                    for (int p = 0; p < PTS; p++) dd[p] = fma(dd[p], y[p], ce[k]);
            } else {
                double2 t, u;
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(t.x), "=d"(t.y) : "r"(nb + 16u * (KN - 1)));
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(u.x), "=d"(u.y) : "r"(nb + 16u * (KN - 2)));
#pragma unroll
                for (int p = 0; p < PTS; p++) { re[p] = fma(t.x, y[p], u.x); ro[p] = fma(t.y, y[p], u.y); }
#pragma unroll
                for (int k = KN - 3; k >= 0; k--) {
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(u.x), "=d"(u.y) : "r"(nb + 16u * k));
#pragma unroll
                    for (int p = 0; p < PTS; p++) { re[p] = fma(re[p], y[p], u.x); ro[p] = fma(ro[p], y[p], u.y); }
                }
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(t.x), "=d"(t.y) : "r"(eb + 8u * (KE - 2)));
#pragma unroll
                for (int p = 0; p < PTS; p++) dd[p] = fma(t.y, y[p], t.x);
#pragma unroll
                for (int k = KE - 4; k >= 0; k -= 2) {
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(u.x), "=d"(u.y) : "r"(eb + 8u * k));
#pragma unroll
                    for (int p = 0; p < PTS; p++) { dd[p] = fma(dd[p], y[p], u.y); dd[p] = fma(dd[p], y[p], u.x); }
                }
            }
#pragma unroll
            for (int p = 0; p < PTS; p++) {
                const double t = ro[p] * ro[p], n2 = fma(-y[p], t, re[p] * re[p]);
                acc |= (unsigned int)__double2hiint(fma(thr, dd[p], -n2));
            }
        }
        fails += __reduce_or_sync(0xffffffffu, acc) >> 31;
        __syncwarp();
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = fails;
}

template <int KN, int KE, int PTS, int TPB, int MINB, bool INREG> static void run(const char *name, int nsm, const double2 *y)
{
    const int blocks = nsm * MINB, nsamp = 64, npl = 128;
    double *d;
    cudaMalloc(&d, (size_t)blocks * TPB * sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        probe<KN, KE, PTS, TPB, MINB, INREG><<<blocks, TPB>>>(d, nsamp, npl, y, 0.37);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double per_pt = 2.0 * (KN - 1) + (KE - 1) + 4.0;      // Num chains, E, n2 (2 DMUL + DFMA), sign DFMA
    const double fp64 = (double)blocks * TPB * nsamp * npl * per_pt;
    const double rate = fp64 / (best * 1e-3);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, probe<KN, KE, PTS, TPB, MINB, INREG>);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe<KN, KE, PTS, TPB, MINB, INREG>, TPB, 0);
    printf("%-34s KN %2d KE %2d PTS %d  %3d regs, %d x %d thr/SM (occupancy limit %d blocks): %7.3f ms  %5.1f %% FP64 pipe  %.3e evals/s\n", name, KN, KE, PTS,
           fa.numRegs, MINB, TPB, occ, best, rate / (148.0 * 64 * 1.965e9) * 100, (double)blocks * TPB * nsamp * npl / (best * 1e-3));
    cudaFree(d);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int nsm = p.multiProcessorCount;
    double2 *y; cudaMalloc(&y, 32 * 64 * sizeof(double2));
    static double2 hy[32 * 64]; for (int i = 0; i < 32 * 64; i++) { hy[i].x = -0.9 + 1e-4 * i; hy[i].y = -0.8 + 1e-4 * i; }
    cudaMemcpy(y, hy, sizeof hy, cudaMemcpyHostToDevice);
    printf("%s, %d SMs\n", p.name, nsm);
    run<9, 14, 8, 128, 4, false>("shared-memory coefficients", nsm, y);
    run<9, 14, 4, 128, 4, false>("shared-memory coefficients", nsm, y);
    run<9, 14, 2, 128, 4, true>("register coefficients", nsm, y);
    run<9, 14, 4, 128, 4, true>("register coefficients", nsm, y);
    run<9, 14, 2, 128, 5, true>("register coefficients", nsm, y);
    run<9, 14, 4, 128, 5, true>("register coefficients", nsm, y);
    run<9, 14, 6, 128, 4, true>("register coefficients", nsm, y);
    run<9, 14, 8, 128, 3, true>("register coefficients", nsm, y);
    run<8, 8, 4, 128, 4, true>("register coefficients", nsm, y);
    run<8, 8, 4, 128, 5, true>("register coefficients", nsm, y);
    return 0;
}
