// horner_probe.cu -- development microbenchmark: what FP64-pipe utilisation does the Horner inner structure of
// qo_tf.cuh reach on sm_100a, as a function of warps per scheduler, coefficient source and operand pattern?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o qo-100-tools_b200/lib/horner_probe tools/horner_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: r = fma(r, y[p], c) with c from registers (loop-invariant pair per step parity)
// MODE 1: c from shared memory, loaded at the top of each step (no prefetch)
// MODE 2: c from shared memory, loaded one step ahead
// MODE 3: like 0 but y shared by all chains (one unique operand per DFMA, as in the DFMA-peak probe)
// MODE 4: c from shared memory, loaded TWO steps ahead          MODE 5: three steps ahead
// MODE 7: one 8-byte load per step (both chains use the same coefficient): half the bytes of MODE 2
// MODE 8: like 2 but only lane 0 loads and the pair is broadcast with two SHFL pairs (is it the register-file write-back?)
// MODE 9: like 2, two 8-byte loads per step instead of one 16-byte load
// MODE 6: like 2, with the warps of a scheduler started 0 / 1500 / 3000 / 4500 cycles apart (does a convoy matter?)
template <int MODE, int NCH, int NPT, int TPB>
__global__ void __launch_bounds__(TPB) probe(double *out, int iters, int steps, const double *yin)
{
    __shared__ __align__(16) double coef[64];
    if (threadIdx.x < 64) coef[threadIdx.x] = 1.0 + 1e-3 * threadIdx.x;
    __syncthreads();
    double y[NPT], r[NCH][NPT];
#pragma unroll
    for (int p = 0; p < NPT; p++) y[p] = yin[(threadIdx.x & 31) * NPT + p];
#pragma unroll
    for (int c = 0; c < NCH; c++)
#pragma unroll
        for (int p = 0; p < NPT; p++) r[c][p] = 0.5 + c + p;
    const unsigned sb = (unsigned)__cvta_generic_to_shared(coef);
    double acc = 0;
    if (MODE == 6) { const long long t0 = clock64(), w = (long long)((blockIdx.x / 148) % 4) * 1500; while (clock64() - t0 < w) { } }
    for (int it = 0; it < iters; it++) {
        double2 nx, n2, n3;
        if (MODE == 7) nx.y = 0.0;
        if (MODE == 2 || MODE >= 4) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(nx.x), "=d"(nx.y) : "r"(sb));
        if (MODE == 4 || MODE == 5) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(n2.x), "=d"(n2.y) : "r"(sb + 16u));
        if (MODE == 5) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(n3.x), "=d"(n3.y) : "r"(sb + 32u));
#pragma unroll 4
        for (int k = 0; k < steps; k++) {
            double2 cc;
            if (MODE == 1) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(cc.x), "=d"(cc.y) : "r"(sb + 16u * (k & 3)));
            else if (MODE == 2 || MODE == 6) { cc = nx; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(nx.x), "=d"(nx.y) : "r"(sb + 16u * ((k + 1) & 3))); }
            else if (MODE == 7) { cc.x = nx.x; cc.y = nx.x; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(nx.x) : "r"(sb + 16u * ((k + 1) & 3))); }
            else if (MODE == 9) { cc = nx; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(nx.x) : "r"(sb + 16u * ((k + 1) & 3))); asm volatile("ld.shared.f64 %0, [%1];" : "=d"(nx.y) : "r"(sb + 8u + 16u * ((k + 1) & 3))); }
            else if (MODE == 8) {
                cc = nx;
                if ((threadIdx.x & 31) == 0) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(nx.x), "=d"(nx.y) : "r"(sb + 16u * ((k + 1) & 3)));
                nx.x = __shfl_sync(0xffffffffu, nx.x, 0); nx.y = __shfl_sync(0xffffffffu, nx.y, 0);
            }
            else if (MODE == 4) { cc = nx; nx = n2; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(n2.x), "=d"(n2.y) : "r"(sb + 16u * ((k + 2) & 3))); }
            else if (MODE == 5) { cc = nx; nx = n2; n2 = n3; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(n3.x), "=d"(n3.y) : "r"(sb + 16u * ((k + 3) & 3))); }
            else { cc.x = 1.0009765625; cc.y = 0.99951171875; }
#pragma unroll
            for (int c = 0; c < NCH; c++)
#pragma unroll
                for (int p = 0; p < NPT; p++) r[c][p] = fma(r[c][p], MODE == 3 ? y[0] : y[p], (c & 1) ? cc.y : cc.x);
        }
#pragma unroll
        for (int c = 0; c < NCH; c++)
#pragma unroll
            for (int p = 0; p < NPT; p++) { acc += r[c][p] * 1e-30; r[c][p] = 0.5 + 1e-9 * acc; }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE, int NCH, int NPT, int TPB> static void run(const char *name, int blocks_per_sm, int nsm, const double *y)
{
    const int blocks = nsm * blocks_per_sm, iters = 256, steps = 16;
    double *d;
    cudaMalloc(&d, (size_t)blocks * TPB * sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        probe<MODE, NCH, NPT, TPB><<<blocks, TPB>>>(d, iters, steps, y);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double fp64 = (double)blocks * TPB * iters * ((double)steps * NCH * NPT + 2.0 * NCH * NPT);
    const double rate = fp64 / (best * 1e-3);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe<MODE, NCH, NPT, TPB>, TPB, 0);
    printf("%-44s warps/SMSP %4.1f (max resident blocks %d) chains %dx%d : %7.3f ms  %6.2f T FP64-instr/s  (%5.1f %% of 64/clk/SM @1.965GHz)\n", name,
           blocks_per_sm * (TPB / 32) / 4.0, occ, NCH, NPT, best, rate * 1e-12, rate / (148.0 * 64 * 1.965e9) * 100);
    cudaFree(d);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int nsm = p.multiProcessorCount;
    double *y; cudaMalloc(&y, 32 * 8 * sizeof(double));
    double hy[256]; for (int i = 0; i < 256; i++) hy[i] = -0.9 + 1e-3 * i;
    cudaMemcpy(y, hy, sizeof hy, cudaMemcpyHostToDevice);
    printf("%s, %d SMs\n", p.name, nsm);
    run<3, 2, 8, 128>("shared y operand (DFMA-peak pattern), regs", 4, nsm, y);
    run<0, 2, 8, 128>("Horner 2x8, coefficients in registers", 4, nsm, y);
    run<0, 2, 8, 128>("Horner 2x8, coefficients in registers", 2, nsm, y);
    run<0, 2, 8, 128>("Horner 2x8, coefficients in registers", 1, nsm, y);
    run<1, 2, 8, 128>("Horner 2x8, LDS per step, no prefetch", 4, nsm, y);
    run<2, 2, 8, 128>("Horner 2x8, LDS one step ahead", 4, nsm, y);
    run<2, 2, 8, 128>("Horner 2x8, LDS one step ahead", 2, nsm, y);
    run<2, 1, 8, 128>("Horner 1x8 (E polynomial), LDS ahead", 4, nsm, y);
    run<2, 2, 4, 128>("Horner 2x4, LDS one step ahead", 4, nsm, y);
    run<2, 2, 4, 128>("Horner 2x4, LDS one step ahead", 6, nsm, y);
    run<2, 2, 4, 128>("Horner 2x4, LDS one step ahead", 8, nsm, y);
    run<2, 4, 4, 128>("Horner 4x4, LDS one step ahead", 4, nsm, y);
    run<2, 2, 6, 128>("Horner 2x6, LDS one step ahead", 5, nsm, y);
    run<4, 2, 8, 128>("Horner 2x8, LDS two steps ahead", 4, nsm, y);
    run<5, 2, 8, 128>("Horner 2x8, LDS three steps ahead", 4, nsm, y);
    run<4, 1, 8, 128>("Horner 1x8, LDS two steps ahead", 4, nsm, y);
    run<6, 2, 8, 128>("Horner 2x8, LDS one step ahead, staggered", 4, nsm, y);
    run<7, 2, 8, 128>("Horner 2x8, one LDS.64 per step", 4, nsm, y);
    run<9, 2, 8, 128>("Horner 2x8, two LDS.64 per step", 4, nsm, y);
    run<8, 2, 8, 128>("Horner 2x8, lane-0 LDS.128 + 4 SHFL", 4, nsm, y);
    return 0;
}
