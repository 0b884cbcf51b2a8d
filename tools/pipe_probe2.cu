// pipe_probe2.cu -- does an independent non-FP64 instruction issued next to DFMA steal FP64-pipe time on sm_100a?
// 8 independent DFMA chains + R independent "other" instructions per 8 DFMA, 4 warps per SMSP.
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND, int R>
__global__ void __launch_bounds__(256) probe(double *out, int iters, double x, double y, unsigned k)
{
    double a[8];
    unsigned u[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x + i; u[i] = threadIdx.x * 7u + i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                a[i] = fma(a[i], x, y);
                if (i < R) {
                    if (KIND == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(k), "r"(u[(i + 1) & 7]));
                    if (KIND == 2) asm volatile("{ .reg .pred p; setp.gt.u32 p, %1, 5; selp.b32 %0, %0, %2, p; }" : "+r"(u[i]) : "r"(k), "r"(u[(i + 1) & 7]));
                    if (KIND == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(k), "r"(u[(i + 1) & 7]));
                    if (KIND == 4) asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
                    if (KIND == 5) asm volatile("mov.b32 %0, %1;" : "=r"(u[i]) : "r"(u[(i + 1) & 7]));
                    if (KIND == 6) { float f; asm volatile("fma.rn.f32 %0, %1, %1, %1;" : "=f"(f) : "f"(__uint_as_float(u[i]))); u[i] = __float_as_uint(f); }
                }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i] + (double)u[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND, int R> static void run(const char *name, int nsm)
{
    const int blocks = nsm * 2, iters = 4096;
    double *d;
    cudaMalloc(&d, (size_t)blocks * 256 * sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        probe<KIND, R><<<blocks, 256>>>(d, iters, 0.999999, 1e-9, 12345u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double rate = (double)blocks * 256 * iters * 64 / (best * 1e-3);
    printf("%-10s %d per 8 DFMA : %7.3f ms  %6.2f T DFMA/s  (%5.1f %% of 64/clk/SM @1.965GHz)\n", name, R, best, rate * 1e-12,
           rate / (148.0 * 64 * 1.965e9) * 100);
    cudaFree(d);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int nsm = p.multiProcessorCount;
    run<0, 0>("none", nsm);
    run<1, 2>("lop3", nsm); run<1, 4>("lop3", nsm); run<1, 8>("lop3", nsm);
    run<2, 2>("setp+selp", nsm); run<2, 4>("setp+selp", nsm); run<2, 8>("setp+selp", nsm);
    run<3, 2>("imad", nsm); run<3, 4>("imad", nsm); run<3, 8>("imad", nsm);
    run<4, 4>("prmt", nsm); run<4, 8>("prmt", nsm);
    run<5, 4>("mov", nsm); run<5, 8>("mov", nsm);
    run<6, 4>("ffma", nsm); run<6, 8>("ffma", nsm);
    return 0;
}
