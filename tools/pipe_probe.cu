// pipe_probe.cu -- development microbenchmark: how much FP64-pipe time do other instruction classes
// cost when interleaved with DFMA on sm_100a?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -o qo-100-tools_b200/lib/pipe_probe tools/pipe_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int NCH>
__global__ void __launch_bounds__(256) probe(double *out, int iters, double x, double y, unsigned k)
{
    double a[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) a[i] = threadIdx.x + i;
    unsigned u0 = threadIdx.x, u1 = k, u2 = k * 3u, u3 = 7u;
    __shared__ double sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = x;
    __syncthreads();
    double acc = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < NCH; i++) a[i] = fma(a[i], x, y);
            if (MODE == 1) {   // + NCH/2 integer ALU ops (LOP3/IADD), independent chain
#pragma unroll
                for (int i = 0; i < NCH / 2; i++) { u0 = (u0 ^ u1) + u2; u1 = (u1 & u3) | u0; }
            }
            if (MODE == 2) {   // + NCH/4 MUFU.RCP64H
#pragma unroll
                for (int i = 0; i < NCH / 4; i++) { double r0; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(a[i])); acc += r0; }
            }
            if (MODE == 3) {   // + NCH/4 LDS.64 broadcast
#pragma unroll
                for (int i = 0; i < NCH / 4; i++) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(sm + ((r + i) & 63)))); acc = fma(v, x, acc); }
            }
            if (MODE == 4) {   // DMUL instead of half the DFMA
#pragma unroll
                for (int i = 0; i < NCH; i += 2) a[i] = a[i] * x;
            }
            if (MODE == 5) {   // + NCH/2 FSEL/SEL style selects
#pragma unroll
                for (int i = 0; i < NCH / 2; i++) { u0 = (u1 > u2) ? u0 : u3; u1 = (u0 > u3) ? u1 + 1 : u2; }
            }
            if (MODE == 6) {   // + NCH/2 32-bit moves (IMAD.MOV-like): rotate registers
#pragma unroll
                for (int i = 0; i < NCH / 2; i++) { unsigned t = u0; u0 = u1; u1 = u2; u2 = u3; u3 = t + 1; }
            }
        }
    }
    double s = acc;
#pragma unroll
    for (int i = 0; i < NCH; i++) s += a[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s + (double)(u0 + u1 + u2 + u3);
}

template <int MODE, int NCH> static void run(const char *name, int blocks_per_sm, int nsm, double extra_dfma_per_round)
{
    const int blocks = nsm * blocks_per_sm, iters = 2048;
    double *d;
    cudaMalloc(&d, (size_t)blocks * 256 * sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        probe<MODE, NCH><<<blocks, 256>>>(d, iters, 0.999999, 1e-9, 12345u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double fp64_instr = (double)blocks * 256 * iters * 8 * (NCH + extra_dfma_per_round);
    const double rate = fp64_instr / (best * 1e-3);             // thread-level FP64 instr / s
    printf("%-34s warps/SMSP %2d  chains %2d : %7.3f ms  %6.2f T FP64-instr/s  (%5.1f %% of 64/clk/SM @1.965GHz)\n", name,
           blocks_per_sm * 2, NCH, best, rate * 1e-12, rate / (148.0 * 64 * 1.965e9) * 100);
    cudaFree(d);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int nsm = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, nsm);
    run<0, 8>("pure DFMA", 8, nsm, 0);
    run<0, 8>("pure DFMA", 2, nsm, 0);
    run<0, 8>("pure DFMA", 1, nsm, 0);
    run<0, 4>("pure DFMA", 2, nsm, 0);
    run<0, 2>("pure DFMA", 2, nsm, 0);
    run<0, 1>("pure DFMA (latency probe)", 1, nsm, 0);
    run<0, 16>("pure DFMA", 2, nsm, 0);
    run<1, 8>("DFMA + 1.0 int ALU per DFMA", 2, nsm, 0);
    run<5, 8>("DFMA + 1.0 select per DFMA", 2, nsm, 0);
    run<6, 8>("DFMA + 2.0 mov per DFMA", 2, nsm, 0);
    run<2, 8>("DFMA + 0.25 MUFU.RCP64H (+.25 DADD)", 2, nsm, 2);
    run<3, 8>("DFMA + 0.25 LDS (+.25 DFMA)", 2, nsm, 2);
    run<4, 8>("DFMA + 0.5 DMUL", 2, nsm, 4);
    return 0;
}
