// horner_probe3.cu -- development microbenchmark: THREAD-per-sample frequency loop.  Each thread keeps its own sample's
// polynomial coefficients in registers (compile-time lengths); the grid value y is the same for the whole warp and comes from
// constant memory; P points are in flight per thread for ILP.  No shared-memory traffic in the loop at all.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o qo-100-tools_b200/lib/horner_probe3 tools/horner_probe3.cu
#include <cstdio>
#include <cuda_runtime.h>

__constant__ double c_y[4096];

template <int KN, int KE, int P, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) probe(double *out, int nbatch, int nf, double thr, const double *coef_src)
{
    unsigned int fails = 0;
    for (int b = 0; b < nbatch; b++) {
        double cn[2 * KN], ce[KE];
#pragma unroll
        for (int k = 0; k < 2 * KN; k++) cn[k] = coef_src[(k * 37 + threadIdx.x + b) & 1023];
#pragma unroll
        for (int k = 0; k < KE; k++) ce[k] = coef_src[(k * 53 + threadIdx.x + 7 * b) & 1023];
        unsigned int acc = 0;
        for (int j = 0; j < nf; j += P) {
            double y[P], re[P], ro[P], dd[P];
#pragma unroll
            for (int p = 0; p < P; p++) y[p] = c_y[j + p];
#pragma unroll
            for (int p = 0; p < P; p++) { re[p] = fma(cn[2 * KN - 2], y[p], cn[2 * KN - 4]); ro[p] = fma(cn[2 * KN - 1], y[p], cn[2 * KN - 3]); dd[p] = fma(ce[KE - 1], y[p], ce[KE - 2]); }
#pragma unroll
            for (int k = KN - 3; k >= 0; k--)
#pragma unroll
                for (int p = 0; p < P; p++) { re[p] = fma(re[p], y[p], cn[2 * k]); ro[p] = fma(ro[p], y[p], cn[2 * k + 1]); }
#pragma unroll
            for (int k = KE - 3; k >= 0; k--)
#pragma unroll
                for (int p = 0; p < P; p++) dd[p] = fma(dd[p], y[p], ce[k]);
#pragma unroll
            for (int p = 0; p < P; p++) {
                const double t = ro[p] * ro[p], n2 = fma(-y[p], t, re[p] * re[p]);
                acc |= (unsigned int)__double2hiint(fma(thr, dd[p], -n2));
            }
        }
        fails += acc >> 31;
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = fails;
}

template <int KN, int KE, int P, int TPB, int MINB> static void run(int nsm, const double *src)
{
    const int blocks = nsm * MINB, nbatch = 2, nf = 4096;
    double *d;
    cudaMalloc(&d, (size_t)blocks * TPB * sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        probe<KN, KE, P, TPB, MINB><<<blocks, TPB>>>(d, nbatch, nf, 0.37, src);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double per_pt = 2.0 * (KN - 1) + (KE - 1) + 4.0;
    const double evals = (double)blocks * TPB * nbatch * nf;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, probe<KN, KE, P, TPB, MINB>);
    printf("thread-per-sample KN %2d KE %2d P %d  %3d regs, %d x %d thr/SM: %7.3f ms  %5.1f %% FP64 pipe  %.3e evals/s\n", KN, KE, P, fa.numRegs, MINB, TPB, best,
           evals * per_pt / (best * 1e-3) / (148.0 * 64 * 1.965e9) * 100, evals / (best * 1e-3));
    cudaFree(d);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int nsm = p.multiProcessorCount;
    static double hy[4096]; for (int i = 0; i < 4096; i++) hy[i] = -0.9 + 2e-4 * i;
    cudaMemcpyToSymbol(c_y, hy, sizeof hy);
    double *src; cudaMalloc(&src, 1024 * sizeof(double));
    static double hs[1024]; for (int i = 0; i < 1024; i++) hs[i] = 1.0 / (1.0 + i);
    cudaMemcpy(src, hs, sizeof hs, cudaMemcpyHostToDevice);
    printf("%s, %d SMs\n", p.name, nsm);
    run<9, 14, 1, 128, 4>(nsm, src);
    run<9, 14, 2, 128, 4>(nsm, src);
    run<9, 14, 3, 128, 4>(nsm, src);
    run<9, 14, 4, 128, 4>(nsm, src);
    run<9, 14, 2, 128, 5>(nsm, src);
    run<9, 14, 2, 128, 6>(nsm, src);
    run<9, 14, 4, 128, 5>(nsm, src);
    run<9, 14, 2, 64, 8>(nsm, src);
    run<8, 8, 4, 128, 4>(nsm, src);
    return 0;
}
