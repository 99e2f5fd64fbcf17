// Microbenchmark: peak rate of DMMA.8x8x4 (FP64 tensor pipe) vs DFMA (FP64 FMA pipe) on this GPU.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void k_dmma(double* out, int iters) {
    double c[NACC][2];
    for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = 0;
    double a = threadIdx.x * 1e-3, b = blockIdx.x * 1e-3 + 1.0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dfma(double* out, int iters) {
    double c[NACC];
    for (int i = 0; i < NACC; i++) c[i] = i;
    double a = threadIdx.x * 1e-3, b = blockIdx.x * 1e-3 + 1.0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) c[i] = fma(a, b, c[i]);
    }
    double s = 0;
    for (int i = 0; i < NACC; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// both pipes at once: even warps DMMA, odd warps DFMA
__global__ void k_mixed(double* out, int iters) {
    double c[16][2];
    for (int i = 0; i < 16; i++) c[i][0] = c[i][1] = 0;
    double a = threadIdx.x * 1e-3, b = blockIdx.x * 1e-3 + 1.0;
    if ((threadIdx.x >> 5) & 1) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++) { c[i][0] = fma(a, b, c[i][0]); c[i][1] = fma(a, b, c[i][1]); }
        }
    } else {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++) dmma(c[i][0], c[i][1], a, b);
        }
    }
    double s = 0;
    for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F>
float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    int iters = 20000;
    for (int wpb : {4, 8, 16, 32}) {
        for (int bps : {1, 2}) {
            if (wpb * bps > 64) continue;
            int threads = wpb * 32, blocks = sms * bps;
            float ms = timeit([&] { k_dmma<16><<<blocks, threads>>>(out, iters); });
            double flops = 2.0 * 256 * 16 * (double)iters * wpb * blocks;
            printf("DMMA  warps/SM=%2d : %7.2f TFLOP/s\n", wpb * bps, flops / (ms * 1e-3) * 1e-12);
            ms = timeit([&] { k_dfma<16><<<blocks, threads>>>(out, iters); });
            flops = 2.0 * 32 * 16 * (double)iters * wpb * blocks;
            printf("DFMA  warps/SM=%2d : %7.2f TFLOP/s\n", wpb * bps, flops / (ms * 1e-3) * 1e-12);
        }
    }
    {
        int threads = 512, blocks = sms * 2;
        float ms = timeit([&] { k_mixed<<<blocks, threads>>>(out, iters); });
        double fl_dmma = 2.0 * 256 * 16 * (double)iters * 8 * blocks, fl_dfma = 2.0 * 32 * 32 * (double)iters * 8 * blocks;
        printf("MIXED warps/SM=32 (16 DMMA + 16 DFMA): DMMA part %7.2f + DFMA part %7.2f = %7.2f TFLOP/s\n",
               fl_dmma / (ms * 1e-3) * 1e-12, fl_dfma / (ms * 1e-3) * 1e-12, (fl_dmma + fl_dfma) / (ms * 1e-3) * 1e-12);
    }
    return 0;
}
