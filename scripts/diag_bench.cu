// Developer microbenchmark: variants of the 64 x 64 diagonal-block LU (k_diag), one CTA and many CTAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/diag_bench scripts/diag_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../nk_ocn_tracer_jacobian_precond_b200/csrc/nkp_internal.hpp"
using namespace nkp;
constexpr int NBMAX = 64;

// ---- v0: first version (block in shared memory) ----
__global__ void __launch_bounds__(256) diag_v0(const DiagTask* __restrict__ tasks, double* __restrict__ heap, double tiny, int* __restrict__ n_replaced) {
    __shared__ double D[NBMAX * (NBMAX + 1)];
    const DiagTask tk = tasks[blockIdx.x];
    const int kb = tk.kb, ld = tk.ld;
    double* G = heap + tk.Doff;
    const int LDS = NBMAX + 1;
    for (int e = threadIdx.x; e < NBMAX * NBMAX; e += blockDim.x) {
        int i = e % NBMAX, j = e / NBMAX;
        D[i + j * LDS] = (i < kb && j < kb) ? G[i + (int64_t)j * ld] : (i == j ? 1.0 : 0.0);
    }
    __syncthreads();
    const int i = threadIdx.x & 63, seg = threadIdx.x >> 6;
    for (int k = 0; k < kb; k++) {
        double p = D[k + k * LDS];
        if (fabs(p) < tiny) { p = p < 0 ? -tiny : tiny; if (threadIdx.x == 0) atomicAdd(n_replaced, 1); }
        double l = 0.0;
        const int j0 = seg * 16;
        if (i > k && i < kb && j0 + 15 > k) {
            l = D[i + k * LDS] / p;
            double own[16], piv[16];
#pragma unroll
            for (int jj = 0; jj < 16; jj++) { own[jj] = D[i + (j0 + jj) * LDS]; piv[jj] = D[k + (j0 + jj) * LDS]; }
#pragma unroll
            for (int jj = 0; jj < 16; jj++) { int j = j0 + jj; if (j > k && j < kb) D[i + j * LDS] = own[jj] - l * piv[jj]; }
        } else if (i > k && i < kb) l = D[i + k * LDS] / p;
        __syncthreads();
        if (seg == 0 && i < kb) { if (i > k) D[i + k * LDS] = l; else if (i == k) D[k + k * LDS] = p; }
    }
    __syncthreads();
    double* GU = heap + tk.UTDoff;
    for (int e = threadIdx.x; e < kb * kb; e += blockDim.x) {
        int a = e % kb, b = e / kb;
        double v = D[a + b * LDS];
        G[a + (int64_t)b * ld] = v;
        if (a <= b) GU[b + (int64_t)a * ld] = v;
    }
}

// ---- v1: register tiles, one barrier per column; RCP selects the reciprocal flavour ----
template <int RCP>
__device__ __forceinline__ double recip(double p) {
    if (RCP == 0) return 1.0 / p;
    if (RCP == 1) return __drcp_rn(p);
    double r;   // approximate reciprocal + two Newton steps
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(p));
    double e = fma(-p, r, 1.0);
    r = fma(r, e, r);
    e = fma(-p, r, 1.0);
    r = fma(r, e, r);
    return r;
}
template <int RCP>
__global__ void __launch_bounds__(256) diag_v1(const DiagTask* __restrict__ tasks, double* __restrict__ heap, double tiny, int* __restrict__ n_replaced) {
    __shared__ __align__(16) double rowb[2][NBMAX], colb[2][NBMAX];
    const DiagTask tk = tasks[blockIdx.x];
    const int kb = tk.kb, ld = tk.ld;
    double* G = heap + tk.Doff;
    const int ti = threadIdx.x & 15, tj = threadIdx.x >> 4;
    const int i0 = 4 * ti, j0 = 4 * tj;
    double a[4][4];
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int r = 0; r < 4; r++) { const int i = i0 + r, j = j0 + c; a[r][c] = (i < kb && j < kb) ? G[i + (int64_t)j * ld] : (i == j ? 1.0 : 0.0); }
    if (ti == 0)
#pragma unroll
        for (int c = 0; c < 4; c++) rowb[0][j0 + c] = a[0][c];
    if (tj == 0)
#pragma unroll
        for (int r = 0; r < 4; r++) colb[0][i0 + r] = a[r][0];
    __syncthreads();
    int nrep = 0;
    for (int k = 0; k < kb; k++) {
        const int cur = k & 1;
        double p = rowb[cur][k];
        if (fabs(p) < tiny) { p = p < 0 ? -tiny : tiny; nrep++; }
        const double inv = recip<RCP>(p);
        const int kq = k >> 2, kr = k & 3;
        double l[4], u[4];
#pragma unroll
        for (int r = 0; r < 4; r++) l[r] = colb[cur][i0 + r] * inv;
#pragma unroll
        for (int c = 0; c < 4; c++) u[c] = rowb[cur][j0 + c];
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (i0 + r > k && j0 + c > k) a[r][c] -= l[r] * u[c];
        if (tj == kq) {
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++)
                    if (c == kr) { if (i0 + r > k) a[r][c] = l[r]; else if (i0 + r == k) a[r][c] = p; }
        }
        if (k + 1 < kb) {
            const int nq = (k + 1) >> 2, nr = (k + 1) & 3;
            if (ti == nq)
#pragma unroll
                for (int r = 0; r < 4; r++)
                    if (r == nr)
#pragma unroll
                        for (int c = 0; c < 4; c++) rowb[cur ^ 1][j0 + c] = a[r][c];
            if (tj == nq)
#pragma unroll
                for (int c = 0; c < 4; c++)
                    if (c == nr)
#pragma unroll
                        for (int r = 0; r < 4; r++) colb[cur ^ 1][i0 + r] = a[r][c];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && nrep) atomicAdd(n_replaced, nrep);
    double* GU = heap + tk.UTDoff;
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int r = 0; r < 4; r++) { const int i = i0 + r, j = j0 + c; if (i < kb && j < kb) { G[i + (int64_t)j * ld] = a[r][c]; if (i <= j) GU[j + (int64_t)i * ld] = a[r][c]; } }
}

// ---- v2: ONE WARP per block, 64 x 64 in registers (lane owns rows lane and lane + 32), warp shuffles, no barriers
__global__ void __launch_bounds__(32) diag_v2(const DiagTask* __restrict__ tasks, double* __restrict__ heap, double tiny, int* __restrict__ n_replaced) {
    const DiagTask tk = tasks[blockIdx.x];
    const int kb = tk.kb, ld = tk.ld;
    double* G = heap + tk.Doff;
    const int lane = threadIdx.x;
    double a0[64], a1[64];   // rows lane, lane + 32
#pragma unroll
    for (int j = 0; j < 64; j++) {
        a0[j] = (lane < kb && j < kb) ? G[lane + (int64_t)j * ld] : (lane == j ? 1.0 : 0.0);
        a1[j] = (lane + 32 < kb && j < kb) ? G[lane + 32 + (int64_t)j * ld] : (lane + 32 == j ? 1.0 : 0.0);
    }
    int nrep = 0;
#pragma unroll
    for (int k = 0; k < 64; k++) {
        if (k < kb) {
            const int src = k & 31;
            double p = __shfl_sync(0xffffffffu, k < 32 ? a0[k] : a1[k], src);
            if (fabs(p) < tiny) { p = p < 0 ? -tiny : tiny; nrep++; }
            const double inv = 1.0 / p;
            const bool up0 = lane > k, up1 = lane + 32 > k;
            const double l0 = up0 ? a0[k] * inv : 0.0, l1 = up1 ? a1[k] * inv : 0.0;
            if (up0) a0[k] = l0;
            if (up1) a1[k] = l1;
            if (lane == src) { if (k < 32) a0[k] = p; else a1[k] = p; }
#pragma unroll
            for (int j = k + 1; j < 64; j++) {
                const double u = __shfl_sync(0xffffffffu, k < 32 ? a0[j] : a1[j], src);
                a0[j] -= l0 * u;
                a1[j] -= l1 * u;
            }
        }
    }
    if (lane == 0 && nrep) atomicAdd(n_replaced, nrep);
    double* GU = heap + tk.UTDoff;
#pragma unroll
    for (int j = 0; j < 64; j++) {
        if (j < kb) {
            if (lane < kb) { G[lane + (int64_t)j * ld] = a0[j]; if (lane <= j) GU[j + (int64_t)lane * ld] = a0[j]; }
            if (lane + 32 < kb) { G[lane + 32 + (int64_t)j * ld] = a1[j]; if (lane + 32 <= j) GU[j + (int64_t)(lane + 32) * ld] = a1[j]; }
        }
    }
}


// ---- v3: blocked right-looking LU, 8-column micro-panels: the panel is factored by ONE warp in registers (shuffles,
// no barrier inside), the row block of U by one thread per column, the rank-8 trailing update by 4 x 4 register tiles.
// 3 barriers per 8 columns instead of 8.
__global__ void __launch_bounds__(256) diag_v3(const DiagTask* __restrict__ tasks, double* __restrict__ heap, double tiny, int* __restrict__ n_replaced) {
    __shared__ double D[NBMAX * (NBMAX + 1)];
    const DiagTask tk = tasks[blockIdx.x];
    const int kb = tk.kb, ld = tk.ld;
    double* G = heap + tk.Doff;
    constexpr int LDS = NBMAX + 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < NBMAX * NBMAX; e += 256) {
        int i = e % NBMAX, j = e / NBMAX;
        D[i + j * LDS] = (i < kb && j < kb) ? G[i + (int64_t)j * ld] : (i == j ? 1.0 : 0.0);
    }
    __syncthreads();
    int nrep = 0;
    for (int j0 = 0; j0 < kb; j0 += 8) {
        if (warp == 0) {
            const int r0 = j0 + lane, r1 = j0 + lane + 32;
            double a0[8], a1[8];
#pragma unroll
            for (int c = 0; c < 8; c++) {
                a0[c] = r0 < NBMAX ? D[r0 + (j0 + c) * LDS] : 0.0;
                a1[c] = r1 < NBMAX ? D[r1 + (j0 + c) * LDS] : 0.0;
            }
#pragma unroll
            for (int c = 0; c < 8; c++) {
                double p = __shfl_sync(0xffffffffu, a0[c], c);
                if (fabs(p) < tiny) {
                    p = p < 0 ? -tiny : tiny;
                    nrep++;
                }
                const double inv = 1.0 / p;
                const double l0 = lane > c ? a0[c] * inv : 0.0, l1 = a1[c] * inv;
                if (lane == c) a0[c] = p;
                if (lane > c) a0[c] = l0;
                a1[c] = l1;
#pragma unroll
                for (int j = c + 1; j < 8; j++) {
                    const double u = __shfl_sync(0xffffffffu, a0[j], c);
                    a0[j] -= l0 * u;
                    a1[j] -= l1 * u;
                }
            }
#pragma unroll
            for (int c = 0; c < 8; c++) {
                if (r0 < NBMAX) D[r0 + (j0 + c) * LDS] = a0[c];
                if (r1 < NBMAX) D[r1 + (j0 + c) * LDS] = a1[c];
            }
        }
        __syncthreads();
        const int nrem = NBMAX - j0 - 8;   // rows / columns beyond the micro-panel
        if (tid < nrem) {   // U12 = L11^-1 A12, one column per thread
            const int j = j0 + 8 + tid;
            double x[8];
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = D[j0 + i + j * LDS];
#pragma unroll
            for (int i = 1; i < 8; i++)
#pragma unroll
                for (int k = 0; k < i; k++) x[i] -= D[j0 + i + (j0 + k) * LDS] * x[k];
#pragma unroll
            for (int i = 1; i < 8; i++) D[j0 + i + j * LDS] = x[i];
        }
        __syncthreads();
        const int nt = nrem >> 2;
        if (tid < nt * nt) {   // A22 -= L21 U12, 4 x 4 tile per thread
            const int ti = tid % nt, tj = tid / nt;
            const int i0 = j0 + 8 + 4 * ti, c0 = j0 + 8 + 4 * tj;
            double acc[4][4];
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
                for (int r = 0; r < 4; r++) acc[r][c] = D[i0 + r + (c0 + c) * LDS];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                double l[4], u[4];
#pragma unroll
                for (int r = 0; r < 4; r++) l[r] = D[i0 + r + (j0 + k) * LDS];
#pragma unroll
                for (int c = 0; c < 4; c++) u[c] = D[j0 + k + (c0 + c) * LDS];
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int c = 0; c < 4; c++) acc[r][c] -= l[r] * u[c];
            }
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
                for (int r = 0; r < 4; r++) D[i0 + r + (c0 + c) * LDS] = acc[r][c];
        }
        __syncthreads();
    }
    if (tid == 0 && nrep) atomicAdd(n_replaced, nrep);
    double* GU = heap + tk.UTDoff;
    for (int e = tid; e < kb * kb; e += 256) {
        int a = e % kb, b = e / kb;
        double v = D[a + b * LDS];
        G[a + (int64_t)b * ld] = v;
        if (a <= b) GU[b + (int64_t)a * ld] = v;
    }
}

__global__ void k_empty(const DiagTask*, double*, double, int*) {}

int main() {
    const int nblk = 2048, ld = 64;
    std::vector<double> h((size_t)nblk * 2 * 64 * 64);
    srand(1);
    for (int b = 0; b < nblk; b++)
        for (int j = 0; j < 64; j++)
            for (int i = 0; i < 64; i++) h[(size_t)b * 8192 + i + j * 64] = (i == j ? 40.0 : 0.0) + (rand() / (double)RAND_MAX - 0.5);
    std::vector<DiagTask> tasks(nblk);
    for (int b = 0; b < nblk; b++) tasks[b] = DiagTask{(int64_t)b * 8192, (int64_t)b * 8192 + 4096, ld, 64};
    double *d, *d0; DiagTask* dt; int* nr;
    cudaMalloc(&d, h.size() * 8); cudaMalloc(&d0, h.size() * 8); cudaMalloc(&dt, sizeof(DiagTask) * nblk); cudaMalloc(&nr, 4);
    cudaMemcpy(d0, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dt, tasks.data(), sizeof(DiagTask) * nblk, cudaMemcpyHostToDevice);
    cudaMemset(nr, 0, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<double> ref(h.size()), out(h.size());
    typedef void (*kern_t)(const DiagTask*, double*, double, int*);
    struct V { const char* name; kern_t k; int threads; } vs[] = {
        {"empty", k_empty, 32}, {"v0 smem", diag_v0, 256}, {"v1 reg div", diag_v1<0>, 256}, {"v1 reg drcp", diag_v1<1>, 256},
        {"v1 reg approx+newton", diag_v1<2>, 256}, {"v2 one warp", diag_v2, 32}, {"v3 blocked 8", diag_v3, 256}};
    for (auto& v : vs) {
        for (int grid : {1, 8, 2048}) {
            float best = 1e9;
            for (int rep = 0; rep < 5; rep++) {
                cudaMemcpy(d, d0, h.size() * 8, cudaMemcpyDeviceToDevice);
                cudaDeviceSynchronize();
                cudaEventRecord(e0);
                // 8 back-to-back dependent launches (same stream), like the panel loop
                for (int q = 0; q < 8; q++) v.k<<<grid, v.threads>>>(dt + (q * grid) % (nblk - grid + 1), d, 1e-8, nr);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                best = ms < best ? ms : best;
            }
            printf("%-22s grid %5d: %8.2f us per launch (8 back-to-back)  %s\n", v.name, grid, best * 1e3 / 8, cudaGetErrorString(cudaGetLastError()));
        }
        // correctness vs v0 on the first 8 blocks
        cudaMemcpy(d, d0, h.size() * 8, cudaMemcpyDeviceToDevice);
        v.k<<<8, v.threads>>>(dt, d, 1e-8, nr);
        cudaMemcpy(out.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
        if (v.k == (kern_t)diag_v0) ref = out;
        else if (v.k != (kern_t)k_empty) {
            double md = 0;
            for (size_t q = 0; q < 8 * 8192; q++) md = fmax(md, fabs(out[q] - ref[q]));
            printf("   max |diff| vs v0 on 8 blocks: %.3e\n", md);
        }
    }
    return 0;
}
