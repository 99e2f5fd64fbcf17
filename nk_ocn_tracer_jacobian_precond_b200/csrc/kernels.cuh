// kernels.cuh -- sm_100a kernels of the multifrontal numeric phase and the solves.
//
// Replaces the inside of SuperLU_DIST's pdgstrf / pdgstrs / pdgsrfs as called by the
// reference (src/SuperLU_brief_tree.txt:11-24).  Every kernel is driven by a static task
// list built by the host analysis (nkp_internal.hpp); a launch covers ALL fronts of one
// assembly-tree level, so the launch count does not depend on the number of fronts.
//
//   k_scatter        CRS (scaled) -> front storage                       HBM bound
//   k_extend_add     child update matrix -> parent front                 HBM bound
//   k_diag           nb x nb diagonal-block LU, static pivoting          latency bound
//   k_trsm           tall panel  X * T = B  (L panel and U^T panel)      FP64 FMA
//   k_gemm           Schur update C -= A * B^T, FP64 tensor cores (DMMA) FP64 tensor
//   k_invert_diag    64 x 64 diagonal blocks -> their inverses (for the sweeps) latency bound
//   k_sweep_big      level-scheduled sweeps of the big fronts: many CTAs,  HBM bound
//                    DMMA tile stream, counter-driven dataflow, multi-RHS
//   k_fwd_small /    sweeps of the small fronts, one warp per front        HBM / latency
//   k_bwd_small
//   k_gather_fields / k_scatter_fields   tracer fields <-> right-hand sides HBM bound
//   k_bswap64        matrix-file ingest (big-endian doubles)               HBM bound
//   k_residual       r = b - A x and |A||x|+|b| (refinement, berr)       HBM bound
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nkp_internal.hpp"

namespace nkp {

// ------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------

// largest t in [0,n) with key(t) <= idx, keys ascending, key(0) == 0
template <class T, class F>
__device__ __forceinline__ int find_task(const T* tasks, int n, int idx, F key) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (key(tasks[mid]) <= idx) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ int64_t d_front_entry(int a, int b, int s, int m, int ld, int nb,
                                                 int64_t Loff, int64_t UToff, int64_t F22off) {
    if (b < s) {
        if (a >= s || a / nb >= b / nb) return Loff + a + (int64_t)b * ld;
        return UToff + b + (int64_t)a * ld;
    }
    if (a < s) return UToff + b + (int64_t)a * ld;
    return F22off + (a - s) + (int64_t)(b - s) * (m - s);
}

__device__ __forceinline__ void atomic_max_pos_double(double* addr, double v) {
    // valid for non-negative doubles: their bit patterns order like unsigned integers
    atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool pred) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
// same with a precomputed shared-space address and source size (8 = copy, 0 = zero-fill)
__device__ __forceinline__ void cp_async8_s(unsigned saddr, const void* gmem, int sz) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(saddr), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

// D(8x8) += A(8x4) * B(4x8), FP64 tensor core
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------
// equilibration (dgsequ-style row / column scalings, rounded to powers of two)
// ------------------------------------------------------------------------------------------

__global__ void k_row_scale(int n, const int* __restrict__ rowptr, const double* __restrict__ val,
                            double* __restrict__ R) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double mx = 0;
    for (int p = rowptr[i]; p < rowptr[i + 1]; p++) mx = fmax(mx, fabs(val[p]));
    double r = 1.0;
    if (mx > 0) {
        int e;
        frexp(mx, &e);           // mx = f * 2^e, f in [0.5,1)
        r = ldexp(1.0, 1 - e);   // r * mx in [1,2)
    }
    R[i] = r;
}

__global__ void k_col_max(int64_t nnz, const int* __restrict__ rowidx, const int* __restrict__ colind,
                          const double* __restrict__ val, const double* __restrict__ R,
                          double* __restrict__ cmax) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    double v = fabs(val[p]) * R[rowidx[p]];
    atomic_max_pos_double(&cmax[colind[p]], v);
}

__global__ void k_col_scale(int n, double* __restrict__ C) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double mx = C[i];
    double c = 1.0;
    if (mx > 0) {
        int e;
        frexp(mx, &e);
        c = ldexp(1.0, 1 - e);
    }
    C[i] = c;
}

// big-endian IEEE doubles (NetCDF-3 external representation, src/matrix.c:3880) -> host-order doubles;
// src may alias dst (every thread converts its own element)
__global__ void __launch_bounds__(256) k_bswap64(const unsigned long long* src, double* dst, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long v = src[i];
    const unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
    const unsigned long long w = ((unsigned long long)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
    dst[i] = __longlong_as_double((long long)w);
}

__global__ void k_fill(double* p, int64_t n, double v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ------------------------------------------------------------------------------------------
// CRS -> front scatter
// ------------------------------------------------------------------------------------------

__global__ void k_scatter(int64_t nnz, const double* __restrict__ val, const int64_t* __restrict__ dst,
                          const int* __restrict__ rowidx, const int* __restrict__ colind,
                          const double* __restrict__ R, const double* __restrict__ C,
                          double* __restrict__ heap) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    const int64_t d = dst[p];
    if (d < 0) return;   // entry of a front owned by another GPU
    double v = val[p] * R[rowidx[p]] * C[colind[p]];
    atomicAdd(&heap[d], v);   // duplicates in the CRS (if any) are summed, as sum_dup_vals would
}

// ------------------------------------------------------------------------------------------
// extend-add
// ------------------------------------------------------------------------------------------

constexpr int ADD_TILE = 32;

__global__ void __launch_bounds__(256) k_extend_add(const AddTask* __restrict__ tasks, int ntasks,
                                                    const int* __restrict__ rel, double* __restrict__ heap,
                                                    int nb) {
    int t = find_task(tasks, ntasks, (int)blockIdx.x, [](const AddTask& x) { return x.tile0; });
    const AddTask tk = tasks[t];
    int local = blockIdx.x - tk.tile0;
    int ti = local % tk.tiles_m, tj = local / tk.tiles_m;
    const int* rl = rel + tk.rel_off;
    const double* Cc = heap + tk.Coff;
    int a = ti * ADD_TILE + (threadIdx.x & 31);
    int ra = a < tk.rc ? rl[a] : 0;
    // 4 entries per thread: all addresses and loads first, then the stores (distinct targets)
    int64_t d[4];
    double v[4], o[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        int b = tj * ADD_TILE + (threadIdx.x >> 5) + 8 * q;
        d[q] = -1;
        if (a < tk.rc && b < tk.rc) {
            int rb = rl[b];
            v[q] = Cc[a + (int64_t)b * tk.rc];
            d[q] = d_front_entry(ra, rb, tk.sp, tk.mp, tk.ldp, nb, tk.Loff, tk.UToff, tk.F22off);
        }
    }
#pragma unroll
    for (int q = 0; q < 4; q++)
        if (d[q] >= 0) o[q] = heap[d[q]];
#pragma unroll
    for (int q = 0; q < 4; q++)
        if (d[q] >= 0) heap[d[q]] = o[q] + v[q];
}

// ------------------------------------------------------------------------------------------
// diagonal block LU (kb <= 64), static pivoting with tiny-pivot replacement
// ------------------------------------------------------------------------------------------

constexpr int NBMAX = 64;

// Blocked right-looking LU with 8-column micro-panels: the micro-panel (rows below included) is factored by ONE warp in
// registers (lane = two rows, pivot row broadcast by shuffles, no barrier inside), the 8 x rest row block of U by one
// thread per column, the rank-8 trailing update by 4 x 4 register tiles: 3 CTA barriers per 8 columns.
// scripts/diag_bench.cu measured four designs on B200 (one CTA, 8 dependent launches, per launch): column-by-column in
// shared memory 35 us (round 1), 4 x 4 register tiles with one barrier per column 46-50 us, one warp holding the
// whole block 56 us, this one 27 us.  The floor is the chain of dependent FP64 operations per pivot (update of the
// next pivot, reciprocal, multiplier, update), which costs several hundred cycles on this part.
__global__ void __launch_bounds__(256) k_diag(const DiagTask* __restrict__ tasks, double* __restrict__ heap,
                                              double tiny, int* __restrict__ n_replaced) {
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

// ------------------------------------------------------------------------------------------
// panel solve  X * T = B,  T upper triangular kb x kb (kb <= 64), one thread per row
// ------------------------------------------------------------------------------------------

constexpr int TRSM_ROWS = 128;

// X * T = B on a 128-row slab with 256 threads: thread r (< 128) solves columns 0..31 of row r,
// thread 128 + r solves columns 32..63 of the same row after a rank-32 update with the first
// half (exchanged through shared memory).  Each thread keeps only 32 columns in registers, so
// twice as many warps are resident as with one thread per full row and the rank-32 update is
// 32 independent FMA chains.
constexpr int TRSM_THREADS = 256;
constexpr int TRSM_XLD = 33;
constexpr int TRSM_SMEM = (NBMAX * NBMAX + TRSM_ROWS * TRSM_XLD) * 8;

__global__ void __launch_bounds__(TRSM_THREADS, 2) k_trsm(const TrsmTask* __restrict__ tasks, int ntasks,
                                                          double* __restrict__ heap) {
    extern __shared__ __align__(16) double tsm[];
    double* T = tsm;                       // T[p + j*NBMAX], p <= j ; diagonal holds 1/T_jj
    double* xs = tsm + NBMAX * NBMAX;      // xs[r * TRSM_XLD + p]: first half of the solution
    int t = find_task(tasks, ntasks, (int)blockIdx.x, [](const TrsmTask& x) { return x.cta0; });
    const TrsmTask tk = tasks[t];
    const int kb = tk.kb, ld = tk.ld;
    const double* G = heap + tk.Toff;
    for (int e = threadIdx.x; e < NBMAX * NBMAX; e += TRSM_THREADS) {
        int p = e % NBMAX, j = e / NBMAX;
        double v = 0.0;
        if (p < kb && j < kb) {
            if (tk.unit == 0) {
                if (p < j) v = G[p + (int64_t)j * ld];
                else if (p == j) v = 1.0 / G[p + (int64_t)p * ld];
            } else {
                if (p < j) v = G[j + (int64_t)p * ld];   // L_kk^T
                else if (p == j) v = 1.0;
            }
        } else if (p == j) v = 1.0;
        T[e] = v;
    }
    const int half = threadIdx.x >> 7;                 // 0: columns 0..31, 1: columns 32..63
    const int r = threadIdx.x & (TRSM_ROWS - 1);
    const int row = (blockIdx.x - tk.cta0) * TRSM_ROWS + r;
    const bool active = row < tk.nrows && (half == 0 || kb > 32);
    double* X = heap + tk.Xoff + row;
    const int c0 = half * 32;
    double x[32];
#pragma unroll
    for (int j = 0; j < 32; j++) x[j] = (active && c0 + j < kb) ? X[(int64_t)(c0 + j) * ld] : 0.0;
    __syncthreads();
    if (half == 0) {
#pragma unroll
        for (int j = 0; j < 32; j++) {
            double a0 = x[j], a1 = 0.0;
#pragma unroll
            for (int p = 0; p < j; p++) {
                const double tt = x[p] * T[p + j * NBMAX];
                if (p & 1) a1 -= tt;
                else a0 -= tt;
            }
            x[j] = (a0 + a1) * T[j + j * NBMAX];
        }
#pragma unroll
        for (int j = 0; j < 32; j++) xs[r * TRSM_XLD + j] = x[j];
        if (active)
#pragma unroll
            for (int j = 0; j < 32; j++)
                if (j < kb) X[(int64_t)j * ld] = x[j];
    }
    __syncthreads();
    if (half == 1 && active) {
        // rank-32 update with the first half, then the second 32 x 32 triangle
        for (int p = 0; p < 32; p++) {
            const double xp = xs[r * TRSM_XLD + p];
#pragma unroll
            for (int j = 0; j < 32; j++) x[j] -= xp * T[p + (32 + j) * NBMAX];
        }
#pragma unroll
        for (int j = 0; j < 32; j++) {
            double a0 = x[j], a1 = 0.0;
#pragma unroll
            for (int p = 0; p < j; p++) {
                const double tt = x[p] * T[32 + p + (32 + j) * NBMAX];
                if (p & 1) a1 -= tt;
                else a0 -= tt;
            }
            x[j] = (a0 + a1) * T[32 + j + (32 + j) * NBMAX];
        }
#pragma unroll
        for (int j = 0; j < 32; j++)
            if (32 + j < kb) X[(int64_t)(32 + j) * ld] = x[j];
    }
}

// ------------------------------------------------------------------------------------------
// Schur update  C[M x N] -= A[M x K] * B[N x K]^T   (K <= 64), FP64 DMMA m8n8k4
// CTA tile 128 x 64, 8 warps (4 x 2), warp tile 32 x 32.
// ------------------------------------------------------------------------------------------

constexpr int G_TM = 128, G_TN = 64;
#ifndef NKP_G_KC
#define NKP_G_KC 16
#endif
#ifndef NKP_G_ST
#define NKP_G_ST 2
#endif
constexpr int G_KC = NKP_G_KC;    // K chunk of one pipeline stage
constexpr int G_ST = NKP_G_ST;    // stages of the cp.async ring (prefetch distance G_ST - 1 chunks).  16 x 2 was the
                                  // fastest of 14 measured (KC, ST) pairs at gx1v6 (3.26 s; 16 x 4: 3.63 s): the ring
                                  // takes 51 KB per CTA instead of 102 KB and leaves the rest of the SM's 256 KB to L1
constexpr int G_LDA = G_TM + 4;   // == 4 (mod 16): conflict-free 8-byte fragment loads
constexpr int G_LDB = G_TN + 4;
constexpr int G_LDC = G_TM + 2;   // == 2 (mod 16): conflict-free accumulator staging
constexpr int G_STAGE = G_KC * (G_LDA + G_LDB);   // doubles per stage
constexpr int G_SMEM = G_ST * G_STAGE * 8;
static_assert(32 * G_LDC <= G_ST * G_STAGE, "accumulator staging (32 columns at a time) must fit in the ring");

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }

// mbarrier helpers (shared::cta): the ring is a producer/consumer pipeline without CTA-wide
// barriers -- "full" completes when the cp.async copies of all 256 threads have landed,
// "empty" when all 256 threads have finished reading the stage.
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    unsigned done = 0;
    // try_wait suspends for a hardware-defined time before returning false, so this is not a hot
    // spin; the iteration cap turns a protocol bug into a trapped kernel instead of a hung GPU
    for (unsigned spins = 0; !done; spins++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
        if (spins > (1u << 26)) __trap();
    }
}

__global__ void __launch_bounds__(256, 2) k_gemm(const GemmTask* __restrict__ tasks, int ntasks,
                                                 double* __restrict__ heap, int nb) {
    extern __shared__ __align__(16) double smem[];
    __shared__ uint64_t full_bar[G_ST], empty_bar[G_ST];
    int t = find_task(tasks, ntasks, (int)blockIdx.x, [](const GemmTask& x) { return x.tile0; });
    const GemmTask tk = tasks[t];
    int local = blockIdx.x - tk.tile0;
    int ti = local % tk.tiles_m, tj = local / tk.tiles_m;
    int rend = (ti + 1) * G_TM;
    if (tk.skip == 1 && rend <= (tj * G_TN / nb) * nb) return;
    if (tk.skip == 2 && rend <= min((tj * G_TN / nb + 1) * nb, tk.N)) return;
    const int m0 = ti * G_TM, n0 = tj * G_TN;
    const double* A = heap + tk.Aoff + m0;
    const double* B = heap + tk.Boff + n0;
    const int mrem = tk.M - m0, nrem = tk.N - n0, K = tk.K;
    double* C = heap + tk.Coff + m0 + (int64_t)n0 * tk.ldc;
    const int nchunks = (K + G_KC - 1) / G_KC;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < G_ST; q++) {
            mbar_init(&full_bar[q], 256);
            mbar_init(&empty_bar[q], 256);
        }
    }
    __syncthreads();

    // stage one K chunk of A (128 x 16) and B (64 x 16), zero-filling out-of-range elements.
    // Every thread copies a fixed row of 8 (A) + 4 (B) columns: addresses advance by constant strides.
    const int ia = threadIdx.x & (G_TM - 1), ka0 = threadIdx.x >> 7;   // A: row, first column (0..1), step 2
    const int ib = threadIdx.x & (G_TN - 1), kb0 = threadIdx.x >> 6;   // B: row, first column (0..3), step 4
    const bool aok = ia < mrem, bok = ib < nrem;
    const double* pa0 = A + (aok ? ia : 0) + (int64_t)ka0 * tk.lda;
    const double* pb0 = B + (bok ? ib : 0) + (int64_t)kb0 * tk.ldb;
    const int64_t sa = 2 * (int64_t)tk.lda, sb = 4 * (int64_t)tk.ldb;
    auto issue = [&](int ch) {
        const int stg = ch % G_ST;
        double* As = smem + stg * G_STAGE;        // As[k * G_LDA + m]
        double* Bs = As + G_KC * G_LDA;           // Bs[k * G_LDB + n]
        const int kbase = ch * G_KC;
        const double* pa = pa0 + (int64_t)kbase * tk.lda;
        const double* pb = pb0 + (int64_t)kbase * tk.ldb;
        const unsigned da = (unsigned)__cvta_generic_to_shared(As + ka0 * G_LDA + ia);
        const unsigned db = (unsigned)__cvta_generic_to_shared(Bs + kb0 * G_LDB + ib);
        if (kbase + G_KC <= K) {
            const int sza = aok ? 8 : 0, szb = bok ? 8 : 0;
#pragma unroll
            for (int q = 0; q < G_KC / 2; q++) cp_async8_s(da + q * 2 * G_LDA * 8, pa + q * sa, sza);
#pragma unroll
            for (int q = 0; q < G_KC / 4; q++) cp_async8_s(db + q * 4 * G_LDB * 8, pb + q * sb, szb);
        } else {
#pragma unroll
            for (int q = 0; q < G_KC / 2; q++) {
                bool ok = aok && (kbase + ka0 + 2 * q < K);
                cp_async8_s(da + q * 2 * G_LDA * 8, ok ? pa + q * sa : A, ok ? 8 : 0);
            }
#pragma unroll
            for (int q = 0; q < G_KC / 4; q++) {
                bool ok = bok && (kbase + kb0 + 4 * q < K);
                cp_async8_s(db + q * 4 * G_LDB * 8, ok ? pb + q * sb : B, ok ? 8 : 0);
            }
        }
        mbar_cp_async_arrive(&full_bar[stg]);
    };
#pragma unroll
    for (int q = 0; q < G_ST - 1; q++)
        if (q < nchunks) issue(q);
    // pull the C tile towards L2 for the epilogue: 512 lines of 128 B, 2 per thread
    for (int e = threadIdx.x; e < G_TN * 8; e += 256) {
        int j = e >> 3, seg = e & 7;
        if (j < nrem && seg * 16 < mrem) prefetch_l2(C + seg * 16 + (int64_t)j * tk.ldc);
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
    const int lr = lane >> 2, lc = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;

    for (int ch = 0; ch < nchunks; ch++) {
        const int stg = ch % G_ST;
        mbar_wait(&full_bar[stg], (ch / G_ST) & 1);
        // refill the stage that chunk ch-1 used, once every thread has finished reading it
        const int nxt = ch + G_ST - 1;
        if (nxt < nchunks) {
            if (nxt >= G_ST) mbar_wait(&empty_bar[nxt % G_ST], (nxt / G_ST - 1) & 1);
            issue(nxt);
        }
        const double* As = smem + stg * G_STAGE;
        const double* Bs = As + G_KC * G_LDA;
        const int ksteps = min(G_KC, K - ch * G_KC + 3) >> 2;
#pragma unroll
        for (int ks = 0; ks < G_KC / 4; ks++) {
            if (ks < ksteps) {
                const double* ap = As + (ks * 4 + lc) * G_LDA + wm + lr;
                const double* bp = Bs + (ks * 4 + lc) * G_LDB + wn + lr;
                double af[4], bf[4];
#pragma unroll
                for (int a = 0; a < 4; a++) af[a] = ap[a * 8];
#pragma unroll
                for (int b = 0; b < 4; b++) bf[b] = bp[b * 8];
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b < 4; b++) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
            }
        }
        mbar_arrive(&empty_bar[stg]);
    }
    __syncthreads();
    // stage the product through shared memory so that the read-modify-write of C is coalesced;
    // 32 columns at a time, so the staging area (and with it the ring) stays small: the less shared
    // memory the kernel takes, the more L1 is left, which measurably helps
    double* Cs = smem;   // Cs[n * G_LDC + m], n in [0, 32)
    constexpr int PER = 16;   // 16 elements per thread and half
    const int i = threadIdx.x % G_TM, jb = threadIdx.x / G_TM;   // jb in {0,1}
#pragma unroll
    for (int half = 0; half < 2; half++) {
        if (half) __syncthreads();
        if (wn == 32 * half) {
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    int r = wm + a * 8 + lr;
                    int c = b * 8 + 2 * lc;
                    Cs[c * G_LDC + r] = acc[a][b][0];
                    Cs[(c + 1) * G_LDC + r] = acc[a][b][1];
                }
        }
        __syncthreads();
        // all loads of C first (independent, deep memory-level parallelism), then update + store
        double cv[PER];
#pragma unroll
        for (int q = 0; q < PER; q++) {
            int j = 32 * half + jb + 2 * q;
            cv[q] = (i < mrem && j < nrem) ? C[i + (int64_t)j * tk.ldc] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < PER; q++) {
            int j = 32 * half + jb + 2 * q;
            if (i < mrem && j < nrem) C[i + (int64_t)j * tk.ldc] = cv[q] - Cs[(jb + 2 * q) * G_LDC + i];
        }
    }
}

// ------------------------------------------------------------------------------------------
// right-hand-side handling
// ------------------------------------------------------------------------------------------

// y[perm[i], c] = R[i] * b[i, c]
__global__ void k_permute_in(int n, int nrhs, const int* __restrict__ perm, const double* __restrict__ R,
                             const double* __restrict__ b, int ldb, double* __restrict__ y) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int pi = perm[i];
    double r = R[i];
    for (int c = 0; c < nrhs; c++) y[pi + (int64_t)c * n] = r * b[i + (int64_t)c * ldb];
}

// x[i, c] (+)= C[i] * y[perm[i], c]
// columns whose bit is clear in `mask` are left alone: a right-hand side that has met the refinement
// stop rule is frozen, so a batched solve gives every column exactly what a single solve would
__global__ void k_permute_out(int n, int nrhs, const int* __restrict__ perm, const double* __restrict__ Cs,
                              const double* __restrict__ y, double* __restrict__ x, int ldx, int accumulate,
                              unsigned mask) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int pi = perm[i];
    double cs = Cs[i];
    for (int c = 0; c < nrhs; c++) {
        if (!((mask >> c) & 1u)) continue;
        double v = cs * y[pi + (int64_t)c * n];
        if (accumulate) x[i + (int64_t)c * ldx] += v;
        else x[i + (int64_t)c * ldx] = v;
    }
}

// ------------------------------------------------------------------------------------------
// tracer fields <-> right-hand sides (get_B_global / put_B_global of the reference,
// src/solve_ABglobal.c:184-191 and :242-248): system `col` stacks the ocean points of `ct`
// coupled tracer fields, B[t * tsl + s, col] = field_{col * ct + t}[cell[s]].  The scatter
// writes only ocean points, so land values of the fields survive (src/solve_ABglobal.c:236-248).
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) k_gather_fields(int tsl, int ct, int nfields, const int* __restrict__ cell,
                                                       const double* __restrict__ fields, int64_t ncell,
                                                       double* __restrict__ B, int ldb) {
    const int sidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (sidx >= tsl) return;
    const int c = cell[sidx];
    for (int f = 0; f < nfields; f++) {
        const int col = f / ct, t = f - col * ct;
        B[(int64_t)t * tsl + sidx + (int64_t)col * ldb] = fields[(int64_t)f * ncell + c];
    }
}

__global__ void __launch_bounds__(256) k_scatter_fields(int tsl, int ct, int nfields, const int* __restrict__ cell,
                                                        double* __restrict__ fields, int64_t ncell,
                                                        const double* __restrict__ B, int ldb) {
    const int sidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (sidx >= tsl) return;
    const int c = cell[sidx];
    for (int f = 0; f < nfields; f++) {
        const int col = f / ct, t = f - col * ct;
        fields[(int64_t)f * ncell + c] = B[(int64_t)t * tsl + sidx + (int64_t)col * ldb];
    }
}

// ------------------------------------------------------------------------------------------
// triangular sweeps of the small fronts: ONE WARP per front, no barriers.
// W (work vectors) layout: front t, rhs c  ->  W[woff_t * nrtot + c * m_t + a]
//
// The arithmetic is the same block algorithm as for the big fronts (inverted 64 x 64 diagonal
// blocks, everything as DMMA m8n8k4 products with the right-hand sides as the N = 8 dimension),
// but a warp walks its whole front alone: the A fragments come straight from global memory into
// registers (32 independent loads in flight per lane), the work vector lives in global memory
// (L2 resident), and results move from the accumulator layout to the B-fragment layout with
// warp shuffles.  With 8 fronts per CTA and no shared memory, thousands of fronts are in flight.
// ------------------------------------------------------------------------------------------

#ifndef NKP_SMALL_MINB
#define NKP_SMALL_MINB 3
#endif
constexpr int SMALL_WARPS = 4;

// accumulator layout (row 8g + lr, rhs 2 lc + {0,1}) -> B fragment of k-step ks: (k = 4 ks + lc, rhs = lr)
__device__ __forceinline__ double acc_to_bfrag(const double (&acc)[8][2], int ks, int lr, int lc) {
    const int q = 4 * ks + lc;
    const int src = ((q & 7) << 2) | (lr >> 1);
    const double v0 = __shfl_sync(0xffffffffu, acc[ks >> 1][0], src);
    const double v1 = __shfl_sync(0xffffffffu, acc[ks >> 1][1], src);
    return (lr & 1) ? v1 : v0;
}

__global__ void __launch_bounds__(32 * SMALL_WARPS, NKP_SMALL_MINB) k_fwd_small(const SolveTask* __restrict__ tasks, int ntasks,
                                                                const SolveChild* __restrict__ children,
                                                                const int* __restrict__ rel,
                                                                const double* __restrict__ heap, double* W, double* y,
                                                                int n, int nr, int nrtot) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, lr = lane >> 2, lc = lane & 3;
    const int t = blockIdx.x * SMALL_WARPS + warp;
    if (t >= ntasks) return;
    const SolveTask tk = tasks[t];
    const int s = tk.s, m = tk.m, ld = tk.ld;
    double* w = W + tk.woff * nrtot;
    const double* L = heap + tk.Loff;
    // 1. pivots' right-hand side, zero boundary part
    for (int a = lane; a < m; a += 32)
        for (int c = 0; c < nr; c++) w[a + (int64_t)c * m] = a < s ? y[tk.first + a + (int64_t)c * n] : 0.0;
    __syncwarp();
    // 2. children's update vectors (fixed order: deterministic)
    for (int ch = 0; ch < tk.nchild; ch++) {
        const SolveChild sc = children[tk.child_list + ch];
        const int mc = sc.s + sc.r;
        const double* wc = W + sc.woff * nrtot + sc.s;
        const int* rl = rel + sc.rel_off;
        for (int a = lane; a < sc.r; a += 32) {
            const int d = rl[a];
            for (int c = 0; c < nr; c++) w[d + (int64_t)c * m] += wc[a + (int64_t)c * mc];
        }
        __syncwarp();
    }
    // 3. block forward substitution
    for (int c0 = 0; c0 < s; c0 += 64) {
        const int kb = min(64, s - c0);
        // y_i = L_ii^-1 v_i
        double acc[8][2];
        {
            double bf[16];
#pragma unroll
            for (int ks = 0; ks < 16; ks++) {
                const int q = 4 * ks + lc;
                bf[ks] = (q < kb && lr < nr) ? __ldcg(&w[c0 + q + (int64_t)lr * m]) : 0.0;
            }
#pragma unroll
            for (int g = 0; g < 8; g++) {
                acc[g][0] = acc[g][1] = 0.0;
                const int p = 8 * g + lr;
                double af[16];
#pragma unroll
                for (int ks = 0; ks < 16; ks++) {
                    const int q = 4 * ks + lc;
                    double v = p == q ? 1.0 : 0.0;
                    if (ks <= 2 * g + 1 && p < kb && q < p) v = L[c0 + p + (int64_t)(c0 + q) * ld];
                    af[ks] = v;
                }
#pragma unroll
                for (int ks = 0; ks < 16; ks++)
                    if (ks <= 2 * g + 1) dmma884(acc[g][0], acc[g][1], af[ks], bf[ks]);
            }
        }
#pragma unroll
        for (int g = 0; g < 8; g++) {
            const int p = 8 * g + lr;
            if (p < kb) {
                if (2 * lc < nr) y[tk.first + c0 + p + (int64_t)(2 * lc) * n] = acc[g][0];
                if (2 * lc + 1 < nr) y[tk.first + c0 + p + (int64_t)(2 * lc + 1) * n] = acc[g][1];
            }
        }
        // rows below: w -= L[rows, block] y_i
        double bneg[16];
#pragma unroll
        for (int ks = 0; ks < 16; ks++) bneg[ks] = -acc_to_bfrag(acc, ks, lr, lc);
        const int R0 = c0 + kb;
        const int ksn = (kb + 3) >> 2;
        for (int rb = R0; rb < m; rb += 16) {
            double cc[2][2], af[2][16];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int R = rb + 8 * h + lr;
                const bool rok = R < m;
                cc[h][0] = (rok && 2 * lc < nr) ? __ldcg(&w[R + (int64_t)(2 * lc) * m]) : 0.0;
                cc[h][1] = (rok && 2 * lc + 1 < nr) ? __ldcg(&w[R + (int64_t)(2 * lc + 1) * m]) : 0.0;
#pragma unroll
                for (int ks = 0; ks < 16; ks++) {
                    const int q = 4 * ks + lc;
                    af[h][ks] = (rok && q < kb) ? L[R + (int64_t)(c0 + q) * ld] : 0.0;
                }
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
#pragma unroll
                for (int ks = 0; ks < 16; ks++)
                    if (ks < ksn) dmma884(cc[h][0], cc[h][1], af[h][ks], bneg[ks]);
                const int R = rb + 8 * h + lr;
                if (R < m) {
                    if (2 * lc < nr) w[R + (int64_t)(2 * lc) * m] = cc[h][0];
                    if (2 * lc + 1 < nr) w[R + (int64_t)(2 * lc + 1) * m] = cc[h][1];
                }
            }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(32 * SMALL_WARPS, NKP_SMALL_MINB) k_bwd_small(const SolveTask* __restrict__ tasks, int ntasks,
                                                                const int* __restrict__ bidx,
                                                                const double* __restrict__ heap, double* W, double* y,
                                                                int n, int nr, int nrtot) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, lr = lane >> 2, lc = lane & 3;
    const int t = blockIdx.x * SMALL_WARPS + warp;
    if (t >= ntasks) return;
    const SolveTask tk = tasks[t];
    const int s = tk.s, m = tk.m, ld = tk.ld;
    double* w = W + tk.woff * nrtot;
    const double* UT = heap + tk.UToff;
    const int* bi = bidx + tk.bidx_off;
    // 1. boundary values (solutions of ancestors)
    for (int a = lane; a < tk.r; a += 32) {
        const int g = bi[a];
        for (int c = 0; c < nr; c++) w[s + a + (int64_t)c * m] = y[g + (int64_t)c * n];
    }
    __syncwarp();
    // 2. block back substitution, last pivot block first
    for (int c0 = (s - 1) / 64 * 64; c0 >= 0; c0 -= 64) {
        const int kb = min(64, s - c0);
        const int R0 = c0 + kb;
        // z_i = y_i - UT[rows below, block]^T x_rows : out = 64 columns (8 groups), contraction over rows
        double acc[8][2];
#pragma unroll
        for (int g = 0; g < 8; g++) acc[g][0] = acc[g][1] = 0.0;
        for (int rb = R0; rb < m; rb += 16) {
            double bf[4], af[4][8];
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                const int R = rb + 4 * kk + lc;
                const bool rok = R < m;
                bf[kk] = (rok && lr < nr) ? __ldcg(&w[R + (int64_t)lr * m]) : 0.0;
#pragma unroll
                for (int g = 0; g < 8; g++) af[kk][g] = (rok && 8 * g + lr < kb) ? UT[R + (int64_t)(c0 + 8 * g + lr) * ld] : 0.0;
            }
#pragma unroll
            for (int kk = 0; kk < 4; kk++)
#pragma unroll
                for (int g = 0; g < 8; g++) dmma884(acc[g][0], acc[g][1], af[kk][g], bf[kk]);
        }
#pragma unroll
        for (int g = 0; g < 8; g++) {
            const int p = 8 * g + lr;
            const bool ok = p < kb;
            const double y0 = (ok && 2 * lc < nr) ? y[tk.first + c0 + p + (int64_t)(2 * lc) * n] : 0.0;
            const double y1 = (ok && 2 * lc + 1 < nr) ? y[tk.first + c0 + p + (int64_t)(2 * lc + 1) * n] : 0.0;
            acc[g][0] = y0 - acc[g][0];
            acc[g][1] = y1 - acc[g][1];
        }
        // x_i = U_ii^-1 z_i   (U_ii^-1 upper, stored transposed: UT[c0 + q + (c0 + p) ld] = Uinv(p, q), q >= p)
        double zf[16];
#pragma unroll
        for (int ks = 0; ks < 16; ks++) zf[ks] = acc_to_bfrag(acc, ks, lr, lc);
#pragma unroll
        for (int g = 0; g < 8; g++) {
            const int p = 8 * g + lr;
            double x0 = 0.0, x1 = 0.0;
            double af[16];
#pragma unroll
            for (int ks = 0; ks < 16; ks++) {
                const int q = 4 * ks + lc;
                double v = p == q ? 1.0 : 0.0;
                if (ks >= 2 * g && p < kb && q < kb) v = q >= p ? UT[c0 + q + (int64_t)(c0 + p) * ld] : 0.0;
                af[ks] = v;
            }
#pragma unroll
            for (int ks = 0; ks < 16; ks++)
                if (ks >= 2 * g) dmma884(x0, x1, af[ks], zf[ks]);
            if (p < kb) {
                if (2 * lc < nr) {
                    y[tk.first + c0 + p + (int64_t)(2 * lc) * n] = x0;
                    w[c0 + p + (int64_t)(2 * lc) * m] = x0;
                }
                if (2 * lc + 1 < nr) {
                    y[tk.first + c0 + p + (int64_t)(2 * lc + 1) * n] = x1;
                    w[c0 + p + (int64_t)(2 * lc + 1) * m] = x1;
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// small fronts, second generation: PERSISTENT CTAs, one front at a time per CTA, the panel streamed through shared
// memory by 1-D bulk async copies (cp.async.bulk + mbarrier complete_tx -- the TMA engine without a tensor map: every
// column of a panel is contiguous and 16-byte aligned because ld is even).
//
// A front is cut into JOBS; a job is what one ring stage holds: cc consecutive columns of pivot block k, rows from the
// top of the block's diagonal block to the end of the front (one bulk copy per column, 1-6 KB each).
//   forward : blocks and columns ascending.  The inverted diagonal block is lower triangular, so the columns of a job
//             complete y_k for exactly their own rows (yk accumulates the contributions of the columns seen so far), and
//             the rows below the block can be updated with those columns at once.
//   backward: blocks and columns descending.  The job's rows below the block give z for its own columns; (U_kk^-1)(p, .)
//             lies in column p and needs z of the same and of later columns, which are complete.
// Thread 0 issues the copies of job q + SF_NST - 1 while the CTA works on job q; the job stream runs across front
// boundaries (factor data are read-only), so the next front's first stages are already in flight when a front ends.
// The work vector of the front (m x 8 right-hand sides) stays in shared memory; all products are DMMA m8n8k4 with
// the right-hand sides as the N dimension, as in the other sweep kernels.  Two CTA barriers per job.
// The first generation (k_fwd_small / k_bwd_small: one warp per front, fragments straight from global memory) was
// latency bound: 0.8-2.3 TB/s at gx1v6-shape, 0.15 ms for a level of 20 fronts at gx3v7-shape.
// ------------------------------------------------------------------------------------------

// Two launch shapes of the same kernels:
//   wide   256 threads, 4 stages of 36 KB, persistent (one CTA per SM): levels with few, larger fronts -- the pipeline
//          inside a front and across consecutive fronts hides the latency;
//   narrow 128 threads, 2 stages of 12 KB, one CTA per front, 3-4 CTAs per SM: levels with thousands of leaf-sized
//          fronts, where several fronts in flight per SM hide each other's start-up and barrier latencies.
constexpr int SF_MAXNST = 4;               // ring stages (upper bound)
constexpr int SF_STAGE_WIDE = 4608, SF_NST_WIDE = 4, SF_STAGE_NARROW = 1536, SF_NST_NARROW = 2;
constexpr int SF_MAXLD = 1024;             // fronts up to this leading dimension (work vector and >= 4 columns per wide stage)
constexpr int SF_NARROW_MAXLD = 320;       // ... for the narrow shape (>= 4 columns per narrow stage)

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// column stride of a staged chunk with `rows` (even) rows: >= rows, == 4 (mod 16): conflict-light fragment loads both ways
__host__ __device__ __forceinline__ int sf_cs(int rows) { return ((rows + 11) & ~15) + 4; }

// Job iterator over the fronts of one CTA (task indices t0, t0 + stride, ...).  Used twice per CTA with identical
// results: by the producer (thread 0, running ahead) and by all threads as consumers.
template <bool FWD>
struct SfJobs {
    const SolveTask* tasks;
    int ntasks, stride;
    int t;                 // current task index (>= ntasks: exhausted)
    int s, m, ld;
    int64_t base;          // Loff (forward) / UToff (backward)
    int k0, kb;            // current pivot block
    int nch, ich;          // chunks of the block, current chunk
    int j0, cc;            // first column of the chunk inside the block, number of columns
    int cs, ccmax;
    int stage, cccap;      // doubles per ring stage; largest number of columns per job (8 per warp of the CTA)
    bool first_of_front;

    __device__ void load_front() {
        if (t >= ntasks) return;
        const SolveTask& tk = tasks[t];
        s = tk.s;
        m = tk.m;
        ld = tk.ld;
        base = FWD ? tk.Loff : tk.UToff;
        k0 = FWD ? 0 : ((s - 1) / 64) * 64;
        enter_block();
        first_of_front = true;
    }
    __device__ void enter_block() {
        kb = min(64, s - k0);
        cs = sf_cs(ld - k0);
        ccmax = min(cccap, (stage / cs) & ~3);
        nch = (kb + ccmax - 1) / ccmax;
        ich = FWD ? 0 : nch - 1;
        set_chunk();
    }
    __device__ void set_chunk() {
        j0 = ich * ccmax;
        cc = min(ccmax, kb - j0);
    }
    __device__ bool valid() const { return t < ntasks; }
    __device__ bool first_of_block() const { return FWD ? ich == 0 : ich == nch - 1; }
    __device__ bool last_job() const { return FWD ? (k0 + 64 >= s && ich == nch - 1) : (k0 == 0 && ich == 0); }
    __device__ void next() {
        first_of_front = false;
        if (FWD) {
            if (++ich < nch) {
                set_chunk();
                return;
            }
            if (k0 + 64 < s) {
                k0 += 64;
                enter_block();
                return;
            }
        } else {
            if (--ich >= 0) {
                set_chunk();
                return;
            }
            if (k0 > 0) {
                k0 -= 64;
                enter_block();
                return;
            }
        }
        t += stride;
        load_front();
    }
};

// copies of one job into a ring stage (called by thread 0 only): cc columns, rows [k0, ld)
template <bool FWD>
__device__ __forceinline__ void sf_issue(const SfJobs<FWD>& jb, const double* __restrict__ heap, double* stage, uint64_t* bar) {
    const double* src = heap + jb.base + jb.k0 + (int64_t)(jb.k0 + jb.j0) * jb.ld;
    const unsigned bytes = (unsigned)((jb.ld - jb.k0) * 8);
    mbar_expect_tx(bar, (unsigned)jb.cc * bytes);
    for (int c = 0; c < jb.cc; c++) bulk_g2s(stage + c * jb.cs, src + (int64_t)c * jb.ld, bytes, bar);
}

// FORWARD.  W, y as in the other sweep kernels.  Dynamic shared memory: w[wcap][8] | yk[64][8] | ring.
template <int SF_THREADS, int MINB>
__global__ void __launch_bounds__(SF_THREADS, MINB) k_fwd_front(const SolveTask* __restrict__ tasks, int ntasks,
                                                             const SolveChild* __restrict__ children,
                                                             const int* __restrict__ rel, const double* __restrict__ heap,
                                                             double* W, double* y, int n, int nr, int nrtot, int wcap,
                                                             int SF_NST, int SF_STAGE) {
    constexpr int NW = SF_THREADS / 32;
    extern __shared__ __align__(128) double sfm[];
    __shared__ uint64_t full_bar[SF_MAXNST];
    double* w = sfm;                       // w[a * 8 + c]
    double* yk = w + (size_t)wcap * 8;     // y of the current pivot block: rows below the current chunk still accumulate
    double* ring = yk + 512;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
    if (tid == 0) {
        for (int q = 0; q < SF_NST; q++) mbar_init(&full_bar[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    SfJobs<true> prod, jb;
    prod.tasks = jb.tasks = tasks;
    prod.ntasks = jb.ntasks = ntasks;
    prod.stride = jb.stride = gridDim.x;
    prod.t = jb.t = blockIdx.x;
    prod.stage = jb.stage = SF_STAGE;
    prod.cccap = jb.cccap = 8 * NW;
    prod.load_front();
    jb.load_front();
    unsigned issued = 0, q = 0;
    if (tid == 0)
        for (; (int)issued < SF_NST - 1 && prod.valid(); issued++, prod.next())
            sf_issue(prod, heap, ring + (issued % SF_NST) * SF_STAGE, &full_bar[issued % SF_NST]);
    for (; jb.valid(); jb.next(), q++) {
        const SolveTask tk = tasks[jb.t];
        const int s = tk.s, m = tk.m;
        if (jb.first_of_front) {
            // right-hand side of the pivots, zero boundary part, then the children's update vectors (fixed order)
            for (int a = tid; a < m; a += SF_THREADS)
#pragma unroll
                for (int c = 0; c < 8; c++) w[a * 8 + c] = (a < s && c < nr) ? y[tk.first + a + (int64_t)c * n] : 0.0;
            for (int ch = 0; ch < tk.nchild; ch++) {
                __syncthreads();
                const SolveChild sc = children[tk.child_list + ch];
                const int mc = sc.s + sc.r;
                const double* wc = W + sc.woff * nrtot + sc.s;
                const int* rl = rel + sc.rel_off;
                for (int a = tid; a < sc.r; a += SF_THREADS) {
                    const int d = rl[a];
                    for (int c = 0; c < nr; c++) w[d * 8 + c] += wc[a + (int64_t)c * mc];
                }
            }
        }
        if (jb.first_of_block())
            for (int e = tid; e < 512; e += SF_THREADS) yk[e] = 0.0;
        __syncthreads();
        // keep the ring full: the stage of job q - 1 is free (barrier at the end of that job)
        if (tid == 0 && prod.valid()) {
            fence_proxy_async();
            sf_issue(prod, heap, ring + (issued % SF_NST) * SF_STAGE, &full_bar[issued % SF_NST]);
            issued++;
            prod.next();
        }
        const double* st = ring + (q % SF_NST) * SF_STAGE;
        mbar_wait(&full_bar[q % SF_NST], (q / SF_NST) & 1);
        const int k0 = jb.k0, kb = jb.kb, j0 = jb.j0, cc = jb.cc, cs = jb.cs;
        const int nks = (cc + 3) >> 2;
        // (1) yk[p] += sum over the job's columns q of Linv(p, q) v[q] for the rows p >= j0 of the block
        //     (unit lower triangular inverse: Linv(p, p) = 1, Linv(p, q > p) = 0); warp -> rows 8 warp .. 8 warp + 7
        for (int g8 = warp; g8 < 8; g8 += NW) {
            if (8 * g8 + 7 < j0) continue;
            const int p = 8 * g8 + lr;
            double a0 = yk[p * 8 + 2 * lc], a1 = yk[p * 8 + 2 * lc + 1];
#pragma unroll
            for (int ks = 0; ks < 16; ks++) {
                if (ks < nks) {
                    const int col = 4 * ks + lc, qq = j0 + col;
                    double av = 0.0;
                    if (col < cc && p < kb) av = p == qq ? 1.0 : (p > qq ? st[col * cs + p] : 0.0);
                    const double bv = col < cc ? w[(k0 + qq) * 8 + lr] : 0.0;
                    dmma884(a0, a1, av, bv);
                }
            }
            yk[p * 8 + 2 * lc] = a0;
            yk[p * 8 + 2 * lc + 1] = a1;
            if (p >= j0 && p < j0 + cc) {   // final for the rows of this job's own columns
                if (2 * lc < nr) y[tk.first + k0 + p + (int64_t)(2 * lc) * n] = a0;
                if (2 * lc + 1 < nr) y[tk.first + k0 + p + (int64_t)(2 * lc + 1) * n] = a1;
            }
        }
        __syncthreads();
        // (2) rows below the pivot block: w -= L[rows, j0 .. j0 + cc) y_k[j0 ..]
        const int R0 = k0 + kb;
        if (m > R0) {
            double bneg[16];
#pragma unroll
            for (int ks = 0; ks < 16; ks++) {
                const int col = 4 * ks + lc;
                bneg[ks] = (ks < nks && col < cc) ? -yk[(j0 + col) * 8 + lr] : 0.0;
            }
            const int ngr = (m - R0 + 7) >> 3;
            for (int g = warp; g < ngr; g += NW) {
                const int row = R0 + 8 * g + lr;
                const bool rok = row < m;
                double c0 = rok ? w[row * 8 + 2 * lc] : 0.0, c1 = rok ? w[row * 8 + 2 * lc + 1] : 0.0;
                const double* ap = st + (row - k0);
#pragma unroll
                for (int ks = 0; ks < 16; ks++) {
                    if (ks < nks) {
                        const int col = 4 * ks + lc;
                        const double av = (rok && col < cc) ? ap[col * cs] : 0.0;
                        dmma884(c0, c1, av, bneg[ks]);
                    }
                }
                if (rok) {
                    w[row * 8 + 2 * lc] = c0;
                    w[row * 8 + 2 * lc + 1] = c1;
                }
            }
        }
        if (jb.last_job()) {
            __syncthreads();
            // boundary part of the work vector -> global (the parent's extend-add reads it)
            double* wg = W + tk.woff * nrtot;
            for (int a = s + tid; a < m; a += SF_THREADS)
                for (int c = 0; c < nr; c++) wg[a + (int64_t)c * m] = w[a * 8 + c];
        }
        __syncthreads();
    }
}

// BACKWARD.  Dynamic shared memory: w[wcap][8] | zk[64][8] | red[8][64] | ring.
template <int SF_THREADS, int MINB>
__global__ void __launch_bounds__(SF_THREADS, MINB) k_bwd_front(const SolveTask* __restrict__ tasks, int ntasks,
                                                             const int* __restrict__ bidx, const double* __restrict__ heap,
                                                             double* W, double* y, int n, int nr, int nrtot, int wcap,
                                                             int SF_NST, int SF_STAGE) {
    constexpr int NW = SF_THREADS / 32;
    extern __shared__ __align__(128) double sfm[];
    __shared__ uint64_t full_bar[SF_MAXNST];
    double* w = sfm;
    double* zk = w + (size_t)wcap * 8;     // z of the current pivot block (rows of finished columns)
    double* red = zk + 512;                // one 8 x 8 partial block per warp
    double* ring = red + 512;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
    if (tid == 0) {
        for (int q = 0; q < SF_NST; q++) mbar_init(&full_bar[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    SfJobs<false> prod, jb;
    prod.tasks = jb.tasks = tasks;
    prod.ntasks = jb.ntasks = ntasks;
    prod.stride = jb.stride = gridDim.x;
    prod.t = jb.t = blockIdx.x;
    prod.stage = jb.stage = SF_STAGE;
    prod.cccap = jb.cccap = 8 * NW;
    prod.load_front();
    jb.load_front();
    unsigned issued = 0, q = 0;
    if (tid == 0)
        for (; (int)issued < SF_NST - 1 && prod.valid(); issued++, prod.next())
            sf_issue(prod, heap, ring + (issued % SF_NST) * SF_STAGE, &full_bar[issued % SF_NST]);
    for (; jb.valid(); jb.next(), q++) {
        const SolveTask tk = tasks[jb.t];
        const int s = tk.s, m = tk.m;
        if (jb.first_of_front) {
            // boundary values (solutions of the ancestors); the pivot part is filled block by block
            const int* bi = bidx + tk.bidx_off;
            for (int a = tid; a < tk.r; a += SF_THREADS) {
                const int g = bi[a];
#pragma unroll
                for (int c = 0; c < 8; c++) w[(s + a) * 8 + c] = c < nr ? y[g + (int64_t)c * n] : 0.0;
            }
            __syncthreads();
        }
        if (tid == 0 && prod.valid()) {
            fence_proxy_async();
            sf_issue(prod, heap, ring + (issued % SF_NST) * SF_STAGE, &full_bar[issued % SF_NST]);
            issued++;
            prod.next();
        }
        const double* st = ring + (q % SF_NST) * SF_STAGE;
        mbar_wait(&full_bar[q % SF_NST], (q / SF_NST) & 1);
        const int k0 = jb.k0, kb = jb.kb, j0 = jb.j0, cc = jb.cc, cs = jb.cs;
        const int R0 = k0 + kb;
        const int ng = (cc + 7) >> 3;                          // column groups of 8 in this job (<= 8)
        // (1) z[p] = y[p] - sum over the rows below the block of UT[row, p] x_row, for the job's columns p:
        //     warp = (column group, row slice); the slices are added through shared memory in a fixed order
        const int nsl = ng <= 1 ? NW : (ng <= 2 ? NW / 2 : (ng <= 4 ? (NW >= 4 ? NW / 4 : 1) : 1));   // row slices per group (ng <= NW)
        {
            const int grp = warp / nsl, sl = warp - grp * nsl;
            double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
            if (grp < ng && m > R0) {
                const int p = 8 * grp + lr;                        // column inside the job
                const int nst4 = (m - R0 + 3) >> 2;                // steps of 4 rows
                const double* ap = st + (p < cc ? p : 0) * cs + kb;
                for (int kk = sl; kk < nst4; kk += 2 * nsl) {
                    {
                        const int row = R0 + 4 * kk + lc;
                        const double av = (p < cc && row < m) ? ap[4 * kk + lc] : 0.0;
                        const double bv = row < m ? w[row * 8 + lr] : 0.0;
                        dmma884(a0, a1, av, bv);
                    }
                    const int k2 = kk + nsl;
                    if (k2 < nst4) {
                        const int row = R0 + 4 * k2 + lc;
                        const double av = (p < cc && row < m) ? ap[4 * k2 + lc] : 0.0;
                        const double bv = row < m ? w[row * 8 + lr] : 0.0;
                        dmma884(b0, b1, av, bv);
                    }
                }
                a0 += b0;
                a1 += b1;
            }
            red[warp * 64 + lr * 8 + 2 * lc] = a0;
            red[warp * 64 + lr * 8 + 2 * lc + 1] = a1;
        }
        __syncthreads();
        for (int e = tid; e < ng * 64; e += SF_THREADS) {
            const int g2 = e >> 6, rc = e & 63;
            const int p = 8 * g2 + (rc >> 3), c = rc & 7;
            if (p < cc) {
                double sum = 0.0;
                for (int s2 = 0; s2 < nsl; s2++) sum += red[(g2 * nsl + s2) * 64 + rc];
                zk[(j0 + p) * 8 + c] = (c < nr ? y[tk.first + k0 + j0 + p + (int64_t)c * n] : 0.0) - sum;
            }
        }
        __syncthreads();
        // (2) x[p] = sum over q >= p of Uinv(p, q) z[q] for the job's columns p; Uinv(p, q) is row q of staged column p
        if (warp < ng) {
            const int pl = 8 * warp + lr, p = j0 + pl;             // column inside the job / inside the block
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < 16; ks++) {
                if (4 * ks + 3 >= j0 + 8 * warp) {
                    const int qq = 4 * ks + lc;
                    double av = 0.0;
                    if (pl < cc && qq < kb && qq >= p) av = st[pl * cs + qq];
                    const double bv = (qq < kb && qq >= j0) ? zk[qq * 8 + lr] : 0.0;   // rows before j0: not computed yet
                    dmma884(a0, a1, av, bv);
                }
            }
            if (pl < cc) {
                w[(k0 + p) * 8 + 2 * lc] = a0;
                w[(k0 + p) * 8 + 2 * lc + 1] = a1;
                if (2 * lc < nr) y[tk.first + k0 + p + (int64_t)(2 * lc) * n] = a0;
                if (2 * lc + 1 < nr) y[tk.first + k0 + p + (int64_t)(2 * lc + 1) * n] = a1;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// big fronts: many CTAs per front, counter-driven dataflow (launched cooperatively so that all
// CTAs are co-resident; a CTA only ever waits on work items that precede its own in the item
// list, which some running CTA is processing or has processed).
//
//   forward : item = 64-row slab i of L.        v_i = b_i - sum_{k<i} L[i,k] y_k ;  y_i = L_ii^-1 v_i
//   backward: item = 64-column pivot panel i of U^T, swept from the last panel down.
//             z_i = y_i - sum_{row blocks below} UT[rows, i]^T x_rows ;  x_i = U_ii^-1 z_i
//
// Both are streams of 64 x 64 tiles contracted with 64 x nrhs blocks of the running solution, and
// both run on the FP64 tensor cores (DMMA m8n8k4, N = 8 right-hand sides): a warp owns 8 of the 64
// contraction indices of every tile (columns of L / rows of U^T), fetches that 8 x 64 sub-tile
// with cp.async into a warp-private shared-memory ring (no CTA barrier in the stream loop) and
// keeps a 64 x 8 partial product in 16 registers; the eight partials are added once per item.
// The diagonal blocks are stored INVERTED (k_invert_diag, after the factorisation), so the
// triangular solve at the end of an item is one more 64 x 64 x 8 product instead of a 64-step
// substitution: the dependency chain of a front has npiv links of a few microseconds each.
// Progress is published through one 64-bit counter per front and direction: (epoch << 32 | blocks
// done); the blocks of a front complete in order, so a consumer far behind the wavefront learns
// about many finished blocks with one poll.
// ------------------------------------------------------------------------------------------

constexpr int SW_LD = 68;                   // column stride of a staged tile: == 4 (mod 16) -> conflict-free fragment
                                            // loads in both directions, and room for 64 + 2 rows (alignment shift)
#ifndef NKP_SW_ST
#define NKP_SW_ST 3
#endif
constexpr int SW_ST = NKP_SW_ST;            // ring stages
constexpr int SW_TILE = 64 * SW_LD;         // doubles per stage: tile[col * SW_LD + (row - aligned first row)]
constexpr int SW_SMEM = SW_ST * SW_TILE * 8;   // dynamic shared memory per CTA (bytes): 2 CTAs per SM
static_assert(SW_ST * SW_TILE >= 8 * 64 * 8, "the ring doubles as the buffer of the eight 64 x 8 partial sums");

__device__ __forceinline__ void cp_async16_s(unsigned saddr, const void* gmem, int sz) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(saddr), "l"(gmem), "r"(sz));
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// x boundary values of the big fronts of one level -> their work vectors (backward sweep):
// w[s + a, c] = y[bidx[a], c].  grid (fronts, row chunks)
__global__ void __launch_bounds__(256) k_gather_bnd(const BigFront* __restrict__ bfs, const int* __restrict__ bidx,
                                                    const double* __restrict__ y, double* __restrict__ W, int n,
                                                    int nr, int nrtot) {
    const BigFront bf = bfs[blockIdx.x];
    const int a = blockIdx.y * 256 + threadIdx.x;
    if (a >= bf.r) return;
    const int g = bidx[bf.bidx_off + a];
    double* w = W + bf.woff * nrtot + bf.s + a;
    for (int c = 0; c < nr; c++) w[(int64_t)c * bf.m] = y[g + (int64_t)c * n];
}

// MODE 0: forward sweep (slab items).
// MODE 1: backward sweep, triangular part (panel items, pivot row blocks only); starts from
//         y_i minus the partial products of the rectangular part.
// MODE 2: backward sweep, rectangular part: item = (panel, chunk of BWD_CHUNK boundary row blocks);
//         all inputs are known when the level starts, so these items are independent and equally
//         sized; each stores its 64 x 8 partial product in a scratch slot (summed in a fixed order
//         by the MODE 1 item of the panel: deterministic).
enum { SWEEP_FWD = 0, SWEEP_BWD_TRI = 1, SWEEP_BWD_RECT = 2 };

template <int MODE>
__global__ void __launch_bounds__(256, 2) k_sweep_big(const BigFront* __restrict__ bfs, const BigItem* __restrict__ items,
                                                      int nitems, const SolveChild* __restrict__ children,
                                                      const int* __restrict__ rel, const int* __restrict__ clo,
                                                      const double* __restrict__ heap, double* W, double* y, double* part,
                                                      int n, int nr, int nrtot, unsigned long long* cnt, unsigned epoch) {
    constexpr bool FWD = MODE == SWEEP_FWD;
    constexpr bool RECT = MODE == SWEEP_BWD_RECT;
    extern __shared__ __align__(16) double ring[];
    __shared__ double bv[64 * 8];   // item start values b_i / y_i, then v_i / z_i : [entry * 8 + rhs]
    __shared__ uint64_t full_bar[SW_ST], empty_bar[SW_ST];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < SW_ST; q++) {
            mbar_init(&full_bar[q], 256);
            mbar_init(&empty_bar[q], 256);
        }
    }
    __syncthreads();
    unsigned gbase = 0;   // tiles streamed by this CTA before the current item (ring position / barrier phase)
    for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
        const BigItem item = items[it];
        const BigFront bf = bfs[item.front];
        const int s = bf.s, m = bf.m, ld = bf.ld;
        const double* F = heap + (FWD ? bf.Loff : bf.UToff);
        double* w = W + bf.woff * nrtot;
        const unsigned long long* mycnt = cnt + item.front;
        bool pivot = true;
        int i = item.idx, r0, nrow, ntiles, tile0 = 0;
        if (FWD) {
            pivot = i < bf.npiv;
            r0 = pivot ? 64 * i : s + 64 * (i - bf.npiv);
            nrow = min(64, (pivot ? s : m) - r0);
            ntiles = pivot ? i : bf.npiv;
        } else if (RECT) {
            const int chunk = item.idx % bf.nchunk;
            i = item.idx / bf.nchunk;
            tile0 = chunk * BWD_CHUNK;                       // first boundary row block of the chunk
            ntiles = min(BWD_CHUNK, ((bf.r + 63) >> 6) - tile0);
            r0 = 64 * i;
            nrow = min(64, s - r0);
        } else {
            r0 = 64 * i;
            nrow = min(64, s - r0);
            ntiles = bf.npiv - 1 - i;
        }
        // ---- start values of the 64 out entries (kept in shared memory until the end of the item)
        if (!RECT && tid < 64) {
            double base[8];
#pragma unroll
            for (int c = 0; c < 8; c++) base[c] = 0.0;
            if (tid < nrow) {
                if (pivot)
#pragma unroll
                    for (int c = 0; c < 8; c++)
                        if (c < nr) base[c] = y[bf.first + r0 + tid + (int64_t)c * n];
                if (!FWD) {
                    // partial products of the rectangular part, chunk by chunk
                    const double* pp = part + ((bf.part_off + (int64_t)i * bf.nchunk) * 64 + tid) * 8;
                    for (int ch = 0; ch < bf.nchunk; ch++) {
#pragma unroll
                        for (int c = 0; c < 8; c++) base[c] -= pp[c];
                        pp += 512;
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 8; c++) bv[tid * 8 + c] = base[c];
        }
        if (FWD && bf.nchild > 0) {
            // children's update vectors (fixed order: deterministic).  rel[] of a child is ascending,
            // so the entries that map into this slab are the run starting at the precomputed
            // child_lo[child][slab]; thread k looks at the k-th entry of the run.
            for (int ch = 0; ch < bf.nchild; ch++) {
                __syncthreads();
                if (tid < 64) {
                    const SolveChild sc = children[bf.child_list + ch];
                    const int e = clo[bf.clo_off + (int64_t)ch * bf.nslab + item.idx] + tid;
                    if (e < sc.r) {
                        const int row = rel[sc.rel_off + e] - r0;
                        if (row < nrow) {
                            const int mc = sc.s + sc.r;
                            const double* wc = W + sc.woff * nrtot + sc.s + e;
#pragma unroll
                            for (int c = 0; c < 8; c++)
                                if (c < nr) bv[row * 8 + c] += wc[(int64_t)c * mc];
                        }
                    }
                }
            }
        }
        // ---- fragments of the inverted diagonal block for the final product: entry (p, q) with
        //      p = 8 warp + lr (out), q = 4 ks + lc (contraction)
        double ainv[16];
        if (!RECT && pivot) {
            const int p = 8 * warp + lr;
#pragma unroll
            for (int ks = 0; ks < 16; ks++) {
                const int q = 4 * ks + lc;
                double v = (p == q) ? 1.0 : 0.0;
                if (p < nrow && q < nrow) {
                    if (FWD) {
                        if (p > q) v = F[r0 + p + (int64_t)(r0 + q) * ld];            // L_ii^-1 (unit lower)
                    } else {
                        v = q >= p ? F[r0 + q + (int64_t)(r0 + p) * ld] : 0.0;        // U_ii^-1 (upper), stored transposed
                    }
                }
                ainv[ks] = v;
            }
        }
        // ---- tile stream.  Tile t covers rows [R0, R0 + nrows) x columns [C0, C0 + ncols) of F:
        //        forward : R0 = r0 (out rows),          C0 = 64 t (contraction)
        //        backward: R0 = row block (contraction), C0 = r0   (out columns)
        auto tile_geom = [&](int t, int& R0, int& nrows, int& C0, int& ncols) {
            if (FWD) {
                R0 = r0;
                nrows = nrow;
                C0 = 64 * t;
                ncols = min(64, s - C0);
            } else {
                if (RECT) {
                    R0 = s + 64 * (tile0 + t);
                    nrows = min(64, m - R0);
                } else {
                    R0 = 64 * (bf.npiv - 1 - t);
                    nrows = min(64, s - R0);
                }
                C0 = r0;
                ncols = nrow;
            }
        };
        // The whole CTA stages the tile with 16-byte cp.async: a warp copies 512 contiguous bytes of
        // one column per instruction.  Columns start 16-byte aligned (even ld), so the copy starts at
        // the even row Ra <= R0 and the fragment loads skip `R0 - Ra` (0 or 1) rows.
        auto issue = [&](int t) {
            const unsigned g = gbase + (unsigned)t;
            const int stg = g % SW_ST;
            if (g >= SW_ST) mbar_wait(&empty_bar[stg], (g / SW_ST - 1) & 1);
            int R0, nrows, C0, ncols;
            tile_geom(t, R0, nrows, C0, ncols);
            const int Ra = R0 & ~1;
            const int nval = (R0 - Ra) + nrows;     // rows [Ra, Ra + nval) are wanted (the first may be a dummy)
            const unsigned st = ring_s + (unsigned)(stg * SW_TILE * 8);
            const int cr = tid & 31, cg = tid >> 5;
            const int szr = 2 * cr + 1 < nval ? 16 : (2 * cr < nval ? 8 : 0);
            const double* src = F + Ra + 2 * cr + (int64_t)(C0 + cg) * ld;
            const unsigned dst = st + (unsigned)((cg * SW_LD + 2 * cr) * 8);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int sz = cg + 8 * j < ncols ? szr : 0;
                cp_async16_s(dst + (unsigned)(8 * j * SW_LD * 8), sz ? src + (int64_t)(8 * j) * ld : F, sz);
            }
            if (nval > 64 && tid < 64) {   // 33rd chunk of every column (odd first row)
                const int sz = tid < ncols ? (65 < nval ? 16 : 8) : 0;
                cp_async16_s(st + (unsigned)((tid * SW_LD + 64) * 8), sz ? F + Ra + 64 + (int64_t)(C0 + tid) * ld : F, sz);
            }
            mbar_cp_async_arrive(&full_bar[stg]);
        };
        // B fragments of tile t: entry (q, rhs) with q = 8 warp + 4 ks + lc, rhs = lr
        auto load_b = [&](int t, double* b) {
            int R0, nrows, C0, ncols;
            tile_geom(t, R0, nrows, C0, ncols);
#pragma unroll
            for (int ks = 0; ks < 2; ks++) {
                const int q = 8 * warp + 4 * ks + lc;
                double v = 0.0;
                if (FWD) {
                    if (lr < nr && q < ncols) v = __ldcg(&y[bf.first + C0 + q + (int64_t)lr * n]);
                } else if (lr < nr && q < nrows) {
                    v = RECT ? __ldcg(&w[R0 + q + (int64_t)lr * m]) : __ldcg(&y[bf.first + R0 + q + (int64_t)lr * n]);
                }
                b[ks] = v;
            }
        };
        // number of published blocks of this front (forward: y_0..; backward: x_{npiv-1}, ...)
        auto poll = [&]() -> int {
            unsigned long long v = 0;
            if (lane == 0) v = ld_acquire_u64(mycnt);
            v = __shfl_sync(0xffffffffu, v, 0);
            return (unsigned)(v >> 32) == epoch ? (int)(unsigned)(v & 0xffffffffu) : 0;
        };
        double acc[8][2];
#pragma unroll
        for (int g = 0; g < 8; g++) acc[g][0] = acc[g][1] = 0.0;
#pragma unroll
        for (int q = 0; q < SW_ST - 1; q++)
            if (q < ntiles) issue(q);
        int ready = RECT ? (1 << 30) : 0;   // tile t needs `ready > t`
        bool have_next = false;
        double bn[2] = {0.0, 0.0};
        for (int t = 0; t < ntiles; t++) {
            double b[2];
            if (have_next) {
                b[0] = bn[0];
                b[1] = bn[1];
            } else {
                if (ready <= t) {
                    ready = poll();
                    // the iteration cap turns a protocol bug into a trapped kernel instead of a hung GPU
                    for (unsigned spins = 0; ready <= t; spins++) {
                        __nanosleep(20);
                        ready = poll();
                        if (spins > (1u << 24)) __trap();
                    }
                }
                load_b(t, b);
            }
            // prefetch the fragments of the next tile when its block is already published
            have_next = false;
            if (t + 1 < ntiles) {
                if (ready <= t + 1) ready = poll();
                if (ready > t + 1) {
                    load_b(t + 1, bn);
                    have_next = true;
                }
            }
            const unsigned g = gbase + (unsigned)t;
            const int stg = g % SW_ST;
            mbar_wait(&full_bar[stg], (g / SW_ST) & 1);
            if (t + SW_ST - 1 < ntiles) issue(t + SW_ST - 1);
            int R0, nrows, C0, ncols;
            tile_geom(t, R0, nrows, C0, ncols);
            const double* st = ring + stg * SW_TILE + (R0 & 1);
#pragma unroll
            for (int ks = 0; ks < 2; ks++) {
                const int kq = 8 * warp + 4 * ks + lc;   // contraction index inside the tile
                double af[8];
#pragma unroll
                for (int g2 = 0; g2 < 8; g2++)
                    af[g2] = FWD ? st[kq * SW_LD + 8 * g2 + lr] : st[(8 * g2 + lr) * SW_LD + kq];
#pragma unroll
                for (int g2 = 0; g2 < 8; g2++) dmma884(acc[g2][0], acc[g2][1], af[g2], b[ks]);
            }
            mbar_arrive(&empty_bar[stg]);
        }
        gbase += (unsigned)ntiles;
        __syncthreads();   // every warp has finished reading the ring
        // ---- add the eight partial products (staged in the idle ring)
        double* myred = ring + warp * 512;
#pragma unroll
        for (int g = 0; g < 8; g++) {
            myred[(8 * g + lr) * 8 + 2 * lc] = acc[g][0];
            myred[(8 * g + lr) * 8 + 2 * lc + 1] = acc[g][1];
        }
        __syncthreads();
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int e = tid + 256 * h;
            double sum = 0.0;
#pragma unroll
            for (int q = 0; q < 8; q++) sum += ring[q * 512 + e];
            if (RECT) {
                part[(bf.part_off + item.idx) * 512 + e] = sum;
            } else {
                const double v = bv[e] - sum;
                if (pivot) bv[e] = v;
                else {
                    const int a = e >> 3, c = e & 7;
                    if (a < nrow && c < nr) w[r0 + a + (int64_t)c * m] = v;
                }
            }
        }
        __syncthreads();
        if (RECT || !pivot) continue;
        // ---- out = D_ii^-1 v : warp computes entries 8 warp .. 8 warp + 7
        double c4[4][2];
#pragma unroll
        for (int q = 0; q < 4; q++) c4[q][0] = c4[q][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < 16; ks++) dmma884(c4[ks & 3][0], c4[ks & 3][1], ainv[ks], bv[(4 * ks + lc) * 8 + lr]);
        const double o0 = (c4[0][0] + c4[1][0]) + (c4[2][0] + c4[3][0]);
        const double o1 = (c4[0][1] + c4[1][1]) + (c4[2][1] + c4[3][1]);
        const int p = 8 * warp + lr;
        if (p < nrow) {
            if (2 * lc < nr) y[bf.first + r0 + p + (int64_t)(2 * lc) * n] = o0;
            if (2 * lc + 1 < nr) y[bf.first + r0 + p + (int64_t)(2 * lc + 1) * n] = o1;
        }
        // publish: the CTA barrier orders every thread's stores before thread 0's fence + release store
        // (the grid-sync pattern), so the other 255 threads do not each pay a device-wide fence
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            st_release_u64(cnt + item.front, ((unsigned long long)epoch << 32) | (unsigned)(FWD ? i + 1 : bf.npiv - i));
        }
    }
}

// In-place inversion of the 64 x 64 diagonal blocks, after the factorisation of their level:
// strictly lower part of the Larr block <- strictly lower part of L_kk^-1 (unit diagonal implied),
// lower part of the UTarr block <- (U_kk^-1)^T = (U_kk^T)^-1.  Both are inverses of a LOWER
// triangular matrix M stored in place (L_kk with unit diagonal; U_kk^T as stored), so one code path
// serves both: grid (blocks, 2), y = 0 inverts L_kk, y = 1 U_kk^T.
// Blocked: the eight 8 x 8 diagonal blocks are inverted by substitution, then the block size doubles
// three times with  inv([[A, 0], [B, C]]) = [[A^-1, 0], [-C^-1 B A^-1, C^-1]]  (two small products per
// off-diagonal block, all entries of a step in parallel): short dependency chains, small code.
constexpr int INV_THREADS = 128;

__global__ void __launch_bounds__(INV_THREADS) k_invert_diag(const DiagTask* __restrict__ tasks, double* __restrict__ heap) {
    __shared__ double T[64 * 65];    // T[a + 65 b] = M(a,b), a >= b; inverted in place, block by block
    __shared__ double Wk[32 * 32];   // B * A^-1 of the current step (all pairs)
    const DiagTask tk = tasks[blockIdx.x];
    const int kb = tk.kb, ld = tk.ld;
    const bool lower = blockIdx.y == 0;
    double* G = heap + (lower ? tk.Doff : tk.UTDoff);
    const int tid = threadIdx.x;
    for (int e = tid; e < 64 * 64; e += INV_THREADS) {
        const int a = e & 63, b = e >> 6;
        double v = 0.0;
        if (a == b) v = (lower || a >= kb) ? 1.0 : G[a + (int64_t)a * ld];
        else if (a > b && a < kb) v = G[a + (int64_t)b * ld];
        T[a + 65 * b] = v;
    }
    __syncthreads();
    {   // 8 x 8 diagonal blocks: thread = (block, column); in place via registers
        const int base = 8 * ((tid & 63) >> 3), c = tid & 7;
        double x[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            double acc = i == c ? 1.0 : 0.0;
#pragma unroll
            for (int p = 0; p < i; p++) acc -= T[base + i + 65 * (base + p)] * x[p];
            x[i] = acc / T[base + i + 65 * (base + i)];
        }
        __syncthreads();
        if (tid < 64)
#pragma unroll
            for (int i = 0; i < 8; i++) T[base + i + 65 * (base + c)] = x[i];
    }
    __syncthreads();
    for (int h = 8; h < 64; h *= 2) {
        const int nent = 32 * h;   // (32 / h) pairs of h x h entries
        // W = B * A^-1,  B = M[R0.., C0..] (still original),  A^-1 = T[C0.., C0..] (lower triangular)
        for (int e = tid; e < nent; e += INV_THREADS) {
            const int r = e % h, c = (e / h) % h, q = e / (h * h);
            const int C0 = 2 * q * h, R0 = C0 + h;
            double a0 = 0.0, a1 = 0.0;
            for (int p = c; p < h; p++) {
                const double t = T[R0 + r + 65 * (C0 + p)] * T[C0 + p + 65 * (C0 + c)];
                if (p & 1) a1 += t;
                else a0 += t;
            }
            Wk[e] = a0 + a1;
        }
        __syncthreads();
        // the B block receives -C^-1 * W,  C^-1 = T[R0.., R0..] (lower triangular)
        for (int e = tid; e < nent; e += INV_THREADS) {
            const int r = e % h, c = (e / h) % h, q = e / (h * h);
            const int C0 = 2 * q * h, R0 = C0 + h;
            const double* wq = Wk + q * h * h + h * c;
            double a0 = 0.0, a1 = 0.0;
            for (int p = 0; p <= r; p++) {
                const double t = T[R0 + r + 65 * (R0 + p)] * wq[p];
                if (p & 1) a1 += t;
                else a0 += t;
            }
            T[R0 + r + 65 * (C0 + c)] = -(a0 + a1);
        }
        __syncthreads();
    }
    for (int e = tid; e < 64 * 64; e += INV_THREADS) {
        const int a = e & 63, b = e >> 6;
        if (a < kb && b < kb && (lower ? a > b : a >= b)) G[a + (int64_t)b * ld] = T[a + 65 * b];
    }
}

// ------------------------------------------------------------------------------------------
// residual and componentwise backward error (pdgsrfs: pdgsmv_AXglobal, pdgsmv_AXglobal_abs)
//   r = b - A x ;  berr_c = max_i |r_i| / (|A||x| + |b|)_i
// ------------------------------------------------------------------------------------------

// One thread per row, ALL right-hand sides in one pass over A (12 nnz + 4 (n+1) + 24 n nrhs bytes,
// BASELINE.md section 4): a and its column index are loaded once per nonzero and used for NR columns.
// x is column-major; consecutive rows are consecutive cells of a water column (src/matrix.c:239-251), so
// the gathers x[col + c ldx] of a warp are contiguous runs per neighbour direction and per column.
// berr follows pdgsrfs: rows with (|A||x| + |b|)_i > safe2 contribute |r_i| / den_i, rows with a non-zero
// smaller denominator (|r_i| + safe1) / (den_i + safe1), rows with a zero denominator nothing.
// EXTRA: the residual is accumulated in twice the working precision (error-free products by FMA, error-free
// sums, Ogita-Rump-Oishi Dot2) and rounded once -- LAPACK's "extra-precise iterative refinement" (xGERFSX);
// the refined solution then no longer carries cond(A) times the rounding errors of a working-precision residual.
template <int NR, bool EXTRA>
__global__ void __launch_bounds__(256) k_residual(int n, int nrhs, const int* __restrict__ rowptr,
                                                  const int* __restrict__ colind, const double* __restrict__ val,
                                                  const double* __restrict__ x, int ldx, const double* __restrict__ b,
                                                  int ldb, double* __restrict__ r, double* __restrict__ berr,
                                                  double safe1, double safe2) {
    __shared__ double wmax[8][NR];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double e[NR];
#pragma unroll
    for (int c = 0; c < NR; c++) e[c] = 0.0;
    if (i < n) {
        double acc[NR], aabs[NR], lo[NR];
#pragma unroll
        for (int c = 0; c < NR; c++) {
            const double bi = c < nrhs ? b[i + (int64_t)c * ldb] : 0.0;
            acc[c] = bi;
            lo[c] = 0.0;
            aabs[c] = fabs(bi);
        }
        const int p1 = rowptr[i + 1];
        for (int p = rowptr[i]; p < p1; p++) {
            const double a = val[p], aa = fabs(a);
            const double* xp = x + colind[p];
#pragma unroll
            for (int c = 0; c < NR; c++) {
                const double xv = c < nrhs ? xp[(int64_t)c * ldx] : 0.0;
                if (EXTRA) {
                    const double pr = a * xv, pe = fma(a, xv, -pr);        // a * xv == pr + pe exactly
                    const double sm = acc[c] - pr, bb = sm - acc[c];
                    const double se = (acc[c] - (sm - bb)) + (-pr - bb);   // acc - pr == sm + se exactly
                    acc[c] = sm;
                    lo[c] += se - pe;
                } else {
                    acc[c] = fma(-a, xv, acc[c]);
                }
                aabs[c] = fma(aa, fabs(xv), aabs[c]);
            }
        }
        if (EXTRA)
#pragma unroll
            for (int c = 0; c < NR; c++) acc[c] += lo[c];
#pragma unroll
        for (int c = 0; c < NR; c++) {
            if (c < nrhs) {
                r[i + (int64_t)c * n] = acc[c];
                if (aabs[c] > safe2) e[c] = fabs(acc[c]) / aabs[c];
                else if (aabs[c] != 0.0) e[c] = (fabs(acc[c]) + safe1) / (aabs[c] + safe1);
                if (e[c] != e[c]) e[c] = 1e300;   // a NaN residual must not pass for "converged" (fmax drops NaNs)
            }
        }
    }
    if (berr == nullptr) return;
#pragma unroll
    for (int c = 0; c < NR; c++) {
        double m = e[c];
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5][c] = m;
    }
    __syncthreads();
    if (threadIdx.x < NR && threadIdx.x < nrhs) {
        double m = wmax[0][threadIdx.x];
        for (int w = 1; w < 8; w++) m = fmax(m, wmax[w][threadIdx.x]);
        atomic_max_pos_double(&berr[threadIdx.x], m);   // max is order-independent: deterministic
    }
}

// sum of squares per column (normwise stopping rule), deterministic: every block writes its partial
// sum to part[block * nrhs + c]; k_sumsq_final adds the partials in block order.
constexpr int SUMSQ_BLOCKS = 256;
__global__ void __launch_bounds__(256) k_sumsq(int n, int nrhs, const double* __restrict__ v, int ld, double* __restrict__ part) {
    __shared__ double red[256];
    for (int c = 0; c < nrhs; c++) {
        double acc = 0;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            double t = v[i + (int64_t)c * ld];
            acc += t * t;
        }
        red[threadIdx.x] = acc;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) part[blockIdx.x * nrhs + c] = red[0];
        __syncthreads();
    }
}
__global__ void k_sumsq_final(int nblocks, int nrhs, const double* __restrict__ part, double* __restrict__ out) {
    const int c = threadIdx.x;
    if (c >= nrhs) return;
    double acc = 0;
    for (int b = 0; b < nblocks; b++) acc += part[b * nrhs + c];
    out[c] = acc;
}

// multi-GPU: parts of the solution y (n x nr, column-major) <-> contiguous per-range blocks of `pack`
// (range [lo, hi) occupies pack[lo * nr, hi * nr), column c at offset c * (hi - lo)), so that every range
// travels as ONE broadcast.  dir 0 packs the ranges this rank publishes, dir 1 unpacks the others.
__global__ void __launch_bounds__(256) k_pub_pack(const PubRange* __restrict__ ranges, int me, int dir, double* __restrict__ y,
                                                  int n, int nr, double* __restrict__ pack) {
    const PubRange pr = ranges[blockIdx.y];
    if ((dir == 0) != (pr.root == me)) return;
    const int len = pr.hi - pr.lo;
    double* pk = pack + (int64_t)pr.lo * nr;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x)
        for (int c = 0; c < nr; c++) {
            if (dir == 0) pk[(int64_t)c * len + i] = y[pr.lo + i + (int64_t)c * n];
            else y[pr.lo + i + (int64_t)c * n] = pk[(int64_t)c * len + i];
        }
}

// diagnostics (NKP_CHECK=1): number of non-finite entries
__global__ void __launch_bounds__(256) k_count_nonfinite(const double* __restrict__ p, int64_t n, unsigned long long* __restrict__ out) {
    unsigned long long c = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) c += !isfinite(p[i]);
    if (c) atomicAdd(out, c);
}

// max |a| over the stored values (tiny-pivot threshold of an unequilibrated factorisation)
__global__ void __launch_bounds__(256) k_absmax(int64_t nnz, const double* __restrict__ val, double* __restrict__ out) {
    double m = 0.0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x)
        m = fmax(m, fabs(val[p]));
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomic_max_pos_double(out, m);
}

}  // namespace nkp
