// crs_device.cu -- the matrix post-processing of the reference's generator on the device.
//
// gen_sparse_matrix (src/matrix.c:3775-3840) finishes every assembly with
//     sum_dup_vals ()         src/matrix.c:3621-3650   entries of a row that name the same column are summed into the first
//     strip_matrix_zeros ()   src/matrix.c:3657-3688   exact zeros are removed, rowptr is rebuilt
//     sort_cols_all_rows ()   src/matrix.c:3753-3765   columns ascending within every row (insertion sort, stable)
// on the host.  For a Newton sequence that stays on the GPU (new circulation -> new values in the SAME slots ->
// nkp_factor_device) the same three steps are needed on a CRS held in device memory.  Rows are independent and at most
// 21 entries long (src/matrix.c:621-650), so one thread owns one row and walks it exactly as the reference's loops do:
// the floating-point additions of sum_dup_vals happen in the reference's order (bit-exact), the insertion sort is the
// same stable sort.  strip_matrix_zeros makes the PATTERN value-dependent; a same-pattern refactorisation must keep the
// explicit zeros (strip_zeros = 0), which is why it is optional here.
//
// HBM-bound: 12 B read + 12 B written per entry and step; the exclusive scan of the row counts is a three-launch
// block scan (n <= 7.4 M rows).
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/nkprecond.h"

namespace {

thread_local std::string g_crs_err;

__global__ void __launch_bounds__(256) k_crs_sum_dup(int n, const int* __restrict__ rowptr, const int* __restrict__ colind,
                                                     double* __restrict__ val, int* __restrict__ dup_cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p0 = rowptr[i], p1 = rowptr[i + 1];
    int cnt = 0;
    for (int p = p0; p < p1; p++)
        for (int q = p + 1; q < p1; q++)
            if (colind[q] == colind[p]) {
                val[p] = __dadd_rn(val[p], val[q]);   // the reference's order of additions
                val[q] = 0.0;
                cnt++;
            }
    if (cnt) atomicAdd(dup_cnt, cnt);
}

__global__ void __launch_bounds__(256) k_crs_count_nonzero(int n, const int* __restrict__ rowptr, const double* __restrict__ val,
                                                           int* __restrict__ cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = 0;
    for (int p = rowptr[i]; p < rowptr[i + 1]; p++) c += val[p] != 0.0;
    cnt[i] = c;
}

// exclusive scan, step 1: every block scans SCAN_ITEMS consecutive counts in place and stores its total
constexpr int SCAN_THREADS = 256, SCAN_PER = 4, SCAN_ITEMS = SCAN_THREADS * SCAN_PER;
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_blocks(int n, int* __restrict__ data, int* __restrict__ block_sum) {
    __shared__ int sh[SCAN_THREADS];
    const int base = blockIdx.x * SCAN_ITEMS + threadIdx.x * SCAN_PER;
    int v[SCAN_PER], tot = 0;
#pragma unroll
    for (int q = 0; q < SCAN_PER; q++) {
        v[q] = base + q < n ? data[base + q] : 0;
        tot += v[q];
    }
    sh[threadIdx.x] = tot;
    __syncthreads();
    for (int o = 1; o < SCAN_THREADS; o <<= 1) {
        int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    int run = sh[threadIdx.x] - tot;   // exclusive prefix of this thread inside the block
#pragma unroll
    for (int q = 0; q < SCAN_PER; q++) {
        if (base + q < n) data[base + q] = run;
        run += v[q];
    }
    if (threadIdx.x == SCAN_THREADS - 1) block_sum[blockIdx.x] = sh[threadIdx.x];
}

// step 2: one block turns the block totals into exclusive offsets (sequential over chunks of 256)
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_block_sums(int nblocks, int* __restrict__ block_sum, int* __restrict__ total) {
    __shared__ int sh[SCAN_THREADS];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nblocks; b0 += SCAN_THREADS) {
        const int idx = b0 + threadIdx.x;
        const int v = idx < nblocks ? block_sum[idx] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < SCAN_THREADS; o <<= 1) {
            int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (idx < nblocks) block_sum[idx] = carry + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == SCAN_THREADS - 1) carry += sh[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// step 3: add the block offsets; out has n + 1 entries, out[n] = total
__global__ void __launch_bounds__(256) k_scan_finish(int n, const int* __restrict__ scanned, const int* __restrict__ block_sum,
                                                     const int* __restrict__ total, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = scanned[i] + block_sum[i / SCAN_ITEMS];
    if (i == n) out[n] = *total;
}

__global__ void __launch_bounds__(256) k_crs_compact(int n, const int* __restrict__ rowptr_old, const int* __restrict__ rowptr_new,
                                                     const int* __restrict__ colind_in, const double* __restrict__ val_in,
                                                     int* __restrict__ colind_out, double* __restrict__ val_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int o = rowptr_new[i];
    for (int p = rowptr_old[i]; p < rowptr_old[i + 1]; p++)
        if (val_in[p] != 0.0) {
            val_out[o] = val_in[p];
            colind_out[o] = colind_in[p];
            o++;
        }
}

// sort_cols_one_row (src/matrix.c:3730-3747): insertion sort, strict comparison => stable
__global__ void __launch_bounds__(256) k_crs_sort_rows(int n, const int* __restrict__ rowptr, int* __restrict__ colind,
                                                       double* __restrict__ val) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int* c = colind + rowptr[r];
    double* v = val + rowptr[r];
    const int len = rowptr[r + 1] - rowptr[r];
    for (int i = 1; i < len; i++) {
        const int key = c[i];
        const double kv = v[i];
        int j = i - 1;
        for (; j >= 0 && c[j] > key; --j) {
            c[j + 1] = c[j];
            v[j + 1] = v[j];
        }
        c[j + 1] = key;
        v[j + 1] = kv;
    }
}

// big-endian 32-bit integers of the matrix file (NC_INT colind / rowptr, src/matrix.c:3884,3888) -> host order, in place
__global__ void __launch_bounds__(256) k_bswap32(unsigned* __restrict__ a, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = __byte_perm(a[i], 0, 0x0123);
}

#define CKC(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            g_crs_err = std::string(#call) + " failed: " + cudaGetErrorString(e_);                 \
            rc = NKP_ECUDA;                                                                        \
            goto done;                                                                             \
        }                                                                                          \
    } while (0)

}  // namespace

extern "C" int nkp_crs_finalize_device(int n, int* d_rowptr, int* d_colind, double* d_val, int strip_zeros,
                                       long long* nnz_out, int* dup_cnt_out) {
    if (n <= 0 || !d_rowptr || !d_colind || !d_val) return NKP_EINVAL;
    int rc = NKP_OK;
    int *d_cnt = nullptr, *d_bsum = nullptr, *d_tot = nullptr, *d_rp_new = nullptr, *d_ci_new = nullptr;
    double* d_v_new = nullptr;
    int nnz_old = 0, nnz_new = 0, dup = 0;
    const int g = (n + 255) / 256;
    const int nblocks = (n + SCAN_ITEMS - 1) / SCAN_ITEMS;
    CKC(cudaMemcpy(&nnz_old, d_rowptr + n, sizeof(int), cudaMemcpyDeviceToHost));
    CKC(cudaMalloc((void**)&d_tot, 2 * sizeof(int)));
    CKC(cudaMemset(d_tot, 0, 2 * sizeof(int)));
    k_crs_sum_dup<<<g, 256>>>(n, d_rowptr, d_colind, d_val, d_tot + 1);
    nnz_new = nnz_old;
    if (strip_zeros) {
        CKC(cudaMalloc((void**)&d_cnt, sizeof(int) * (size_t)n));
        CKC(cudaMalloc((void**)&d_bsum, sizeof(int) * (size_t)nblocks));
        CKC(cudaMalloc((void**)&d_rp_new, sizeof(int) * ((size_t)n + 1)));
        k_crs_count_nonzero<<<g, 256>>>(n, d_rowptr, d_val, d_cnt);
        k_scan_blocks<<<nblocks, SCAN_THREADS>>>(n, d_cnt, d_bsum);
        k_scan_block_sums<<<1, SCAN_THREADS>>>(nblocks, d_bsum, d_tot);
        k_scan_finish<<<(n + 1 + 255) / 256, 256>>>(n, d_cnt, d_bsum, d_tot, d_rp_new);
        CKC(cudaMemcpy(&nnz_new, d_tot, sizeof(int), cudaMemcpyDeviceToHost));
        CKC(cudaMalloc((void**)&d_ci_new, sizeof(int) * (size_t)(nnz_new > 0 ? nnz_new : 1)));
        CKC(cudaMalloc((void**)&d_v_new, sizeof(double) * (size_t)(nnz_new > 0 ? nnz_new : 1)));
        k_crs_compact<<<g, 256>>>(n, d_rowptr, d_rp_new, d_colind, d_val, d_ci_new, d_v_new);
        CKC(cudaMemcpy(d_colind, d_ci_new, sizeof(int) * (size_t)nnz_new, cudaMemcpyDeviceToDevice));
        CKC(cudaMemcpy(d_val, d_v_new, sizeof(double) * (size_t)nnz_new, cudaMemcpyDeviceToDevice));
        CKC(cudaMemcpy(d_rowptr, d_rp_new, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToDevice));
    }
    k_crs_sort_rows<<<g, 256>>>(n, d_rowptr, d_colind, d_val);
    CKC(cudaGetLastError());
    CKC(cudaMemcpy(&dup, d_tot + 1, sizeof(int), cudaMemcpyDeviceToHost));
    CKC(cudaDeviceSynchronize());
    if (nnz_out) *nnz_out = nnz_new;
    if (dup_cnt_out) *dup_cnt_out = dup;
done:
    cudaFree(d_cnt);
    cudaFree(d_bsum);
    cudaFree(d_tot);
    cudaFree(d_rp_new);
    cudaFree(d_ci_new);
    cudaFree(d_v_new);
    return rc;
}

extern "C" int nkp_bswap32_device(void* d_data, long long count) {
    if (!d_data || count < 0) return NKP_EINVAL;
    if (count == 0) return NKP_OK;
    k_bswap32<<<(unsigned)((count + 255) / 256), 256>>>(static_cast<unsigned*>(d_data), (int64_t)count);
    return cudaGetLastError() == cudaSuccess ? NKP_OK : NKP_ECUDA;
}

extern "C" const char* nkp_crs_last_error(void) { return g_crs_err.c_str(); }
