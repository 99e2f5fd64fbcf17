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

// ------------------------------------------------------------------------------------------------------------
// Stencil values on the device for the option set  adv_type centered / hmix_type const / vmix_type const /
// sink_type const_shallow  (SURVEY.md appendix B "minimal input"): what gen_sparse_matrix computes between init_matrix and
// sum_dup_vals (src/matrix.c:3790-3827) for that option set -- add_UTE_coeffs :1239-1273, add_VTN_coeffs :1320-1360,
// add_WVEL_coeffs :1401-1430, adv_enforce_divfree :2094-2206, add_hmix_const :2656-2710, add_vmix_const :2978-3004,
// add_sink_pure_diag :3084-3091 -- one thread per tracer-state entry, every floating-point operation in the reference's
// order with explicit round-to-nearest intrinsics (no FMA contraction), so that the CRS after nkp_crs_finalize_device is
// bit-identical to the file gen_A writes.  Rows come out in slot order (self, k-1, k+1, east, west, north, south;
// absent neighbours skipped), exact zeros kept.  Two launches: slot counts (-> exclusive scan -> rowptr), then the fill.
// With this, a Newton sequence on one sparsity pattern never leaves the GPU: new circulation fields -> values in
// the same slots -> nkp_crs_finalize_device(strip_zeros = 0) -> nkp_factor_device.
// ------------------------------------------------------------------------------------------------------------

namespace {

struct MinFields {
    int imt, jmt, km, n;
    const int *KMT, *ind_i, *ind_j, *ind_k, *int3;
    const double *dz, *z_t, *TAREA, *HTE, *HUS, *HTN, *HUW, *DXU, *DYU, *UVEL, *VVEL, *WVEL;
    double delta_t, year_cnt, sink_rate, sink_depth, fill;
};

__device__ __forceinline__ double fv0(double a, double fill) { return a == fill ? 0.0 : a; }
__device__ __forceinline__ int kmu_of(const MinFields& f, int j, int i) {   // src/grid.c:187-203
    if (j >= f.jmt - 1) return 0;
    const int ip1 = i < f.imt - 1 ? i + 1 : 0;
    const int a = f.KMT[j * f.imt + i], b = f.KMT[(j + 1) * f.imt + i];
    const int c = f.KMT[j * f.imt + ip1], d = f.KMT[(j + 1) * f.imt + ip1];
    return min(min(a, b), min(c, d));
}
// 0.5 * vel * len on U points below the sea floor of the U cell: 0
__device__ __forceinline__ double half_transport(const MinFields& f, const double* vel, const double* len, int k, int j, int i) {
    if (k >= kmu_of(f, j, i)) return 0.0;
    const double v = fv0(vel[((int64_t)k * f.jmt + j) * f.imt + i], f.fill);
    return __dmul_rn(__dmul_rn(0.5, v), len[j * f.imt + i]);
}
__device__ __forceinline__ double ute_at(const MinFields& f, int k, int j, int i) {   // load_UTE, src/matrix.c:1024-1031
    if (j < 1 || j > f.jmt - 2) return 0.0;
    return __dadd_rn(__dadd_rn(0.0, half_transport(f, f.UVEL, f.DYU, k, j, i)), half_transport(f, f.UVEL, f.DYU, k, j - 1, i));
}
__device__ __forceinline__ double vtn_at(const MinFields& f, int k, int j, int i) {   // load_VTN, src/matrix.c:1103-1111
    if (j < 1 || j > f.jmt - 2) return 0.0;
    const int im1 = i > 0 ? i - 1 : f.imt - 1;
    return __dadd_rn(__dadd_rn(0.0, half_transport(f, f.VVEL, f.DXU, k, j, i)), half_transport(f, f.VVEL, f.DXU, k, j, im1));
}
__device__ __forceinline__ double w_at(const MinFields& f, int k, int j, int i) {   // load_WVEL, src/matrix.c:1168-1198
    if (k >= f.km || k >= f.KMT[j * f.imt + i] || j == 0 || j == f.jmt - 1 || k == 0) return 0.0;
    return __dadd_rn(0.0, fv0(f.WVEL[((int64_t)k * f.jmt + j) * f.imt + i], f.fill));
}

// MODE 0: number of slots of every row -> cnt;  MODE 1: fill colind / val at rowptr
template <int MODE>
__global__ void __launch_bounds__(256) k_assemble_min(MinFields f, int* __restrict__ cnt, const int* __restrict__ rowptr,
                                                      int* __restrict__ colind, double* __restrict__ val) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= f.n) return;
    const int i = f.ind_i[q], j = f.ind_j[q], k = f.ind_k[q];
    const int imt = f.imt;
    const int ip1 = i < imt - 1 ? i + 1 : 0, im1 = i > 0 ? i - 1 : imt - 1;
    const bool has_up = k - 1 >= 0, has_dn = k + 1 < f.KMT[j * imt + i];
    const bool has_e = k < f.KMT[j * imt + ip1], has_w = k < f.KMT[j * imt + im1];
    const bool has_n = k < f.KMT[(j + 1) * imt + i], has_s = k < f.KMT[(j - 1) * imt + i];
    if (MODE == 0) {
        cnt[q] = 1 + has_up + has_dn + has_e + has_w + has_n + has_s;
        return;
    }
    const double dt = f.delta_t, ta = f.TAREA[j * imt + i], dzk = f.dz[k];
    auto adv = [&](double flux, double den, bool plus) {   // 0.0 -/+ (1 - w) * flux / den * delta_t, w = 0.5
        const double t = __dmul_rn(__ddiv_rn(__dmul_rn(0.5, flux), den), dt);
        return plus ? __dadd_rn(0.0, t) : __dsub_rn(0.0, t);
    };
    const double a_e = adv(ute_at(f, k, j, i), ta, false), a_w = adv(ute_at(f, k, j, im1), ta, true);
    const double a_n = adv(vtn_at(f, k, j, i), ta, false), a_s = adv(vtn_at(f, k, j - 1, i), ta, true);
    const double a_up = adv(w_at(f, k, j, i), dzk, false), a_dn = adv(w_at(f, k + 1, j, i), dzk, true);
    // adv_enforce_divfree: self = -(sum of the present non-self slots in slot order)
    double ssum = 0.0;
    if (has_up) ssum = __dadd_rn(ssum, a_up);
    if (has_dn) ssum = __dadd_rn(ssum, a_dn);
    if (has_e) ssum = __dadd_rn(ssum, a_e);
    if (has_w) ssum = __dadd_rn(ssum, a_w);
    if (has_n) ssum = __dadd_rn(ssum, a_n);
    if (has_s) ssum = __dadd_rn(ssum, a_s);
    double v_self = -ssum;
    // hmix const
    const double ah = 4.0e6;
    auto hm = [&](const double* num, const double* den, int jj, int ii) {
        const double a = fv0(num[jj * imt + ii], f.fill), b = fv0(den[jj * imt + ii], f.fill);
        return __dmul_rn(__ddiv_rn(__ddiv_rn(__dmul_rn(ah, a), b), ta), dt);
    };
    const double ce = has_e ? hm(f.HTE, f.HUS, j, i) : 0.0, cw = has_w ? hm(f.HTE, f.HUS, j, im1) : 0.0;
    const double cn = has_n ? hm(f.HTN, f.HUW, j, i) : 0.0, cs = has_s ? hm(f.HTN, f.HUW, j - 1, i) : 0.0;
    v_self = __dsub_rn(v_self, __dadd_rn(__dadd_rn(__dadd_rn(ce, cw), cn), cs));
    // vmix const
    const double vdc = 0.1;
    const double dz_up = f.dz[max(k - 1, 0)], dz_dn = f.dz[min(k + 1, f.km - 1)];
    const double ct = has_up ? __dmul_rn(__ddiv_rn(__ddiv_rn(vdc, __dmul_rn(0.5, __dadd_rn(dz_up, dzk))), dzk), dt) : 0.0;
    const double cb = has_dn ? __dmul_rn(__ddiv_rn(__ddiv_rn(vdc, __dmul_rn(0.5, __dadd_rn(dzk, dz_dn))), dzk), dt) : 0.0;
    v_self = __dsub_rn(v_self, __dadd_rn(ct, cb));
    // sink const_shallow
    if (f.z_t[k] < f.sink_depth) v_self = __dadd_rn(v_self, -__dmul_rn(f.year_cnt, f.sink_rate));
    int p = rowptr[q];
    auto put = [&](int col, double v) {
        colind[p] = col;
        val[p] = v;
        p++;
    };
    auto nbr = [&](int jj, int ii) { return f.int3[((int64_t)k * f.jmt + jj) * imt + ii]; };
    put(q, v_self);
    if (has_up) put(q - 1, __dadd_rn(a_up, ct));
    if (has_dn) put(q + 1, __dadd_rn(a_dn, cb));
    if (has_e) put(nbr(j, ip1), __dadd_rn(a_e, ce));
    if (has_w) put(nbr(j, im1), __dadd_rn(a_w, cw));
    if (has_n) put(nbr(j + 1, i), __dadd_rn(a_n, cn));
    if (has_s) put(nbr(j - 1, i), __dadd_rn(a_s, cs));
}

}  // namespace

extern "C" int nkp_assemble_min_device(const nkp_min_fields* fd, double day_cnt, double sink_rate, double sink_depth,
                                       int* d_rowptr, int* d_colind, double* d_val, long long capacity, long long* nnz_out) {
    if (!fd || !d_rowptr || !d_colind || !d_val || fd->n <= 0 || fd->imt < 3 || fd->jmt < 3 || fd->km < 1) return NKP_EINVAL;
    MinFields f;
    f.imt = fd->imt; f.jmt = fd->jmt; f.km = fd->km; f.n = fd->n;
    f.KMT = fd->KMT; f.ind_i = fd->ind_i; f.ind_j = fd->ind_j; f.ind_k = fd->ind_k; f.int3 = fd->int3_to_tracer_state_ind;
    f.dz = fd->dz; f.z_t = fd->z_t; f.TAREA = fd->TAREA; f.HTE = fd->HTE; f.HUS = fd->HUS; f.HTN = fd->HTN; f.HUW = fd->HUW;
    f.DXU = fd->DXU; f.DYU = fd->DYU; f.UVEL = fd->UVEL; f.VVEL = fd->VVEL; f.WVEL = fd->WVEL;
    f.delta_t = 60.0 * 60.0 * 24.0 * day_cnt;
    f.year_cnt = day_cnt / 365.0;
    f.sink_rate = sink_rate;
    f.sink_depth = sink_depth;
    f.fill = fd->fill_value;
    int rc = NKP_OK;
    const int n = fd->n, g = (n + 255) / 256, nblocks = (n + SCAN_ITEMS - 1) / SCAN_ITEMS;
    int *d_cnt = nullptr, *d_bsum = nullptr, *d_tot = nullptr;
    int total = 0;
    CKC(cudaMalloc((void**)&d_cnt, sizeof(int) * (size_t)n));
    CKC(cudaMalloc((void**)&d_bsum, sizeof(int) * (size_t)nblocks));
    CKC(cudaMalloc((void**)&d_tot, sizeof(int)));
    k_assemble_min<0><<<g, 256>>>(f, d_cnt, nullptr, nullptr, nullptr);
    k_scan_blocks<<<nblocks, SCAN_THREADS>>>(n, d_cnt, d_bsum);
    k_scan_block_sums<<<1, SCAN_THREADS>>>(nblocks, d_bsum, d_tot);
    k_scan_finish<<<(n + 1 + 255) / 256, 256>>>(n, d_cnt, d_bsum, d_tot, d_rowptr);
    CKC(cudaMemcpy(&total, d_tot, sizeof(int), cudaMemcpyDeviceToHost));
    if (nnz_out) *nnz_out = total;
    if ((long long)total > capacity) {
        g_crs_err = "nkp_assemble_min_device: colind / val capacity too small";
        rc = NKP_ENOMEM;
        goto done;
    }
    k_assemble_min<1><<<g, 256>>>(f, nullptr, d_rowptr, d_colind, d_val);
    CKC(cudaGetLastError());
    CKC(cudaDeviceSynchronize());
done:
    cudaFree(d_cnt);
    cudaFree(d_bsum);
    cudaFree(d_tot);
    return rc;
}

extern "C" int nkp_bswap32_device(void* d_data, long long count) {
    if (!d_data || count < 0) return NKP_EINVAL;
    if (count == 0) return NKP_OK;
    k_bswap32<<<(unsigned)((count + 255) / 256), 256>>>(static_cast<unsigned*>(d_data), (int64_t)count);
    return cudaGetLastError() == cudaSuccess ? NKP_OK : NKP_ECUDA;
}

extern "C" const char* nkp_crs_last_error(void) { return g_crs_err.c_str(); }
