// solver.cu -- device-side orchestration and the C ABI (include/nkprecond.h).
//
// One nkp_solver owns: the analysis Plan (host), its task lists mirrored in device memory,
// the device heap (factors + update-matrix pools), the CRS operand on the device and the
// solve work space.  nkp_factor replays the static launch sequence of the plan on one
// stream; nkp_solve runs permute/scale -> forward sweep -> backward sweep -> refinement
// (residual SpMV + correction solves, SuperLU's pdgsrfs stopping rule).
//
// There is no CPU fallback: every failure of a CUDA call is reported as NKP_ECUDA.
#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/nkprecond.h"
#include "kernels.cuh"
#include "nkp_internal.hpp"

using namespace nkp;

static thread_local std::string g_err;

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char b_[512];                                                                          \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            g_err = b_;                                                                            \
            return NKP_ECUDA;                                                                      \
        }                                                                                          \
    } while (0)

// NCCL is bound lazily with dlopen: single-GPU users never load it, and a process that already
// carries a (possibly newer) libnccl.so.2 -- e.g. the one bundled with PyTorch -- keeps using
// that copy instead of getting a second, conflicting one.
#include <dlfcn.h>
namespace {
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, ncclConfig_t*) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;
bool nccl_load() {
    if (g_nccl.ok) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return false;
    g_nccl.handle = h;
#define NKP_SYM(field, name) *(void**)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) return false;
    NKP_SYM(GetUniqueId, "ncclGetUniqueId")
    NKP_SYM(CommInitRank, "ncclCommInitRank")
    NKP_SYM(CommDestroy, "ncclCommDestroy")
    NKP_SYM(CommSplit, "ncclCommSplit")
    NKP_SYM(AllGather, "ncclAllGather")
    NKP_SYM(GroupStart, "ncclGroupStart")
    NKP_SYM(GroupEnd, "ncclGroupEnd")
    NKP_SYM(Send, "ncclSend")
    NKP_SYM(Recv, "ncclRecv")
    NKP_SYM(Broadcast, "ncclBroadcast")
    NKP_SYM(GetErrorString, "ncclGetErrorString")
#undef NKP_SYM
    g_nccl.ok = true;
    return true;
}
}  // namespace
#define ncclGetUniqueId g_nccl.GetUniqueId
#define ncclCommInitRank g_nccl.CommInitRank
#define ncclCommDestroy g_nccl.CommDestroy
#define ncclCommSplit g_nccl.CommSplit
#define ncclAllGather g_nccl.AllGather
#define ncclGroupStart g_nccl.GroupStart
#define ncclGroupEnd g_nccl.GroupEnd
#define ncclSend g_nccl.Send
#define ncclRecv g_nccl.Recv
#define ncclBroadcast g_nccl.Broadcast
#define ncclGetErrorString g_nccl.GetErrorString

#define CKN(call)                                                                                  \
    do {                                                                                           \
        ncclResult_t e_ = (call);                                                                  \
        if (e_ != ncclSuccess) {                                                                   \
            char b_[512];                                                                          \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, ncclGetErrorString(e_), __FILE__, __LINE__); \
            g_err = b_;                                                                            \
            return NKP_ECUDA;                                                                      \
        }                                                                                          \
    } while (0)

static const int MAX_NR = 8;

// ---- host staging ---------------------------------------------------------------------------------
// Caller buffers are plain C arrays (src/matrix.c:84, src/solve_ABglobal.c:344).  Page-locked ones
// (cudaMallocHost / cudaHostRegister by the caller) are handed to the copy engine directly; pageable
// ones go through the solver's pinned staging area with a multi-threaded copy -- one core moves about
// 10 GB/s, a PCIe 5 x16 link 50+.
#include <omp.h>
static bool host_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}
static void par_memcpy(void* dst, const void* src, size_t bytes) {
    const size_t chunk = (size_t)1 << 21;
    const long nchunks = (long)((bytes + chunk - 1) / chunk);
    const int nt = std::max(1, std::min(16, omp_get_max_threads()));
    if (nchunks <= 1 || nt == 1) {
        memcpy(dst, src, bytes);
        return;
    }
#pragma omp parallel for schedule(static) num_threads(nt)
    for (long c = 0; c < nchunks; c++) {
        const size_t o = (size_t)c * chunk;
        memcpy((char*)dst + o, (const char*)src + o, std::min(chunk, bytes - o));
    }
}

struct nkp_solver {
    Plan plan;
    nkp_options opt;
    int n = 0;
    int64_t nnz = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_col[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // per-column D2H completion
    // device data
    double* heap = nullptr;
    int* d_rowptr = nullptr;
    int* d_colind = nullptr;
    int* d_rowidx = nullptr;
    double* d_val = nullptr;
    int64_t* d_scatter = nullptr;
    int* d_perm = nullptr;
    int* d_permr = nullptr;     // row numbering of the permuted system: d_perm, or perm[rowmap[.]] with a static row permutation
    bool fixed_scale = false;   // d_R / d_C were supplied at creation (nkp_create_rowperm) and are reused by every factorisation
    int* d_bidx = nullptr;
    int* d_rel = nullptr;
    int* d_clo = nullptr;
    double* d_R = nullptr;
    double* d_C = nullptr;
    DiagTask* d_diag = nullptr;
    TrsmTask* d_trsm = nullptr;
    GemmTask* d_gemm = nullptr;
    AddTask* d_add = nullptr;
    SolveTask* d_solve = nullptr;
    SolveTask* d_small = nullptr;
    SolveChild* d_children = nullptr;
    BigFront* d_big = nullptr;
    BigItem* d_fwd_items = nullptr;
    BigItem* d_bwd_items = nullptr;
    BigItem* d_rect_items = nullptr;
    double* d_part = nullptr;   // 64 x 8 partial products of the backward sweep's rectangular part
    int epoch = 0;                        // sweep counter: tag in the upper half of the progress counters
    unsigned long long* d_cnt = nullptr;  // progress counters of the big fronts: [0, nbig) forward, [nbig, 2 nbig) backward
    DiagTask* d_inv = nullptr;            // diagonal blocks inverted after the factorisation
    int coop_ctas = 0;          // co-resident CTAs for the dataflow sweeps
    int num_sms = 0;
    int small_v1 = 0;           // NKP_SMALL_V1: bit 0 / bit 1 = first-generation warp-per-front kernel for the forward /
                                // backward sweep of the small fronts (A/B against k_fwd_front / k_bwd_front)
    bool small_force = false;   // NKP_SMALL_FORCE=1: k_*_front for every level of small fronts (tests)
    std::vector<int> small_wcap;   // per level: largest m of its small fronts (shared-memory work vector of k_*_front)
    ncclComm_t comm = nullptr;  // multi-GPU only
    int rank = 0, nranks = 1;
    std::vector<ncclComm_t> gcomm;   // per Plan::groups entry: communicator of the group (null: this rank is not in it, or size 1)
    cudaStream_t cstream = nullptr;  // high-priority stream of the panel / update-matrix broadcasts (top fronts)
    cudaEvent_t ev_p = nullptr, ev_b = nullptr;   // main -> comm ("data ready"), comm -> main ("broadcast done")
    PubRange* d_slab = nullptr;      // nkp_solve_dist: row slabs of all ranks (+ one slot for this rank's own)
    std::vector<PubRange> h_slab;
    PubRange* d_pub = nullptr;       // Plan::pub
    double* d_pack = nullptr;        // n x MAX_NR staging of the published solution ranges
    double* d_W = nullptr;      // solve work vectors, MAX_NR columns
    double* d_y = nullptr;      // n x MAX_NR permuted rhs / solution
    double* d_r = nullptr;      // n x MAX_NR residual
    double* d_xb = nullptr;     // n x MAX_NR staging for host-pointer solves (B)
    double* d_x = nullptr;      // n x MAX_NR solution accumulator
    double* d_berr = nullptr;   // MAX_NR (+ MAX_NR sums)
    double* d_sumsq = nullptr;  // per-block partial sums of k_sumsq (fixed-order final reduction)
    int* d_nrepl = nullptr;
    // tracer fields (nkp_set_tracer_maps / nkp_solve_fields)
    int tsl = 0, ct = 0;        // tracer_state_len, coupled_tracer_cnt
    int64_t ncell = 0;          // imt * jmt * km
    int* d_cell = nullptr;      // flat (k, j, i) cell index of every tracer-state entry
    double* d_fields = nullptr; // MAX_NR * ct device copies of the 3-D fields
    double* h_field = nullptr;  // pinned staging, one slot per field of a batch
    std::vector<cudaEvent_t> ev_field;   // per-field D2H completion
    double* h_pinned = nullptr; // pinned staging for values / rhs
    size_t pinned_bytes = 0;
    bool factored = false;
    double amax = 0;
    // stats
    double t_analysis = 0, t_factor = 0, t_scatter = 0, t_solve = 0;
    int refine_steps = 0, tiny_pivots = 0;
    int64_t launches = 0;
    // optional per-kernel-class event profiling of the factorisation
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;
    std::vector<int> prof_cls;
    int prof_used = 0;
    double t_cls[5] = {0, 0, 0, 0, 0};
    int64_t n_cls[5] = {0, 0, 0, 0, 0};
    double t_sweeps = 0;
    std::vector<cudaEvent_t> trace_ev;   // verbose >= 2: timeline of the last factorisation
    std::vector<std::string> trace_name;
};

enum { KC_OTHER = 0, KC_ADD = 1, KC_DIAG = 2, KC_TRSM = 3, KC_GEMM = 4 };

// record "everything launched so far belongs to class cls" on the solver's stream
static void prof_mark(nkp_solver* s, int cls) {
    if (!s->prof_on) return;
    if (s->prof_used == (int)s->prof_ev.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        s->prof_ev.push_back(e);
        s->prof_cls.push_back(0);
    }
    cudaEventRecord(s->prof_ev[s->prof_used], s->stream);
    s->prof_cls[s->prof_used] = cls;
    s->prof_used++;
}

static void prof_collect(nkp_solver* s) {
    for (int c = 0; c < 5; c++) s->t_cls[c] = 0, s->n_cls[c] = 0;
    for (int i = 1; i < s->prof_used; i++) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, s->prof_ev[i - 1], s->prof_ev[i]) != cudaSuccess) continue;
        s->t_cls[s->prof_cls[i]] += ms * 1e-3;
        s->n_cls[s->prof_cls[i]]++;
    }
    s->prof_used = 0;
}

template <class T>
static int upload(T** dptr, const std::vector<T>& v) {
    size_t bytes = sizeof(T) * (v.size() ? v.size() : 1);
    CK(cudaMalloc((void**)dptr, bytes));
    if (!v.empty()) CK(cudaMemcpy(*dptr, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
    return 0;
}

const char* nkp_last_error(void) { return g_err.c_str(); }
const char* nkp_version(void) { return "nkprecond-b200 0.1 (sm_100a)"; }

int nkp_set_analysis_cache(const char* dir) {
    set_analysis_cache_dir(dir);
    return NKP_OK;
}

void nkp_default_options(nkp_options* o) {
    memset(o, 0, sizeof(*o));
    o->nb = 64;
    o->leaf = 96;
    o->equil = 1;
    o->refine_max = 20;   // ITMAX of pdgsrfs
    o->device = 0;
    o->verbose = 0;
    o->residual_extra = 1;   // see include/nkprecond.h: costs nothing (the SpMV is HBM bound), converges in fewer steps
    const char* e;
    if ((e = getenv("NKP_LEAF"))) o->leaf = atoi(e);
    if ((e = getenv("NKP_VERBOSE"))) o->verbose = atoi(e);
    if ((e = getenv("NKP_EQUIL"))) o->equil = atoi(e);
    if ((e = getenv("NKP_REFINE_RULE"))) o->refine_rule = atoi(e);
    if ((e = getenv("NKP_RESIDUAL_EXTRA"))) o->residual_extra = atoi(e);
}

static int create_impl(nkp_solver** out, int n, const int* rowptr, const int* colind, const int* ci,
                       const int* cj, const int* ck, const nkp_options* opt_in, int rank, int nranks,
                       const void* unique_id, int** d_rowptr_adopt = nullptr, int** d_colind_adopt = nullptr,
                       const int* rowmap = nullptr, const double* row_scale = nullptr, const double* col_scale = nullptr) {
    if (!out || n <= 0 || !rowptr || !colind || rank < 0 || nranks < 1 || rank >= nranks ||
        (nranks > 1 && !unique_id)) {
        g_err = "nkp_create: invalid argument";
        return NKP_EINVAL;
    }
    nkp_options o;
    if (opt_in) o = *opt_in;
    else nkp_default_options(&o);
    if (o.nb <= 0 || o.nb > NBMAX || o.nb % G_TN != 0) {
        g_err = "nkp_create: nb must be 64";
        return NKP_EINVAL;
    }
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev <= 0) {
        g_err = "no CUDA device";
        return NKP_ECUDA;
    }
    CK(cudaSetDevice(o.device));

    nkp_solver* s = new nkp_solver();
    s->opt = o;
    s->n = n;
    s->nnz = rowptr[n];
    Options po;
    po.nb = o.nb;
    po.leaf = o.leaf;
    po.tm = G_TM;
    po.tn = G_TN;
    po.trsm_rows = TRSM_ROWS;
    po.add_tile = ADD_TILE;
    po.verbose = (rank == 0) ? o.verbose : 0;
    po.rank = rank;
    po.nranks = nranks;
    s->rank = rank;
    s->nranks = nranks;
    if (d_rowptr_adopt && d_colind_adopt) {   // pattern already on the device (nkp_create_be): ours from here on
        s->d_rowptr = *d_rowptr_adopt;
        s->d_colind = *d_colind_adopt;
        *d_rowptr_adopt = *d_colind_adopt = nullptr;
    }
    if (getenv("NKP_BIG_ENTRIES")) po.big_entries = atoll(getenv("NKP_BIG_ENTRIES"));
    if (getenv("NKP_BIG_ROWS")) po.big_rows = atoi(getenv("NKP_BIG_ROWS"));
    if (getenv("NKP_OUTER")) po.outer = po.top_outer = std::max(1, atoi(getenv("NKP_OUTER")));
    if (getenv("NKP_TOP_OUTER")) po.top_outer = std::max(1, atoi(getenv("NKP_TOP_OUTER")));
    if (getenv("NKP_SPLIT_TOL")) po.split_tol = atof(getenv("NKP_SPLIT_TOL"));
    if (getenv("NKP_SPLIT_MAX")) po.split_max = std::max(1, atoi(getenv("NKP_SPLIT_MAX")));
    if (getenv("NKP_SUPERNODES")) po.etree_supernodes = atoi(getenv("NKP_SUPERNODES"));
    if (getenv("NKP_RELAX_FRAC")) po.relax_frac = atof(getenv("NKP_RELAX_FRAC"));
    if (getenv("NKP_RELAX_SMALL")) po.relax_small = atoi(getenv("NKP_RELAX_SMALL"));
    const int* coords[3] = {ci, cj, ck};
    auto t0 = std::chrono::steady_clock::now();
    int rc = analyse(n, rowptr, colind, (ci || cj || ck) ? coords : nullptr, po, s->plan, rowmap);
    s->t_analysis = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rc) {
        char b[128];
        snprintf(b, sizeof b, "analysis failed with code %d", rc);
        g_err = b;
        nkp_destroy(s);
        return NKP_EANALYSIS;
    }
    Plan& P = s->plan;
#define CKD(x)                 \
    do {                       \
        int r_ = (x);          \
        if (r_) {              \
            nkp_destroy(s);    \
            return r_;         \
        }                      \
    } while (0)
    auto body = [&]() -> int {
        CK(cudaStreamCreate(&s->stream));
        for (int i = 0; i < 4; i++) CK(cudaEventCreate(&s->ev[i]));
        for (int i = 0; i < MAX_NR; i++) CK(cudaEventCreateWithFlags(&s->ev_col[i], cudaEventDisableTiming));
        if (nranks > 1) {
            if (!nccl_load()) {
                g_err = "cannot load libnccl.so.2";
                return NKP_ECUDA;
            }
            ncclUniqueId id;
            static_assert(sizeof(ncclUniqueId) <= NKP_UNIQUE_ID_BYTES, "unique id size");
            memcpy(&id, unique_id, sizeof(id));
            CKN(ncclCommInitRank(&s->comm, nranks, id, rank));
            // one communicator per group of a top front (every rank takes part in every split)
            s->gcomm.assign(P.groups.size(), nullptr);
            for (size_t q = 0; q < P.groups.size(); q++) {
                const std::vector<int>& grp = P.groups[q];
                const bool member = std::binary_search(grp.begin(), grp.end(), rank);
                if ((int)grp.size() == nranks) {
                    s->gcomm[q] = s->comm;
                    continue;
                }
                if (grp.size() < 2) continue;
                ncclComm_t sub = nullptr;
                CKN(ncclCommSplit(s->comm, member ? 0 : NCCL_SPLIT_NOCOLOR, rank, &sub, nullptr));
                s->gcomm[q] = member ? sub : nullptr;
            }
            int least = 0, greatest = 0;
            CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
            CK(cudaStreamCreateWithPriority(&s->cstream, cudaStreamNonBlocking, greatest));
            CK(cudaEventCreateWithFlags(&s->ev_p, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&s->ev_b, cudaEventDisableTiming));
            if (upload(&s->d_pub, P.pub)) return NKP_ECUDA;
            CK(cudaMalloc((void**)&s->d_pack, sizeof(double) * (size_t)n * MAX_NR));
        }
        CK(cudaMalloc((void**)&s->heap, sizeof(double) * (size_t)std::max<int64_t>(P.heap_len, 1)));
        std::vector<int> rp(rowptr, rowptr + n + 1), cidx(colind, colind + s->nnz), ridx((size_t)s->nnz);
        for (int i = 0; i < n; i++)
            for (int p = rowptr[i]; p < rowptr[i + 1]; p++) ridx[p] = i;
        if (!s->d_rowptr) {
            if (upload(&s->d_rowptr, rp)) return NKP_ECUDA;
            if (upload(&s->d_colind, cidx)) return NKP_ECUDA;
        }
        if (upload(&s->d_rowidx, ridx)) return NKP_ECUDA;
        if (upload(&s->d_scatter, P.scatter)) return NKP_ECUDA;
        if (upload(&s->d_perm, P.perm)) return NKP_ECUDA;
        if (rowmap) {
            std::vector<int> permr((size_t)n);
            for (int i = 0; i < n; i++) permr[i] = P.perm[rowmap[i]];
            if (upload(&s->d_permr, permr)) return NKP_ECUDA;
        } else {
            s->d_permr = s->d_perm;
        }
        if (upload(&s->d_bidx, P.bidx)) return NKP_ECUDA;
        if (upload(&s->d_rel, P.rel)) return NKP_ECUDA;
        if (upload(&s->d_clo, P.child_lo)) return NKP_ECUDA;
        if (upload(&s->d_diag, P.diag_tasks)) return NKP_ECUDA;
        if (upload(&s->d_trsm, P.trsm_tasks)) return NKP_ECUDA;
        if (upload(&s->d_gemm, P.gemm_tasks)) return NKP_ECUDA;
        if (upload(&s->d_add, P.add_tasks)) return NKP_ECUDA;
        if (upload(&s->d_solve, P.solve_tasks)) return NKP_ECUDA;
        if (upload(&s->d_children, P.solve_children)) return NKP_ECUDA;
        if (upload(&s->d_small, P.solve_small)) return NKP_ECUDA;
        if (upload(&s->d_big, P.big_fronts)) return NKP_ECUDA;
        if (upload(&s->d_fwd_items, P.big_fwd_items)) return NKP_ECUDA;
        if (upload(&s->d_bwd_items, P.big_bwd_items)) return NKP_ECUDA;
        if (upload(&s->d_rect_items, P.big_rect_items)) return NKP_ECUDA;
        CK(cudaMalloc((void**)&s->d_part, sizeof(double) * 512 * (size_t)std::max<int64_t>(P.bwd_part_slots, 1)));
        if (upload(&s->d_inv, P.inv_tasks)) return NKP_ECUDA;
        CK(cudaMalloc((void**)&s->d_cnt, sizeof(unsigned long long) * (2 * P.big_fronts.size() + 2)));
        CK(cudaMemset(s->d_cnt, 0, sizeof(unsigned long long) * (2 * P.big_fronts.size() + 2)));
        {
            cudaDeviceProp prop;
            CK(cudaGetDeviceProperties(&prop, o.device));
            int occ = 0, minocc = 1 << 30;
            CK(cudaFuncSetAttribute(k_sweep_big<SWEEP_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_SMEM));
            CK(cudaFuncSetAttribute(k_sweep_big<SWEEP_BWD_TRI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_SMEM));
            CK(cudaFuncSetAttribute(k_sweep_big<SWEEP_BWD_RECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_SMEM));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sweep_big<SWEEP_FWD>, 256, SW_SMEM));
            minocc = std::min(minocc, occ);
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sweep_big<SWEEP_BWD_TRI>, 256, SW_SMEM));
            minocc = std::min(minocc, occ);
            if (!prop.cooperativeLaunch || minocc < 1) {
                g_err = "device cannot run the cooperative dataflow sweeps";
                return NKP_ECUDA;
            }
            s->coop_ctas = prop.multiProcessorCount * std::min(minocc, 2);
            s->num_sms = prop.multiProcessorCount;
            s->small_v1 = getenv("NKP_SMALL_V1") ? atoi(getenv("NKP_SMALL_V1")) : 0;
            s->small_force = getenv("NKP_SMALL_FORCE") && atoi(getenv("NKP_SMALL_FORCE"));
            s->small_wcap.assign(P.nlevels, 0);
            int wmax = 0;
            for (int l = 0; l < P.nlevels; l++) {
                int w = 0;
                for (int q = P.levels[l].small_begin; q < P.levels[l].small_end; q++) w = std::max(w, P.solve_small[q].ld);
                s->small_wcap[l] = (w + 15) & ~15;
                wmax = std::max(wmax, s->small_wcap[l]);
            }
            if (wmax > SF_MAXLD || sf_cs(wmax) * 4 > SF_STAGE_WIDE) {
                g_err = "a small front exceeds the shared-memory work vector of the front sweep kernels";
                return NKP_EANALYSIS;
            }
            const int sf_smem = (int)(sizeof(double) * ((size_t)wmax * 8 + 1024 + (size_t)SF_NST_WIDE * SF_STAGE_WIDE));
            const int sf_smem_n = (int)(sizeof(double) * ((size_t)SF_NARROW_MAXLD * 8 + 1024 + (size_t)SF_NST_NARROW * SF_STAGE_NARROW));
            CK(cudaFuncSetAttribute(k_fwd_front<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, sf_smem));
            CK(cudaFuncSetAttribute(k_bwd_front<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, sf_smem));
            CK(cudaFuncSetAttribute(k_fwd_front<128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, sf_smem_n));
            CK(cudaFuncSetAttribute(k_bwd_front<128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, sf_smem_n));
        }
        CK(cudaMalloc((void**)&s->d_val, sizeof(double) * (size_t)s->nnz));
        CK(cudaMalloc((void**)&s->d_R, sizeof(double) * n));
        CK(cudaMalloc((void**)&s->d_C, sizeof(double) * n));
        if (row_scale && col_scale) {
            // rounded to powers of two like the solver's own equilibration: scaling then adds no rounding error
            std::vector<double> sc((size_t)n);
            for (int pass = 0; pass < 2; pass++) {
                const double* src = pass ? col_scale : row_scale;
                for (int i = 0; i < n; i++) {
                    if (!(src[i] > 0) || !std::isfinite(src[i])) {
                        g_err = "nkp_create_rowperm: scalings must be positive and finite";
                        return NKP_EINVAL;
                    }
                    sc[i] = std::ldexp(1.0, (int)std::lround(std::log2(src[i])));
                }
                CK(cudaMemcpy(pass ? s->d_C : s->d_R, sc.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
            }
            s->fixed_scale = true;
        }
        CK(cudaMalloc((void**)&s->d_W, sizeof(double) * (size_t)std::max<int64_t>(P.solve_pool_len, 1) * MAX_NR));
        CK(cudaMalloc((void**)&s->d_y, sizeof(double) * (size_t)n * MAX_NR));
        CK(cudaMalloc((void**)&s->d_r, sizeof(double) * (size_t)n * MAX_NR));
        CK(cudaMalloc((void**)&s->d_x, sizeof(double) * (size_t)n * MAX_NR));
        CK(cudaMalloc((void**)&s->d_xb, sizeof(double) * (size_t)n * MAX_NR));
        CK(cudaMalloc((void**)&s->d_berr, sizeof(double) * 4 * MAX_NR));
        CK(cudaMalloc((void**)&s->d_sumsq, sizeof(double) * SUMSQ_BLOCKS * MAX_NR));
        CK(cudaMalloc((void**)&s->d_nrepl, sizeof(int)));
        s->pinned_bytes = sizeof(double) * std::max<size_t>((size_t)s->nnz, (size_t)n * MAX_NR);
        CK(cudaMallocHost((void**)&s->h_pinned, s->pinned_bytes));
        CK(cudaFuncSetAttribute(k_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM));
        CK(cudaFuncSetAttribute(k_trsm, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSM_SMEM));
        // the scatter map and the heap can be large: release the host copies
        std::vector<int64_t>().swap(P.scatter);
        return 0;
    };
    rc = body();
    if (rc) {
        nkp_destroy(s);
        return rc;
    }
    // drop the host copy of big plan arrays we no longer need on the host
    *out = s;
    return NKP_OK;
}

int nkp_create(nkp_solver** out, int n, const int* rowptr, const int* colind, const int* ci,
               const int* cj, const int* ck, const nkp_options* opt_in) {
    return create_impl(out, n, rowptr, colind, ci, cj, ck, opt_in, 0, 1, nullptr);
}

extern "C" int nkp_bswap32_device(void* d_data, long long count);

// Pattern arrays in FILE byte order: the big-endian NC_INT bytes of `rowptr` and `colind` as they lie in the matrix
// file go to the device unchanged and are converted there (k_bswap32); the host-order copy the analysis needs comes
// back from the device.  Replaces the host loops behind nc_get_var_int in get_sparse_matrix (src/matrix.c:3944-4031).
int nkp_create_be(nkp_solver** out, int n, long long nnz, const void* rowptr_be, const void* colind_be, const int* ci,
                  const int* cj, const int* ck, const nkp_options* opt_in) {
    if (!out || n <= 0 || nnz <= 0 || nnz > 2147483647LL || !rowptr_be || !colind_be) {
        g_err = "nkp_create_be: invalid argument";
        return NKP_EINVAL;
    }
    nkp_options o;
    if (opt_in) o = *opt_in;
    else nkp_default_options(&o);
    CK(cudaSetDevice(o.device));
    int *d_rp = nullptr, *d_ci = nullptr;
    CK(cudaMalloc((void**)&d_rp, sizeof(int) * ((size_t)n + 1)));
    CK(cudaMalloc((void**)&d_ci, sizeof(int) * (size_t)nnz));
    std::vector<int> rp((size_t)n + 1), cidx((size_t)nnz);
    auto body = [&]() -> int {
        CK(cudaMemcpy(d_rp, rowptr_be, sizeof(int) * ((size_t)n + 1), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_ci, colind_be, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice));
        if (nkp_bswap32_device(d_rp, (long long)n + 1) || nkp_bswap32_device(d_ci, nnz)) {
            g_err = "nkp_create_be: byte-swap kernel failed";
            return NKP_ECUDA;
        }
        CK(cudaMemcpy(rp.data(), d_rp, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(cidx.data(), d_ci, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToHost));
        if (rp[0] != 0 || rp[n] != nnz) {
            g_err = "nkp_create_be: rowptr does not match nnz (wrong byte order or variable?)";
            return NKP_EINVAL;
        }
        return 0;
    };
    int rc = body();
    if (rc) {
        cudaFree(d_rp);
        cudaFree(d_ci);
        return rc;
    }
    rc = create_impl(out, n, rp.data(), cidx.data(), ci, cj, ck, &o, 0, 1, nullptr, &d_rp, &d_ci);
    // create_impl takes the two arrays over (and nulls our pointers) as soon as its solver object exists
    if (d_rp) cudaFree(d_rp);
    if (d_ci) cudaFree(d_ci);
    return rc;
}

int nkp_create_dist(nkp_solver** out, int n, const int* rowptr, const int* colind, const int* ci,
                    const int* cj, const int* ck, const nkp_options* opt_in, int rank, int nranks,
                    const void* unique_id) {
    return create_impl(out, n, rowptr, colind, ci, cj, ck, opt_in, rank, nranks, unique_id);
}

// static row permutation (+ scalings) chosen by the caller, e.g. with nkp_rowperm_largediag (rowperm.cpp)
int nkp_create_rowperm(nkp_solver** out, int n, const int* rowptr, const int* colind, const int* ci, const int* cj,
                       const int* ck, const nkp_options* opt_in, const int* rowmap, const double* row_scale,
                       const double* col_scale, int rank, int nranks, const void* unique_id) {
    if (!rowmap || (row_scale == nullptr) != (col_scale == nullptr)) {
        g_err = "nkp_create_rowperm: rowmap is required; row_scale and col_scale come together or not at all";
        return NKP_EINVAL;
    }
    return create_impl(out, n, rowptr, colind, ci, cj, ck, opt_in, rank, nranks, unique_id, nullptr, nullptr, rowmap,
                       row_scale, col_scale);
}

int nkp_comm_unique_id(void* unique_id) {
    if (!unique_id) return NKP_EINVAL;
    if (!nccl_load()) {
        g_err = "cannot load libnccl.so.2";
        return NKP_ECUDA;
    }
    ncclUniqueId id;
    CKN(ncclGetUniqueId(&id));
    memset(unique_id, 0, NKP_UNIQUE_ID_BYTES);
    memcpy(unique_id, &id, sizeof(id));
    return NKP_OK;
}


// One top front of the shared part of the tree, factored by its group (nkp_internal.hpp, TopFront): update
// matrices of the children to every member, extend-add by every member into its own full copy, then per outer
// block: panel factorisation by the owner, broadcast of the two factored column blocks, wide Schur update of the
// blocks each member owns -- the block that is factored next first, so that its panel and its broadcast overlap
// everybody's remaining update.  All NCCL calls go to the high-priority communication stream; events carry the
// dependencies between it and the compute stream.
// NKP_CHECK=1: count the non-finite entries of a device array and report them (diagnostics; synchronises)
static void check_finite(nkp_solver* s, const char* what, const double* p, int64_t n) {
    static const bool on = getenv("NKP_CHECK") != nullptr;
    if (!on) return;
    unsigned long long* d = nullptr;
    unsigned long long h = 0;
    cudaMalloc((void**)&d, sizeof(h));
    cudaMemsetAsync(d, 0, sizeof(h), s->stream);
    k_count_nonfinite<<<1024, 256, 0, s->stream>>>(p, n, d);
    cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, s->stream);
    cudaStreamSynchronize(s->stream);
    cudaFree(d);
    if (h) fprintf(stderr, "[nkp] CHECK rank %d: %llu non-finite entries in %s\n", s->rank, h, what);
}

// verbose >= 2: timeline marks on the compute stream (name, event), printed after the factorisation
static void trace_mark(nkp_solver* s, const char* what, int a, int b) {
    if (s->opt.verbose < 2) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, s->stream);
    char buf[96];
    snprintf(buf, sizeof buf, "%s %d/%d", what, a, b);
    s->trace_ev.push_back(e);
    s->trace_name.push_back(buf);
}

static int factor_top_front(nkp_solver* s, const TopFront& tf, double tiny) {
    if (!tf.member) return 0;
    Plan& P = s->plan;
    cudaStream_t st = s->stream, cs = s->cstream;
    trace_mark(s, "top front start", tf.front, P.fronts[tf.front].s);
    const int nb = P.opt.nb;
    const std::vector<int>& grp = P.groups[tf.group];
    const int g = (int)grp.size();
    ncclComm_t gc = s->gcomm[tf.group];
    auto bcast = [&](const TopBcast& b) -> int {
        const int root = (int)(std::lower_bound(grp.begin(), grp.end(), b.root) - grp.begin());
        CKN(ncclBroadcast(s->heap + b.off, s->heap + b.off, (size_t)b.count, ncclDouble, root, gc, cs));
        return 0;
    };
    if (g > 1 && tf.cb_end > tf.cb_begin) {
        CK(cudaEventRecord(s->ev_p, st));
        CK(cudaStreamWaitEvent(cs, s->ev_p, 0));
        CKN(ncclGroupStart());
        for (int q = tf.cb_begin; q < tf.cb_end; q++)
            if (bcast(P.top_child_bcasts[q])) return NKP_ECUDA;
        CKN(ncclGroupEnd());
        CK(cudaEventRecord(s->ev_b, cs));
        CK(cudaStreamWaitEvent(st, s->ev_b, 0));
        prof_mark(s, KC_OTHER);
        trace_mark(s, "  children bcast", tf.front, tf.cb_end - tf.cb_begin);
    }
    for (size_t pass = 0; pass + 1 < tf.add_begin.size(); pass++) {
        const int nt = tf.add_begin[pass + 1] - tf.add_begin[pass];
        if (nt == 0 || tf.add_tiles[pass] == 0) continue;
        k_extend_add<<<tf.add_tiles[pass], 256, 0, st>>>(s->d_add + tf.add_begin[pass], nt, s->d_rel, s->heap, nb);
        s->launches++;
        prof_mark(s, KC_ADD);
    }
    auto panel = [&](const TopBlock& tb) {
        for (int q = tb.step_begin; q < tb.step_end; q++) {
            const TopStep& ts = P.top_steps[q];
            if (ts.diag_end > ts.diag_begin) {
                k_diag<<<ts.diag_end - ts.diag_begin, 256, 0, st>>>(s->d_diag + ts.diag_begin, s->heap, tiny, s->d_nrepl);
                s->launches++;
                prof_mark(s, KC_DIAG);
                if (s->opt.verbose >= 4) trace_mark(s, "      diag", q - tb.step_begin, 0);
            }
            if (ts.trsm_ctas > 0) {
                k_trsm<<<ts.trsm_ctas, TRSM_THREADS, TRSM_SMEM, st>>>(s->d_trsm + ts.trsm_begin, ts.trsm_end - ts.trsm_begin, s->heap);
                s->launches++;
                prof_mark(s, KC_TRSM);
                if (s->opt.verbose >= 4) trace_mark(s, "      trsm ctas", q - tb.step_begin, ts.trsm_ctas);
            }
            if (ts.gemm_tiles > 0) {
                k_gemm<<<ts.gemm_tiles, 256, G_SMEM, st>>>(s->d_gemm + ts.gemm_begin, ts.gemm_end - ts.gemm_begin, s->heap, nb);
                s->launches++;
                prof_mark(s, KC_GEMM);
                if (s->opt.verbose >= 4) trace_mark(s, "      narrow gemm tiles", q - tb.step_begin, ts.gemm_tiles);
            }
        }
    };
    trace_mark(s, "  extend-add", tf.front, 0);
    const int nK = tf.block_end - tf.block_begin;
    for (int K = 0; K < nK; K++) {
        const TopBlock& tb = P.top_blocks[tf.block_begin + K];
        const bool own = tb.owner == s->rank;
        if (s->opt.verbose >= 3) trace_mark(s, "    block", K, nK);
        if (K == 0 && own) {
            panel(tb);
            if (g > 1) CK(cudaEventRecord(s->ev_p, st));
        }
        if (g > 1) {
            if (own) CK(cudaStreamWaitEvent(cs, s->ev_p, 0));
            CKN(ncclGroupStart());
            if (bcast(tb.bl) || bcast(tb.bu)) return NKP_ECUDA;
            CKN(ncclGroupEnd());
            CK(cudaEventRecord(s->ev_b, cs));
            if (!own) CK(cudaStreamWaitEvent(st, s->ev_b, 0));   // the owner already has what it sends
        }
        if (s->opt.verbose >= 4) trace_mark(s, "      wait bcast", K, own);
        if (tb.next_tiles > 0) {
            k_gemm<<<tb.next_tiles, 256, G_SMEM, st>>>(s->d_gemm + tb.next_begin, tb.next_end - tb.next_begin, s->heap, nb);
            s->launches++;
            prof_mark(s, KC_GEMM);
            if (s->opt.verbose >= 4) trace_mark(s, "      next gemm tiles", K, tb.next_tiles);
        }
        if (K + 1 < nK && P.top_blocks[tf.block_begin + K + 1].owner == s->rank) {
            panel(P.top_blocks[tf.block_begin + K + 1]);
            if (g > 1) CK(cudaEventRecord(s->ev_p, st));
        }
        if (tb.rest_tiles > 0) {
            k_gemm<<<tb.rest_tiles, 256, G_SMEM, st>>>(s->d_gemm + tb.rest_begin, tb.rest_end - tb.rest_begin, s->heap, nb);
            s->launches++;
            prof_mark(s, KC_GEMM);
            if (s->opt.verbose >= 4) trace_mark(s, "      rest gemm tiles", K, tb.rest_tiles);
        }
    }
    // every broadcast of this front (our own sends included) is complete before anything changes the panels again
    // (k_invert_diag) or reuses the children's update matrices
    if (g > 1) CK(cudaStreamWaitEvent(st, s->ev_b, 0));
    trace_mark(s, "  blocks", tf.front, nK);
    return 0;
}

static int do_factor(nkp_solver* s) {
    Plan& P = s->plan;
    cudaStream_t st = s->stream;
    const int n = s->n;
    const int64_t nnz = s->nnz;
    const int nb = P.opt.nb;
    CK(cudaEventRecord(s->ev[0], st));
    // equilibration
    if (s->fixed_scale) {
        // R and C came with the static row permutation and stay (SamePattern_SameRowPerm)
    } else if (s->opt.equil) {
        k_row_scale<<<(n + 255) / 256, 256, 0, st>>>(n, s->d_rowptr, s->d_val, s->d_R);
        CK(cudaMemsetAsync(s->d_C, 0, sizeof(double) * n, st));
        k_col_max<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(nnz, s->d_rowidx, s->d_colind, s->d_val, s->d_R, s->d_C);
        k_col_scale<<<(n + 255) / 256, 256, 0, st>>>(n, s->d_C);
        s->launches += 3;
    } else {
        k_fill<<<(n + 255) / 256, 256, 0, st>>>(s->d_R, n, 1.0);
        k_fill<<<(n + 255) / 256, 256, 0, st>>>(s->d_C, n, 1.0);
        s->launches += 2;
    }
    // zero the factor arena, scatter A
    CK(cudaMemsetAsync(s->heap, 0, sizeof(double) * (size_t)P.factor_len, st));
    CK(cudaMemsetAsync(s->d_nrepl, 0, sizeof(int), st));
    k_scatter<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(nnz, s->d_val, s->d_scatter, s->d_rowidx, s->d_colind,
                                                              s->d_R, s->d_C, s->heap);
    s->launches++;
    CK(cudaEventRecord(s->ev[1], st));
    s->prof_used = 0;
    prof_mark(s, KC_OTHER);
    // with equilibration every row/column max is in [1,2): threshold relative to ||A|| ~ 1
    double tiny = std::sqrt(2.220446049250313e-16) * ((s->opt.equil || s->fixed_scale) ? 1.0 : s->amax);
    for (cudaEvent_t e : s->trace_ev) cudaEventDestroy(e);
    s->trace_ev.clear();
    s->trace_name.clear();
    trace_mark(s, "scatter done", 0, 0);
    for (int l = P.nlevels - 1; l >= 0; l--) {
        const LevelPlan& L = P.levels[l];
        if (L.f22_zero_len > 0) {
            CK(cudaMemsetAsync(s->heap + L.f22_zero_off, 0, sizeof(double) * (size_t)L.f22_zero_len, st));
            prof_mark(s, KC_OTHER);
        }
        int npass = (int)L.add_tiles.size();
        for (int pass = 0; pass < npass; pass++) {
            int nt = L.add_begin[pass + 1] - L.add_begin[pass];
            if (nt == 0 || L.add_tiles[pass] == 0) continue;
            k_extend_add<<<L.add_tiles[pass], 256, 0, st>>>(s->d_add + L.add_begin[pass], nt, s->d_rel, s->heap, nb);
            s->launches++;
            prof_mark(s, KC_ADD);
        }
        for (int step = 0; step < L.nsteps; step++) {
            int nd = L.diag_begin[step + 1] - L.diag_begin[step];
            if (nd > 0) {
                k_diag<<<nd, 256, 0, st>>>(s->d_diag + L.diag_begin[step], s->heap, tiny, s->d_nrepl);
                s->launches++;
                prof_mark(s, KC_DIAG);
            }
            int ntr = L.trsm_begin[step + 1] - L.trsm_begin[step];
            if (ntr > 0 && L.trsm_ctas[step] > 0) {
                k_trsm<<<L.trsm_ctas[step], TRSM_THREADS, TRSM_SMEM, st>>>(s->d_trsm + L.trsm_begin[step], ntr, s->heap);
                s->launches++;
                prof_mark(s, KC_TRSM);
            }
            int ng = L.gemm_begin[step + 1] - L.gemm_begin[step];
            if (ng > 0 && L.gemm_tiles[step] > 0) {
                k_gemm<<<L.gemm_tiles[step], 256, G_SMEM, st>>>(s->d_gemm + L.gemm_begin[step], ng, s->heap, nb);
                s->launches++;
                prof_mark(s, KC_GEMM);
            }
        }
        if (!L.mine.empty()) trace_mark(s, "level (rank-private fronts)", l, (int)L.mine.size());
        // the top fronts of this level: factored together with the other members of their groups
        for (int ti : L.tops)
            if (int rc = factor_top_front(s, P.top_fronts[ti], tiny)) return rc;
        // the sweeps use inverted 64 x 64 diagonal blocks; nothing above this level reads them.  (Running
        // this on a second stream next to the upper levels was measured slower: its small CTAs displace
        // Schur-update CTAs.)
        if (L.inv_end > L.inv_begin) {
            k_invert_diag<<<dim3((unsigned)(L.inv_end - L.inv_begin), 2), INV_THREADS, 0, st>>>(s->d_inv + L.inv_begin, s->heap);
            s->launches++;
            prof_mark(s, KC_DIAG);
        }
    }
    CK(cudaEventRecord(s->ev[2], st));
    CK(cudaGetLastError());
    check_finite(s, "the factors", s->heap, P.factor_len);
    int nrepl = 0;
    CK(cudaMemcpyAsync(&nrepl, s->d_nrepl, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    float ms01 = 0, ms02 = 0;
    CK(cudaEventElapsedTime(&ms01, s->ev[0], s->ev[1]));
    CK(cudaEventElapsedTime(&ms02, s->ev[0], s->ev[2]));
    s->t_scatter = ms01 * 1e-3;
    s->t_factor = ms02 * 1e-3;
    s->tiny_pivots = nrepl;
    s->factored = true;
    if (s->prof_on) prof_collect(s);
    if (s->opt.verbose)
        fprintf(stderr, "[nkp] rank %d factor: %.3f ms (scatter %.3f ms), %.2f TFLOP/s, tiny pivots replaced: %d\n", s->rank,
                ms02, ms01, P.flops / (ms02 * 1e-3) * 1e-12, nrepl);
    for (size_t i = 1; i < s->trace_ev.size(); i++) {
        float ms = 0, tot = 0;
        cudaEventElapsedTime(&ms, s->trace_ev[i - 1], s->trace_ev[i]);
        cudaEventElapsedTime(&tot, s->trace_ev[0], s->trace_ev[i]);
        if (ms >= 0.5f || s->opt.verbose >= 3)
            fprintf(stderr, "[nkp] rank %d   %-40s %9.3f ms  (at %9.3f)\n", s->rank, s->trace_name[i].c_str(), ms, tot);
    }
    return NKP_OK;
}

// max |A| of the values in d_val (tiny-pivot threshold without equilibration)
static int device_amax(nkp_solver* s) {
    CK(cudaMemsetAsync(s->d_berr + 3 * MAX_NR, 0, sizeof(double), s->stream));
    k_absmax<<<1024, 256, 0, s->stream>>>(s->nnz, s->d_val, s->d_berr + 3 * MAX_NR);
    s->launches++;
    CK(cudaMemcpyAsync(&s->amax, s->d_berr + 3 * MAX_NR, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

int nkp_factor_device(nkp_solver* s, const double* d_nzval) {
    if (!s || !d_nzval) {
        g_err = "nkp_factor_device: invalid argument";
        return NKP_EINVAL;
    }
    CK(cudaSetDevice(s->opt.device));
    if (d_nzval != s->d_val)
        CK(cudaMemcpyAsync(s->d_val, d_nzval, sizeof(double) * (size_t)s->nnz, cudaMemcpyDeviceToDevice, s->stream));
    if (!s->opt.equil && !s->fixed_scale && device_amax(s)) return NKP_ECUDA;
    return do_factor(s);
}

int nkp_factor(nkp_solver* s, const double* nzval) {
    if (!s || !nzval) {
        g_err = "nkp_factor: invalid argument";
        return NKP_EINVAL;
    }
    CK(cudaSetDevice(s->opt.device));
    double amax = 0;
    if (!s->opt.equil)
        for (int64_t p = 0; p < s->nnz; p++) amax = std::max(amax, std::fabs(nzval[p]));
    s->amax = amax;
    if (host_is_pinned(nzval)) {
        CK(cudaMemcpyAsync(s->d_val, nzval, sizeof(double) * (size_t)s->nnz, cudaMemcpyHostToDevice, s->stream));
    } else {
        // pageable -> pinned staging (all cores) -> device, in four pieces so that DMA overlaps the host copies
        const size_t nnz = (size_t)s->nnz, piece = (nnz + 3) / 4;
        for (size_t o = 0; o < nnz; o += piece) {
            const size_t len = std::min(piece, nnz - o);
            par_memcpy(s->h_pinned + o, nzval + o, sizeof(double) * len);
            CK(cudaMemcpyAsync(s->d_val + o, s->h_pinned + o, sizeof(double) * len, cudaMemcpyHostToDevice, s->stream));
        }
    }
    return do_factor(s);
}

int nkp_factor_be(nkp_solver* s, const void* nzval_be) {
    if (!s || !nzval_be) {
        g_err = "nkp_factor_be: invalid argument";
        return NKP_EINVAL;
    }
    CK(cudaSetDevice(s->opt.device));
    s->amax = 0;
    // raw file bytes: (pageable -> pinned staging ->) device, then swapped in place
    if (host_is_pinned(nzval_be)) {
        CK(cudaMemcpyAsync(s->d_val, nzval_be, sizeof(double) * (size_t)s->nnz, cudaMemcpyHostToDevice, s->stream));
    } else {
        par_memcpy(s->h_pinned, nzval_be, sizeof(double) * (size_t)s->nnz);
        CK(cudaMemcpyAsync(s->d_val, s->h_pinned, sizeof(double) * (size_t)s->nnz, cudaMemcpyHostToDevice, s->stream));
    }
    k_bswap64<<<(unsigned)((s->nnz + 255) / 256), 256, 0, s->stream>>>(reinterpret_cast<const unsigned long long*>(s->d_val),
                                                                      s->d_val, s->nnz);
    s->launches++;
    CK(cudaGetLastError());
    if (!s->opt.equil && !s->fixed_scale && device_amax(s)) return NKP_ECUDA;
    return do_factor(s);
}

// forward + backward sweeps on d_y (n x nr, permuted, scaled), in place
// One forward + backward sweep pair.  The progress counters of the dataflow kernels carry the number of the sweep in
// their upper half, so counters left over from earlier sweeps never have to be cleared.  (Handing the items out from a
// device-wide queue and replaying the sequence as a CUDA graph were both tried this round and are NOT in the tree: the
// item loop restructured for the queue produced intermittent NaNs at gx1v6-shape -- profiles/r02_sweep_graph_experiment.txt.)
template <int NR>
static int sweeps(nkp_solver* s) {
    Plan& P = s->plan;
    cudaStream_t st = s->stream;
    s->epoch++;
    const int epoch = s->epoch;
    int n = s->n;
    const bool trace = s->opt.verbose >= 3;
    std::vector<cudaEvent_t> tev;
    std::vector<std::string> tname;
    auto mark = [&](const char* what, int level, int count) {
        if (!trace) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        tev.push_back(e);
        char b[96];
        snprintf(b, sizeof b, "%s level %d (%d)", what, level, count);
        tname.push_back(b);
    };
    mark("start", -1, 0);
    unsigned long long* cnt_f = s->d_cnt;
    unsigned long long* cnt_b = s->d_cnt + P.big_fronts.size();
    unsigned uepoch = (unsigned)epoch;
    int nr = NR, nrtot = NR;
    const double* heap = s->heap;
    for (int l = P.nlevels - 1; l >= 0; l--) {
        const LevelPlan& L = P.levels[l];
        if (s->nranks > 1 && l + 1 < P.nlevels) {
            // update vectors of children that live on another GPU (whole work vector: m x NR)
            const LevelPlan& Lc = P.levels[l + 1];
            bool any = false;
            for (int q : Lc.xfers) any = any || P.xfers[q].src == s->rank || P.xfers[q].dst == s->rank;
            if (any) {
                CKN(ncclGroupStart());
                for (int q : Lc.xfers) {
                    const Xfer& x = P.xfers[q];
                    const Front& f = P.fronts[x.front];
                    size_t cnt = (size_t)f.m * NR;
                    if (x.src == s->rank) CKN(ncclSend(s->d_W + f.woff * NR, cnt, ncclDouble, x.dst, s->comm, st));
                    if (x.dst == s->rank) CKN(ncclRecv(s->d_W + f.woff * NR, cnt, ncclDouble, x.src, s->comm, st));
                }
                CKN(ncclGroupEnd());
                mark("fwd xfer", l, (int)Lc.xfers.size());
            }
        }
        int nsmall = L.small_end - L.small_begin;
        if (nsmall > 0) {
            // few, larger fronts: persistent CTAs + bulk-copy ring; thousands of leaf-sized fronts: one warp per front
            // (measured: gx3v7-shape 20 fronts 0.150 -> 0.035 ms, 520 fronts 0.129 -> 0.088 ms; 2 093 leaves 0.114 vs 0.148 ms)
            if ((s->small_v1 & 1) || (!s->small_force && nsmall > 4 * s->num_sms)) {
                k_fwd_small<<<(nsmall + SMALL_WARPS - 1) / SMALL_WARPS, 32 * SMALL_WARPS, 0, st>>>(
                    s->d_small + L.small_begin, nsmall, s->d_children, s->d_rel, s->heap, s->d_W, s->d_y, s->n, nr, nrtot);
            } else {
                const int wcap = s->small_wcap[l];
                if (nsmall > 4 * s->num_sms && wcap <= SF_NARROW_MAXLD) {
                    const size_t smem = sizeof(double) * ((size_t)wcap * 8 + 512 + (size_t)SF_NST_NARROW * SF_STAGE_NARROW);
                    k_fwd_front<128, 4><<<nsmall, 128, smem, st>>>(s->d_small + L.small_begin, nsmall, s->d_children, s->d_rel, s->heap,
                                                                   s->d_W, s->d_y, s->n, nr, nrtot, wcap, SF_NST_NARROW, SF_STAGE_NARROW);
                } else {
                    const size_t smem = sizeof(double) * ((size_t)wcap * 8 + 512 + (size_t)SF_NST_WIDE * SF_STAGE_WIDE);
                    k_fwd_front<256, 1><<<std::min(nsmall, s->num_sms), 256, smem, st>>>(
                        s->d_small + L.small_begin, nsmall, s->d_children, s->d_rel, s->heap, s->d_W, s->d_y, s->n, nr, nrtot, wcap,
                        SF_NST_WIDE, SF_STAGE_WIDE);
                }
            }
            s->launches++;
            mark("fwd small", l, nsmall);
        }
        int nitems = L.fwd_item_end - L.fwd_item_begin;
        if (nitems > 0) {
            const BigFront* bfs = s->d_big;
            const BigItem* items = s->d_fwd_items + L.fwd_item_begin;
            const SolveChild* ch = s->d_children;
            const int* rel = s->d_rel;
            double* W = s->d_W;
            double* y = s->d_y;
            double* part = s->d_part;
            const int* clo = s->d_clo;
            void* args[] = {(void*)&bfs, (void*)&items, (void*)&nitems, (void*)&ch,    (void*)&rel,   (void*)&clo,   (void*)&heap, (void*)&W,
                            (void*)&y,   (void*)&part,  (void*)&n,      (void*)&nr,    (void*)&nrtot, (void*)&cnt_f, (void*)&uepoch};
            int grid = std::min(nitems, s->coop_ctas);
            CK(cudaLaunchCooperativeKernel((void*)k_sweep_big<SWEEP_FWD>, dim3(grid), dim3(256), args, SW_SMEM, st));
            s->launches++;
            mark("fwd big", l, nitems);
        }
    }
    for (int l = 0; l < P.nlevels; l++) {
        const LevelPlan& L = P.levels[l];
        int nsmall = L.small_end - L.small_begin;
        if (nsmall > 0) {
            if ((s->small_v1 & 2) || (!s->small_force && nsmall > 4 * s->num_sms)) {
                k_bwd_small<<<(nsmall + SMALL_WARPS - 1) / SMALL_WARPS, 32 * SMALL_WARPS, 0, st>>>(
                    s->d_small + L.small_begin, nsmall, s->d_bidx, s->heap, s->d_W, s->d_y, s->n, nr, nrtot);
            } else {
                const int wcap = s->small_wcap[l];
                if (nsmall > 4 * s->num_sms && wcap <= SF_NARROW_MAXLD) {
                    const size_t smem = sizeof(double) * ((size_t)wcap * 8 + 1024 + (size_t)SF_NST_NARROW * SF_STAGE_NARROW);
                    k_bwd_front<128, 4><<<nsmall, 128, smem, st>>>(s->d_small + L.small_begin, nsmall, s->d_bidx, s->heap, s->d_W, s->d_y,
                                                                   s->n, nr, nrtot, wcap, SF_NST_NARROW, SF_STAGE_NARROW);
                } else {
                    const size_t smem = sizeof(double) * ((size_t)wcap * 8 + 1024 + (size_t)SF_NST_WIDE * SF_STAGE_WIDE);
                    k_bwd_front<256, 1><<<std::min(nsmall, s->num_sms), 256, smem, st>>>(
                        s->d_small + L.small_begin, nsmall, s->d_bidx, s->heap, s->d_W, s->d_y, s->n, nr, nrtot, wcap, SF_NST_WIDE,
                        SF_STAGE_WIDE);
                }
            }
            s->launches++;
            mark("bwd small", l, nsmall);
        }
        int nitems = L.bwd_item_end - L.bwd_item_begin;
        if (nitems > 0) {
            const BigFront* bfs = s->d_big;
            const BigItem* items = s->d_bwd_items + L.bwd_item_begin;
            const SolveChild* ch = s->d_children;
            const int* rel = s->d_rel;
            double* W = s->d_W;
            double* y = s->d_y;
            // boundary values of the level's big fronts -> their work vectors
            int maxr = 0;
            for (int b = L.big_begin; b < L.big_end; b++) maxr = std::max(maxr, P.big_fronts[b].r);
            if (maxr > 0) {
                k_gather_bnd<<<dim3(L.big_end - L.big_begin, (maxr + 255) / 256), 256, 0, st>>>(
                    s->d_big + L.big_begin, s->d_bidx, s->d_y, s->d_W, n, nr, nrtot);
                s->launches++;
            }
            double* part = s->d_part;
            int nrect = L.rect_item_end - L.rect_item_begin;
            if (nrect > 0) {
                // rectangular part: independent items, plain launch
                k_sweep_big<SWEEP_BWD_RECT><<<std::min(nrect, 8 * s->coop_ctas), 256, SW_SMEM, st>>>(
                    s->d_big, s->d_rect_items + L.rect_item_begin, nrect, ch, rel, s->d_clo, heap, W, y, part, n, nr, nrtot,
                    cnt_b, uepoch);
                s->launches++;
                mark("bwd rect", l, nrect);
            }
            const int* clo = s->d_clo;
            void* args[] = {(void*)&bfs, (void*)&items, (void*)&nitems, (void*)&ch,    (void*)&rel,   (void*)&clo,   (void*)&heap, (void*)&W,
                            (void*)&y,   (void*)&part,  (void*)&n,      (void*)&nr,    (void*)&nrtot, (void*)&cnt_b, (void*)&uepoch};
            int grid = std::min(nitems, s->coop_ctas);
            CK(cudaLaunchCooperativeKernel((void*)k_sweep_big<SWEEP_BWD_TRI>, dim3(grid), dim3(256), args, SW_SMEM, st));
            s->launches++;
            mark("bwd big", l, nitems);
        }
    }
    if (s->nranks > 1 && !P.pub.empty()) {
        // The top fronts were swept by every member of their groups: no communication so far.  Now every part of
        // the solution is published by one rank that holds it (rank-private subtree ranges by their owner, top
        // fronts by the first member of their group): packed per range, ONE broadcast per range, one group call.
        const int nrg = (int)P.pub.size();
        k_pub_pack<<<dim3(64, nrg), 256, 0, st>>>(s->d_pub, s->rank, 0, s->d_y, n, NR, s->d_pack);
        CKN(ncclGroupStart());
        for (const PubRange& pr : P.pub) {
            if (pr.hi <= pr.lo) continue;
            double* p = s->d_pack + (size_t)pr.lo * NR;
            CKN(ncclBroadcast(p, p, (size_t)(pr.hi - pr.lo) * NR, ncclDouble, pr.root, s->comm, st));
        }
        CKN(ncclGroupEnd());
        k_pub_pack<<<dim3(64, nrg), 256, 0, st>>>(s->d_pub, s->rank, 1, s->d_y, n, NR, s->d_pack);
        s->launches += 2;
        mark("publish", -1, nrg);
    }
    CK(cudaGetLastError());
    if (trace) {
        cudaStreamSynchronize(st);
        for (size_t i = 1; i < tev.size(); i++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, tev[i - 1], tev[i]);
            fprintf(stderr, "[nkp] sweep %-28s %9.3f ms\n", tname[i].c_str(), ms);
        }
        for (cudaEvent_t e : tev) cudaEventDestroy(e);
    }
    return 0;
}

// r = b - A x for nr right-hand sides in ONE pass over A
static void launch_residual(nkp_solver* s, int nr, const double* x, int ldx, const double* b, int ldb, double* r, double* berr,
                            double safe1, double safe2) {
    const int g = (s->n + 255) / 256;
    cudaStream_t st = s->stream;
#define NKP_RES(NRT, EX) k_residual<NRT, EX><<<g, 256, 0, st>>>(s->n, nr, s->d_rowptr, s->d_colind, s->d_val, x, ldx, b, ldb, r, berr, safe1, safe2)
    const bool ex = s->opt.residual_extra != 0;
    if (nr <= 1) { if (ex) NKP_RES(1, true); else NKP_RES(1, false); }
    else if (nr == 2) { if (ex) NKP_RES(2, true); else NKP_RES(2, false); }
    else if (nr <= 4) { if (ex) NKP_RES(4, true); else NKP_RES(4, false); }
    else { if (ex) NKP_RES(8, true); else NKP_RES(8, false); }
#undef NKP_RES
    s->launches++;
}

static int sweeps_nr(nkp_solver* s, int nr) {
    switch (nr) {
        case 1: return sweeps<1>(s);
        case 2: return sweeps<2>(s);
        case 3:
        case 4: return sweeps<4>(s);
        default: return sweeps<8>(s);
    }
}
static int padded_nr(int nr) { return nr <= 2 ? nr : (nr <= 4 ? 4 : 8); }

// solve one chunk of nr <= MAX_NR right-hand sides held in device memory (in place)
static int solve_chunk(nkp_solver* s, double* dB, int ldb, int nr, double* berr_host, int* steps_out) {
    const int n = s->n;
    cudaStream_t st = s->stream;
    const int nrp = padded_nr(nr);
    const int g = (n + 255) / 256;
    const double eps = 2.220446049250313e-16;
    // pdgsrfs: safe1 = nz * safmin (nz = A->ncol + 1 in the library; any small multiple of safmin serves), safe2 = safe1 / eps
    const double safe1 = 2.2250738585072014e-308 * (double)(s->nnz / n + 2);
    const double safe2 = safe1 / eps;
    // x = 0-th solve
    if (nrp > nr) CK(cudaMemsetAsync(s->d_y, 0, sizeof(double) * (size_t)n * nrp, st));
    k_permute_in<<<g, 256, 0, st>>>(n, nr, s->d_permr, s->d_R, dB, ldb, s->d_y);
    check_finite(s, "the right-hand side", s->d_y, (int64_t)n * nr);
    if (sweeps_nr(s, nrp)) return NKP_ECUDA;
    check_finite(s, "the first sweep pair's result", s->d_y, (int64_t)n * nr);
    k_permute_out<<<g, 256, 0, st>>>(n, nr, s->d_perm, s->d_C, s->d_y, s->d_x, n, 0, 0xffu);
    s->launches += 2;
    double last[MAX_NR];
    for (int c = 0; c < nr; c++) last[c] = 3.0;   // lstres of pdgsrfs
    double berr[MAX_NR] = {0};
    bool done[MAX_NR] = {false};
    const bool normwise = s->opt.refine_rule == 1;
    double bnorm2[MAX_NR] = {0}, rnorm2[MAX_NR] = {0};
    if (normwise) {
        k_sumsq<<<SUMSQ_BLOCKS, 256, 0, st>>>(n, nr, dB, ldb, s->d_sumsq);
        k_sumsq_final<<<1, 32, 0, st>>>(SUMSQ_BLOCKS, nr, s->d_sumsq, s->d_berr + MAX_NR);
        s->launches += 2;
        CK(cudaMemcpyAsync(bnorm2, s->d_berr + MAX_NR, sizeof(double) * nr, cudaMemcpyDeviceToHost, st));
    }
    int it = 0;
    for (;;) {
        // r = b - A x, berr
        CK(cudaMemsetAsync(s->d_berr, 0, sizeof(double) * MAX_NR, st));
        launch_residual(s, nr, s->d_x, n, dB, ldb, s->d_r, s->d_berr, safe1, safe2);
        CK(cudaMemcpyAsync(berr, s->d_berr, sizeof(double) * nr, cudaMemcpyDeviceToHost, st));
        if (normwise) {
            k_sumsq<<<SUMSQ_BLOCKS, 256, 0, st>>>(n, nr, s->d_r, n, s->d_sumsq);
            k_sumsq_final<<<1, 32, 0, st>>>(SUMSQ_BLOCKS, nr, s->d_sumsq, s->d_berr + 2 * MAX_NR);
            s->launches += 2;
            CK(cudaMemcpyAsync(rnorm2, s->d_berr + 2 * MAX_NR, sizeof(double) * nr, cudaMemcpyDeviceToHost, st));
        }
        CK(cudaStreamSynchronize(st));
        if (s->opt.verbose > 1) {
            fprintf(stderr, "[nkp] refine it %d berr:", it);
            for (int c = 0; c < nr; c++) fprintf(stderr, " %.2e", berr[c]);
            fprintf(stderr, "\n");
        }
        bool go = false;
        for (int c = 0; c < nr; c++) {
            // SuperLU pdgsrfs, per right-hand side: continue while berr > eps and berr decreased by
            // at least a factor 2; a column that stops once stays stopped
            // refine_rule 1 applies the same logic to the normwise relative residual with threshold 1e-14
            const double crit = normwise ? (bnorm2[c] > 0 ? std::sqrt(rnorm2[c] / bnorm2[c]) : 0.0) : berr[c];
            const double tol = normwise ? 1e-14 : eps;
            if (!done[c] && crit > tol && crit * 2.0 <= last[c]) go = true;
            else done[c] = true;
            last[c] = crit;
        }
        if (!go || it >= s->opt.refine_max) break;
        it++;
        k_permute_in<<<g, 256, 0, st>>>(n, nr, s->d_permr, s->d_R, s->d_r, n, s->d_y);
        check_finite(s, "a residual", s->d_y, (int64_t)n * nr);
        if (sweeps_nr(s, nrp)) return NKP_ECUDA;
        check_finite(s, "a correction sweep's result", s->d_y, (int64_t)n * nr);
        unsigned active = 0;
        for (int c = 0; c < nr; c++)
            if (!done[c]) active |= 1u << c;
        k_permute_out<<<g, 256, 0, st>>>(n, nr, s->d_perm, s->d_C, s->d_y, s->d_x, n, 1, active);
        s->launches += 2;
    }
    // write the solution over B
    CK(cudaMemcpy2DAsync(dB, sizeof(double) * ldb, s->d_x, sizeof(double) * n, sizeof(double) * n, nr,
                         cudaMemcpyDeviceToDevice, st));
    if (berr_host)
        for (int c = 0; c < nr; c++) berr_host[c] = berr[c];
    *steps_out = it;
    return 0;
}

int nkp_solve_device(nkp_solver* s, double* dB, int ldb, int nrhs, double* berr) {
    if (!s || !dB || ldb < s->n || nrhs < 0) {
        g_err = "nkp_solve_device: invalid argument";
        return NKP_EINVAL;
    }
    if (!s->factored) {
        g_err = "nkp_solve: matrix not factored";
        return NKP_ESTATE;
    }
    CK(cudaSetDevice(s->opt.device));
    CK(cudaEventRecord(s->ev[0], s->stream));
    int maxsteps = 0;
    for (int c0 = 0; c0 < nrhs; c0 += MAX_NR) {
        int nr = std::min(MAX_NR, nrhs - c0);
        int steps = 0;
        int rc = solve_chunk(s, dB + (size_t)c0 * ldb, ldb, nr, berr ? berr + c0 : nullptr, &steps);
        if (rc) return rc;
        maxsteps = std::max(maxsteps, steps);
    }
    CK(cudaEventRecord(s->ev[1], s->stream));
    CK(cudaStreamSynchronize(s->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]));
    s->t_solve = ms * 1e-3;
    s->refine_steps = maxsteps;
    return NKP_OK;
}

int nkp_solve(nkp_solver* s, double* B, int ldb, int nrhs, double* berr) {
    if (!s || !B || ldb < s->n || nrhs < 0) {
        g_err = "nkp_solve: invalid argument";
        return NKP_EINVAL;
    }
    if (!s->factored) {
        g_err = "nkp_solve: matrix not factored";
        return NKP_ESTATE;
    }
    CK(cudaSetDevice(s->opt.device));
    const int n = s->n;
    int maxsteps = 0;
    double tsum = 0;
    for (int c0 = 0; c0 < nrhs; c0 += MAX_NR) {
        int nr = std::min(MAX_NR, nrhs - c0);
        // column by column, so that the DMA of one column overlaps the host copy of the next
        const bool pinned = host_is_pinned(B + (size_t)c0 * ldb);
        for (int c = 0; c < nr; c++) {
            const double* src = B + (size_t)(c0 + c) * ldb;
            if (!pinned) {
                par_memcpy(s->h_pinned + (size_t)c * n, src, sizeof(double) * n);
                src = s->h_pinned + (size_t)c * n;
            }
            CK(cudaMemcpyAsync(s->d_xb + (size_t)c * n, src, sizeof(double) * n, cudaMemcpyHostToDevice, s->stream));
        }
        int rc = nkp_solve_device(s, s->d_xb, n, nr, berr ? berr + c0 : nullptr);
        if (rc) return rc;
        tsum += s->t_solve;
        maxsteps = std::max(maxsteps, s->refine_steps);
        for (int c = 0; c < nr; c++) {
            double* dst = pinned ? B + (size_t)(c0 + c) * ldb : s->h_pinned + (size_t)c * n;
            CK(cudaMemcpyAsync(dst, s->d_xb + (size_t)c * n, sizeof(double) * n, cudaMemcpyDeviceToHost, s->stream));
            CK(cudaEventRecord(s->ev_col[c], s->stream));
        }
        for (int c = 0; c < nr; c++) {
            CK(cudaEventSynchronize(s->ev_col[c]));
            if (!pinned) par_memcpy(B + (size_t)(c0 + c) * ldb, s->h_pinned + (size_t)c * n, sizeof(double) * n);
        }
    }
    s->t_solve = tsum;
    s->refine_steps = maxsteps;
    return NKP_OK;
}

int nkp_solve_dist(nkp_solver* s, double* B, int ldb, int nrhs, int fst_row, int m_loc, double* berr) {
    if (!s || !B || nrhs < 0 || fst_row < 0 || m_loc < 0 || fst_row + (int64_t)m_loc > s->n || ldb < m_loc) {
        g_err = "nkp_solve_dist: invalid argument";
        return NKP_EINVAL;
    }
    if (s->nranks == 1) {
        if (fst_row != 0 || m_loc != s->n) {
            g_err = "nkp_solve_dist: one rank must hold all rows";
            return NKP_EINVAL;
        }
        return nkp_solve(s, B, ldb, nrhs, berr);
    }
    if (!s->factored) {
        g_err = "nkp_solve: matrix not factored";
        return NKP_ESTATE;
    }
    CK(cudaSetDevice(s->opt.device));
    const int n = s->n, P = s->nranks;
    cudaStream_t st = s->stream;
    // who holds which rows (slab table of all ranks)
    if (!s->d_slab) {
        CK(cudaMalloc((void**)&s->d_slab, sizeof(PubRange) * (size_t)(P + 1)));
        s->h_slab.assign(P, PubRange{0, 0, 0});
    }
    const PubRange mine{fst_row, fst_row + m_loc, s->rank};
    CK(cudaMemcpyAsync(s->d_slab + P, &mine, sizeof(PubRange), cudaMemcpyHostToDevice, st));
    CKN(ncclAllGather(s->d_slab + P, s->d_slab, sizeof(PubRange), ncclChar, s->comm, st));
    CK(cudaMemcpyAsync(s->h_slab.data(), s->d_slab, sizeof(PubRange) * (size_t)P, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    {
        std::vector<PubRange> t = s->h_slab;
        std::sort(t.begin(), t.end(), [](const PubRange& a, const PubRange& b) { return a.lo < b.lo || (a.lo == b.lo && a.hi < b.hi); });
        int pos = 0;
        for (const PubRange& r : t) {
            if (r.lo != pos || r.hi < r.lo) {
                g_err = "nkp_solve_dist: the row slabs of the ranks do not tile [0, n)";
                return NKP_EINVAL;
            }
            pos = r.hi;
        }
        if (pos != n) {
            g_err = "nkp_solve_dist: the row slabs of the ranks do not tile [0, n)";
            return NKP_EINVAL;
        }
    }
    int maxsteps = 0;
    double tsum = 0;
    const bool pinned = host_is_pinned(B);
    for (int c0 = 0; c0 < nrhs; c0 += MAX_NR) {
        const int nr = std::min(MAX_NR, nrhs - c0);
        for (int c = 0; c < nr; c++) {
            const double* src = B + (size_t)(c0 + c) * ldb;
            if (!pinned) {
                par_memcpy(s->h_pinned + (size_t)c * m_loc, src, sizeof(double) * m_loc);
                src = s->h_pinned + (size_t)c * m_loc;
            }
            CK(cudaMemcpyAsync(s->d_xb + (size_t)c * n + fst_row, src, sizeof(double) * m_loc, cudaMemcpyHostToDevice, st));
        }
        // slabs -> every GPU: packed per slab, one broadcast per rank, one group call (NVLink instead of P x PCIe)
        k_pub_pack<<<dim3(64, P), 256, 0, st>>>(s->d_slab, s->rank, 0, s->d_xb, n, nr, s->d_pack);
        CKN(ncclGroupStart());
        for (const PubRange& r : s->h_slab) {
            if (r.hi <= r.lo) continue;
            double* p = s->d_pack + (size_t)r.lo * nr;
            CKN(ncclBroadcast(p, p, (size_t)(r.hi - r.lo) * nr, ncclDouble, r.root, s->comm, st));
        }
        CKN(ncclGroupEnd());
        k_pub_pack<<<dim3(64, P), 256, 0, st>>>(s->d_slab, s->rank, 1, s->d_xb, n, nr, s->d_pack);
        s->launches += 2;
        int rc = nkp_solve_device(s, s->d_xb, n, nr, berr ? berr + c0 : nullptr);
        if (rc) return rc;
        tsum += s->t_solve;
        maxsteps = std::max(maxsteps, s->refine_steps);
        for (int c = 0; c < nr; c++) {
            double* dst = pinned ? B + (size_t)(c0 + c) * ldb : s->h_pinned + (size_t)c * m_loc;
            CK(cudaMemcpyAsync(dst, s->d_xb + (size_t)c * n + fst_row, sizeof(double) * m_loc, cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(s->ev_col[c], st));
        }
        for (int c = 0; c < nr; c++) {
            CK(cudaEventSynchronize(s->ev_col[c]));
            if (!pinned) par_memcpy(B + (size_t)(c0 + c) * ldb, s->h_pinned + (size_t)c * m_loc, sizeof(double) * m_loc);
        }
    }
    s->t_solve = tsum;
    s->refine_steps = maxsteps;
    return NKP_OK;
}

int nkp_set_tracer_maps(nkp_solver* s, int tracer_state_len, int coupled_tracer_cnt, const int* ind_i, const int* ind_j,
                        const int* ind_k, int imt, int jmt, int km) {
    if (!s || tracer_state_len <= 0 || coupled_tracer_cnt <= 0 || !ind_i || !ind_j || !ind_k || imt <= 0 || jmt <= 0 ||
        km <= 0 || (int64_t)tracer_state_len * coupled_tracer_cnt != s->n) {
        g_err = "nkp_set_tracer_maps: invalid argument (n must equal coupled_tracer_cnt * tracer_state_len)";
        return NKP_EINVAL;
    }
    CK(cudaSetDevice(s->opt.device));
    std::vector<int> cell((size_t)tracer_state_len);
    for (int q = 0; q < tracer_state_len; q++) {
        if (ind_i[q] < 0 || ind_i[q] >= imt || ind_j[q] < 0 || ind_j[q] >= jmt || ind_k[q] < 0 || ind_k[q] >= km) {
            g_err = "nkp_set_tracer_maps: index map entry out of range";
            return NKP_EINVAL;
        }
        cell[q] = (ind_k[q] * jmt + ind_j[q]) * imt + ind_i[q];
    }
    if (s->d_cell) cudaFree(s->d_cell);
    if (s->d_fields) cudaFree(s->d_fields);
    if (s->h_field) cudaFreeHost(s->h_field);
    s->d_cell = nullptr;
    s->d_fields = nullptr;
    s->h_field = nullptr;
    s->tsl = tracer_state_len;
    s->ct = coupled_tracer_cnt;
    s->ncell = (int64_t)imt * jmt * km;
    if (upload(&s->d_cell, cell)) return NKP_ECUDA;
    CK(cudaMalloc((void**)&s->d_fields, sizeof(double) * (size_t)s->ncell * MAX_NR * s->ct));
    CK(cudaMallocHost((void**)&s->h_field, sizeof(double) * (size_t)s->ncell * MAX_NR * s->ct));
    return NKP_OK;
}

int nkp_solve_fields(nkp_solver* s, double* const* fields, int nfields, double* berr) {
    if (!s || !fields || nfields < 0) {
        g_err = "nkp_solve_fields: invalid argument";
        return NKP_EINVAL;
    }
    if (!s->d_cell) {
        g_err = "nkp_solve_fields: call nkp_set_tracer_maps first";
        return NKP_ESTATE;
    }
    if (!s->factored) {
        g_err = "nkp_solve: matrix not factored";
        return NKP_ESTATE;
    }
    if (nfields % s->ct != 0) {
        // the reference aborts when the variable list runs out inside a group (src/solve_ABglobal.c:376-379)
        g_err = "nkp_solve_fields: number of fields is not a multiple of coupled_tracer_cnt";
        return NKP_EINVAL;
    }
    CK(cudaSetDevice(s->opt.device));
    cudaStream_t st = s->stream;
    const int n = s->n, ct = s->ct, tsl = s->tsl;
    const size_t fbytes = sizeof(double) * (size_t)s->ncell;
    const int ngroups = nfields / ct;
    int maxsteps = 0;
    double tsum = 0;
    for (int g0 = 0; g0 < ngroups; g0 += MAX_NR) {
        const int ng = std::min(MAX_NR, ngroups - g0);
        const int nf = ng * ct;
        // one pinned staging slot per field: no synchronisation between the copies
        for (int f = 0; f < nf; f++) {
            const double* src = fields[g0 * ct + f];
            if (!host_is_pinned(src)) {
                par_memcpy(s->h_field + (size_t)f * s->ncell, src, fbytes);
                src = s->h_field + (size_t)f * s->ncell;
            }
            CK(cudaMemcpyAsync(s->d_fields + (size_t)f * s->ncell, src, fbytes, cudaMemcpyHostToDevice, st));
        }
        k_gather_fields<<<(tsl + 255) / 256, 256, 0, st>>>(tsl, ct, nf, s->d_cell, s->d_fields, s->ncell, s->d_xb, n);
        s->launches++;
        int rc = nkp_solve_device(s, s->d_xb, n, ng, berr ? berr + g0 : nullptr);
        if (rc) return rc;
        tsum += s->t_solve;
        maxsteps = std::max(maxsteps, s->refine_steps);
        k_scatter_fields<<<(tsl + 255) / 256, 256, 0, st>>>(tsl, ct, nf, s->d_cell, s->d_fields, s->ncell, s->d_xb, n);
        s->launches++;
        CK(cudaGetLastError());
        if ((int)s->ev_field.size() < nf) {
            const size_t old = s->ev_field.size();
            s->ev_field.resize(nf, nullptr);
            for (size_t q = old; q < s->ev_field.size(); q++) CK(cudaEventCreateWithFlags(&s->ev_field[q], cudaEventDisableTiming));
        }
        for (int f = 0; f < nf; f++) {
            double* dst = fields[g0 * ct + f];
            if (!host_is_pinned(dst)) dst = s->h_field + (size_t)f * s->ncell;
            CK(cudaMemcpyAsync(dst, s->d_fields + (size_t)f * s->ncell, fbytes, cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(s->ev_field[f], st));
        }
        for (int f = 0; f < nf; f++) {
            CK(cudaEventSynchronize(s->ev_field[f]));
            double* dst = fields[g0 * ct + f];
            if (!host_is_pinned(dst)) par_memcpy(dst, s->h_field + (size_t)f * s->ncell, fbytes);
        }
    }
    s->t_solve = tsum;
    s->refine_steps = maxsteps;
    return NKP_OK;
}

int nkp_residual_device(nkp_solver* s, const double* dx, const double* db, double* dr, int nrhs) {
    if (!s || !dx || !db || !dr || nrhs < 0) return NKP_EINVAL;
    if (nrhs == 0) return NKP_OK;
    CK(cudaSetDevice(s->opt.device));
    const int n = s->n;
    for (int c0 = 0; c0 < nrhs; c0 += MAX_NR)
        launch_residual(s, std::min(MAX_NR, nrhs - c0), dx + (size_t)c0 * n, n, db + (size_t)c0 * n, n, dr + (size_t)c0 * n, nullptr, 0.0, 0.0);
    CK(cudaGetLastError());
    return NKP_OK;
}

int nkp_sweeps_device(nkp_solver* s, double* dB, int ldb, int nrhs) {
    if (!s || !dB || nrhs < 1 || nrhs > MAX_NR) return NKP_EINVAL;
    if (!s->factored) return NKP_ESTATE;
    CK(cudaSetDevice(s->opt.device));
    const int n = s->n;
    const int nrp = padded_nr(nrhs);
    const int g = (n + 255) / 256;
    CK(cudaEventRecord(s->ev[2], s->stream));
    if (nrp > nrhs) CK(cudaMemsetAsync(s->d_y, 0, sizeof(double) * (size_t)n * nrp, s->stream));
    k_permute_in<<<g, 256, 0, s->stream>>>(n, nrhs, s->d_permr, s->d_R, dB, ldb, s->d_y);
    if (sweeps_nr(s, nrp)) return NKP_ECUDA;
    k_permute_out<<<g, 256, 0, s->stream>>>(n, nrhs, s->d_perm, s->d_C, s->d_y, dB, ldb, 0, 0xffu);
    s->launches += 2;
    CK(cudaEventRecord(s->ev[3], s->stream));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, s->ev[2], s->ev[3]));
    s->t_sweeps = ms * 1e-3;
    return NKP_OK;
}

int nkp_get_perm(const nkp_solver* s, int* perm) {
    if (!s || !perm) return NKP_EINVAL;
    memcpy(perm, s->plan.perm.data(), sizeof(int) * s->n);
    return NKP_OK;
}

int nkp_get_stats(const nkp_solver* s, nkp_stats* st) {
    if (!s || !st) return NKP_EINVAL;
    memset(st, 0, sizeof(*st));
    const Plan& P = s->plan;
    st->n = s->n;
    st->nnz = s->nnz;
    st->n_fronts = (int)P.fronts.size();
    st->n_levels = P.nlevels;
    st->max_front = P.max_front;
    st->nnz_lu = P.nnz_lu;
    st->factor_flops = P.flops;
    st->heap_bytes = 8.0 * (double)P.heap_len;
    st->t_analysis = s->t_analysis;
    st->t_factor = s->t_factor;
    st->t_scatter = s->t_scatter;
    st->t_solve = s->t_solve;
    st->refine_steps = s->refine_steps;
    st->tiny_pivots = s->tiny_pivots;
    st->kernel_launches = s->launches;
    // 8 (nnz(L)+nnz(U)) + index bytes + 2*8*n per sweep pair
    st->solve_bytes = 8.0 * (double)P.nnz_lu + 4.0 * (double)(P.bidx.size() + P.rel.size()) + 16.0 * s->n;
    st->t_gemm = s->t_cls[KC_GEMM];
    st->gemm_flops = P.gemm_flops;
    st->n_gemm = s->n_cls[KC_GEMM];
    st->t_trsm = s->t_cls[KC_TRSM];
    st->t_diag = s->t_cls[KC_DIAG];
    st->t_extend_add = s->t_cls[KC_ADD];
    st->t_sweeps = s->t_sweeps;
    st->factor_flops_local = P.flops_local;
    st->nnz_lu_local = (double)P.nnz_lu_local;
    st->n_xfers = (double)P.xfers.size();
    st->order_cached = P.order_cached ? 1.0 : 0.0;
    return NKP_OK;
}

int nkp_set_profile(nkp_solver* s, int on) {
    if (!s) return NKP_EINVAL;
    s->prof_on = on != 0;
    return NKP_OK;
}

int nkp_set_refine_rule(nkp_solver* s, int rule) {
    if (!s || rule < 0 || rule > 1) return NKP_EINVAL;
    s->opt.refine_rule = rule;
    return NKP_OK;
}

int nkp_set_verbose(nkp_solver* s, int level) {
    if (!s) return NKP_EINVAL;
    s->opt.verbose = level;
    return NKP_OK;
}

int nkp_set_residual_extra(nkp_solver* s, int on) {
    if (!s) return NKP_EINVAL;
    s->opt.residual_extra = on != 0;
    return NKP_OK;
}

int nkp_sync(nkp_solver* s) {
    if (!s) return NKP_EINVAL;
    CK(cudaStreamSynchronize(s->stream));
    return NKP_OK;
}

void nkp_destroy(nkp_solver* s) {
    if (!s) return;
    cudaSetDevice(s->opt.device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->cstream) cudaStreamSynchronize(s->cstream);
    for (ncclComm_t c : s->gcomm)
        if (c && c != s->comm) ncclCommDestroy(c);
    if (s->comm) ncclCommDestroy(s->comm);
    if (s->cstream) cudaStreamDestroy(s->cstream);
    if (s->ev_p) cudaEventDestroy(s->ev_p);
    if (s->ev_b) cudaEventDestroy(s->ev_b);
    if (s->d_permr && s->d_permr != s->d_perm) cudaFree(s->d_permr);
    if (s->d_pub) cudaFree(s->d_pub);
    if (s->d_slab) cudaFree(s->d_slab);
    if (s->d_pack) cudaFree(s->d_pack);
    void* ptrs[] = {s->heap,   s->d_rowptr, s->d_colind, s->d_rowidx, s->d_val,  s->d_scatter, s->d_perm,
                    s->d_bidx, s->d_rel,    s->d_R,      s->d_C,      s->d_diag, s->d_trsm,    s->d_gemm,
                    s->d_add,  s->d_solve,  s->d_children, s->d_W,    s->d_y,    s->d_r,       s->d_x,
                    s->d_xb,   s->d_berr,   s->d_nrepl,  s->d_small,  s->d_big,  s->d_fwd_items, s->d_bwd_items,
                    s->d_cnt,  s->d_inv,    s->d_rect_items, s->d_part, s->d_clo, s->d_sumsq};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (s->h_pinned) cudaFreeHost(s->h_pinned);
    if (s->h_field) cudaFreeHost(s->h_field);
    if (s->d_cell) cudaFree(s->d_cell);
    if (s->d_fields) cudaFree(s->d_fields);
    for (int i = 0; i < 4; i++)
        if (s->ev[i]) cudaEventDestroy(s->ev[i]);
    for (int i = 0; i < MAX_NR; i++)
        if (s->ev_col[i]) cudaEventDestroy(s->ev_col[i]);
    for (cudaEvent_t e : s->prof_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : s->trace_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : s->ev_field)
        if (e) cudaEventDestroy(e);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

