// nkp_internal.hpp -- shared definitions between the host analysis (analysis.cpp),
// the CUDA numeric phase (solver.cu / kernels.cuh) and the CPU plan simulator used by
// the tests (oracle/plan_sim.cpp).
//
// The solver replaces what the reference obtains from SuperLU_DIST's pdgssvx*
// (src/solve_ABglobal.c:353,395 ; src/solve_ABdist.c:518,571): analysis on the host
// (ordering, assembly tree, symbolic fronts, memory plan, task lists), numeric
// multifrontal LU + triangular solves + refinement on the GPU.
//
// Storage layout of one front t (s pivots, r boundary rows/cols, m = s + r):
//   Larr  : m x s column-major (ld = m rounded up to even, so every column starts 16-byte aligned).  Holds every entry (a,b) of the front with
//           b < s and (a >= s or blk(a) >= blk(b)), i.e. the block-lower part
//           including the full nb x nb diagonal blocks.  After factorisation:
//           L (unit lower) below the diagonal blocks, packed LU in the diagonal blocks.
//   UTarr : m x s column-major (same ld).  Holds U TRANSPOSED: entry (a,b) of the front
//           with a < s and (b >= s or blk(b) > blk(a)) is stored at UTarr[b + a*m].
//           After factorisation the diagonal blocks additionally receive U_kk^T, so the
//           backward solve reads UTarr only.
//   F22   : r x r column-major (ld = r), the update (Schur complement) matrix, lives in a
//           per-level ping-pong pool and dies once the parent has been assembled.
// Keeping U transposed makes every panel a tall column-major array: the L and U panel
// solves are one kernel (X * T = B, T upper triangular) and every Schur update is the
// same "C -= A * B^T" GEMM with coalesced operand loads.
#pragma once
#include <cstdint>
#include <vector>

namespace nkp {

struct Front {
    int first = 0;       // first pivot (permuted index)
    int s = 0;           // number of pivots
    int r = 0;           // boundary size
    int m = 0;           // s + r
    int ld = 0;          // leading dimension of Larr / UTarr: m rounded up to even (16-byte aligned columns)
    int parent = -1;
    int level = 0;       // depth from the root (roots are level 0)
    int child_rank = 0;  // position among the parent's children (extend-add pass)
    int nchild = 0;
    int64_t Loff = 0;    // offsets in doubles into the device heap
    int64_t UToff = 0;
    int64_t F22off = 0;
    int64_t bidx_off = 0;  // into Plan::bidx (r entries: permuted global index of each boundary row)
    int64_t rel_off = 0;   // into Plan::rel  (r entries: local index in the parent's front)
    int64_t woff = 0;      // solve work vector offset (m entries per rhs) in the solve pool
};

// ---- device task descriptors (plain data, uploaded once per analysis) ----

struct DiagTask {   // factor one kb x kb diagonal block in place (no pivoting, tiny-pivot replacement)
    int64_t Doff;   // Larr diagonal block
    int64_t UTDoff; // UTarr diagonal block (receives U_kk^T)
    int ld;
    int kb;
};

struct TrsmTask {   // X * T = B in place on a tall panel: rows [0,nrows) x kb columns
    int64_t Xoff;
    int64_t Toff;   // factored diagonal block (in Larr)
    int ld;         // leading dimension of X and of T's array
    int nrows;
    int kb;
    int unit;       // 0: T = U_kk (upper, non-unit)   1: T = L_kk^T (upper, unit)
    int cta0;       // first CTA of this task in the launch
    int pad;
};

struct GemmTask {   // C[M x N] -= A[M x K] * B[N x K]^T, all column-major, K <= outer*nb
    int64_t Aoff, Boff, Coff;
    int M, N, K;
    int lda, ldb, ldc;
    int skip;       // tile-skip rule, local coords (row a, col b), bj = b / nb:
                    //   0 none
                    //   1 block-lower target (Larr):  rows a <  bj*nb            are unused
                    //   2 block-upper target (UTarr): rows a <  min((bj+1)*nb,N) are unused
                    // a tile is skipped when all its rows are unused; partially unused tiles are
                    // computed in full (the unused part of the arrays is scratch)
    int tile0;      // first tile of this task in the launch
    int tiles_m;    // number of tile rows
    int pad;
};

struct AddTask {    // extend-add a child's update matrix into its parent front
    int64_t Coff;   // child F22 (rc x rc, ld = rc)
    int64_t rel_off;
    int64_t Loff, UToff, F22off;  // parent arrays
    int rc;
    int sp, mp;     // parent pivots / size
    int tile0;
    int tiles_m;
    int ldp;        // leading dimension of the parent's Larr / UTarr
};

struct SolveTask {  // one front in a forward / backward sweep
    int64_t Loff, UToff;
    int64_t bidx_off;
    int64_t rel_off;
    int64_t woff;         // this front's work vector (m per rhs)
    int64_t child_list;   // offset into Plan::solve_children
    int first, s, r, m;
    int nchild;
    int ld;               // leading dimension of Larr / UTarr
    int big;              // 1: swept by the multi-CTA dataflow kernels (diagonal blocks are stored inverted)
};

struct SolveChild {
    int64_t woff;     // child's work vector
    int64_t rel_off;
    int s, r;
};

// big fronts are swept by many CTAs with a counter-driven dataflow (k_sweep_big)
struct BigFront {
    int64_t Loff, UToff, bidx_off, woff;
    int first, s, r, m;
    int npiv;    // ceil(s / 64) pivot blocks
    int nslab;   // npiv + ceil(r / 64) row slabs
    int pad0;
    int nchild;
    int64_t child_list;  // into Plan::solve_children
    int64_t part_off;    // backward sweep: first 64 x 8 partial-product slot of this front (level scratch)
    int nchunk;          // backward sweep: chunks of BWD_CHUNK boundary row blocks per pivot panel
    int ld;              // leading dimension of Larr / UTarr (even)
    int64_t clo_off;     // forward sweep: into Plan::child_lo, [child * nslab + slab] = first entry of the child's
                         // rel[] list that maps to a row >= the slab's first row
};

constexpr int BWD_CHUNK = 16;   // 64-row blocks of the boundary per item of the rectangular part

struct BigItem {
    int front;   // index into Plan::big_fronts
    int idx;     // row slab (forward) or pivot panel (backward)
};

// multi-GPU, forward sweep: the update vector of `front` (a child of a top front) travels from a rank
// that has swept it to a member of the parent's group that has not
struct Xfer {
    int front;
    int src, dst;
    int level;   // level of `front` (the child)
};

// ---- multi-GPU: distributed factorisation of the top separator fronts --------------------------------
// A top front (an ancestor of the rank-private subtrees) is factored by its GROUP: the ranks that own
// something below it.  Every member keeps a full copy of the front's arrays; the pivot columns are cut
// into outer blocks of W = outer * nb columns, block K (its Larr and UTarr column blocks) belongs to member
// K mod g, the column blocks of the update matrix continue the same cycle.  Per block: the owner factors
// the panel, both column blocks are broadcast over the group's communicator, every member applies the
// wide Schur update to the blocks it owns -- the block that is factored next first (look-ahead), so that
// its broadcast overlaps everybody's remaining update.  After the last block every member holds the
// complete factors of the front, and the sweeps of the top fronts need no communication at all.
struct TopBcast {       // one ncclBroadcast over the group of the receiving front
    int root;           // world rank that holds the data
    int pad;
    int64_t off;        // heap offset on THIS rank (send buffer on the root, receive buffer elsewhere)
    int64_t count;      // doubles
};

struct TopStep {        // one inner panel (nb columns) of the panel factorisation, owner only
    int diag_begin = 0, diag_end = 0;
    int trsm_begin = 0, trsm_end = 0, trsm_ctas = 0;
    int gemm_begin = 0, gemm_end = 0, gemm_tiles = 0;   // narrow update of the rest of the outer block
};

struct TopBlock {       // one outer block of a top front
    int owner = 0;      // world rank
    int step_begin = 0, step_end = 0;                       // Plan::top_steps (empty unless this rank is the owner)
    int next_begin = 0, next_end = 0, next_tiles = 0;       // GemmTasks: wide update of block K + 1 (its owner only)
    int rest_begin = 0, rest_end = 0, rest_tiles = 0;       // GemmTasks: wide update of the other blocks this rank owns
    TopBcast bl, bu;                                        // the factored Larr / UTarr column blocks
};

struct TopFront {
    int front = 0;
    int group = 0;              // index into Plan::groups
    int member = 0;             // this rank belongs to the group (and stores the front)
    int block_begin = 0, block_end = 0;   // Plan::top_blocks
    int cb_begin = 0, cb_end = 0;         // Plan::top_child_bcasts: update matrices of the children
    std::vector<int> add_begin, add_tiles;   // extend-add passes (one per child), ranges into Plan::add_tasks
    int inv_begin = 0, inv_end = 0;       // Plan::inv_tasks: diagonal blocks inverted after the factorisation
    int64_t f22_off = 0, f22_len = 0;     // this front's own update matrix (cleared before the extend-add)
};

// solution ranges published after the backward sweep: [lo, hi) of the permuted numbering, held by `root`
struct PubRange {
    int lo, hi, root;
};

struct LevelPlan {
    int level = 0;
    std::vector<int> fronts;             // fronts on this level (all ranks)
    std::vector<int> mine;               // ... rank-private (non-top) fronts owned by this rank: batched factor launches
    std::vector<int> stored;             // ... every front this rank keeps factors of (mine + its top fronts): sweeps
    std::vector<int> ghosts;             // ... not stored, but children of fronts stored by this rank
    std::vector<int> xfers;              // indices into Plan::xfers whose child lives on this level
    std::vector<int> tops;               // indices into Plan::top_fronts of the top fronts on this level (all ranks)
    int nsteps = 0;                      // ceil(max s / nb)
    // per step task ranges into the flat arrays below
    std::vector<int> diag_begin, trsm_begin, gemm_begin;  // size nsteps+1
    std::vector<int> trsm_ctas, gemm_tiles;               // per step launch sizes
    // extend-add passes (children of this level's fronts live on level+1)
    std::vector<int> add_begin;          // size npass+1
    std::vector<int> add_tiles;          // per pass
    int64_t f22_zero_off = 0, f22_zero_len = 0;  // region of the update pool to clear
    int solve_begin = 0, solve_end = 0;  // into Plan::solve_tasks (all fronts of the level)
    int small_begin = 0, small_end = 0;  // into Plan::solve_small (fronts swept by one CTA each)
    int big_begin = 0, big_end = 0;      // into Plan::big_fronts
    int fwd_item_begin = 0, fwd_item_end = 0;  // into Plan::big_fwd_items
    int bwd_item_begin = 0, bwd_item_end = 0;  // into Plan::big_bwd_items
    int inv_begin = 0, inv_end = 0;            // into Plan::inv_tasks (diagonal blocks of this level's fronts)
    int rect_item_begin = 0, rect_item_end = 0;  // into Plan::big_rect_items (idx = panel * nchunk + chunk)
};

struct Options {
    int nb = 64;           // pivot block width (inner panel)
    int outer = 8;         // inner panels per outer block (wide Schur updates use K = outer*nb)
    int top_outer = 8;     // the same for the distributed top fronts (multi-GPU): width of a distribution block / nb
    int leaf = 96;         // stop dissecting below this many unknowns
    int etree_supernodes = 1;   // fronts = relaxed supernodes of the elimination tree of the dissection ordering (0: the dissection
                                // nodes themselves, one dense front per separator / leaf -- the round-1 assembly tree)
    double relax_frac = 0.10;   // ... a supernode is joined with its parent while the explicit zeros stay below this share
    int relax_small = 32;       // ... or the joined supernode has at most this many columns
    int tm = 128, tn = 64; // GEMM tile (must match the kernel)
    int trsm_rows = 128;   // rows per TRSM CTA
    int add_tile = 32;     // extend-add tile
    int period_i = 0;
    int verbose = 0;
    int64_t big_entries = 1 << 15;  // fronts with m*s >= this use the multi-CTA dataflow sweeps
    int big_rows = 768;             // ... and so do tall fronts (few pivots, long boundary): one warp would crawl
    int rank = 0, nranks = 1;       // multi-GPU: this process' rank (one GPU per rank)
    double split_tol = 0.03;        // multi-GPU: accepted flop imbalance (max / mean - 1) of the two halves of a rank range
    int split_max = 16;             // ... and the largest number of candidate subtrees the splitting may produce per range
};

struct Plan {
    int n = 0;
    int64_t nnz = 0;
    Options opt;
    std::vector<int> perm;     // perm[old] = new
    std::vector<int> iperm;    // iperm[new] = old
    std::vector<Front> fronts; // postorder
    std::vector<int> roots;
    int nlevels = 0;
    std::vector<LevelPlan> levels;   // index = level
    std::vector<int> bidx;           // boundary (permuted) indices, all fronts
    std::vector<int> rel;            // parent-local indices, all fronts
    std::vector<int64_t> scatter;    // per nnz (CRS order): destination offset in the heap
    std::vector<DiagTask> diag_tasks;
    std::vector<TrsmTask> trsm_tasks;
    std::vector<GemmTask> gemm_tasks;
    std::vector<AddTask> add_tasks;
    std::vector<SolveTask> solve_tasks;   // grouped by level (deepest first)
    std::vector<SolveChild> solve_children;
    std::vector<SolveTask> solve_small;   // the non-big subset, grouped by level like solve_tasks
    std::vector<BigFront> big_fronts;
    std::vector<BigItem> big_fwd_items, big_bwd_items, big_rect_items;
    int64_t bwd_part_slots = 0;           // 64 x 8 partial-product slots needed by the largest level
    std::vector<int> child_lo;            // see BigFront::clo_off
    std::vector<DiagTask> inv_tasks;      // diagonal blocks of the big fronts: inverted in place after the factorisation
    int64_t factor_len = 0;    // doubles in the factor arena  [0, factor_len)
    int64_t pool_len[2] = {0, 0};  // update-matrix pools follow the arena
    int64_t pool_off[2] = {0, 0};
    int64_t heap_len = 0;
    int64_t solve_pool_len = 0;    // per rhs
    double flops = 0;          // algorithmic factor flops
    double gemm_flops = 0;     // of which performed by the Schur-update GEMM launches (useful entries only)
    int64_t nnz_lu = 0;        // stored factor entries incl. diagonal
    int max_front = 0;
    double t_order = 0, t_symbolic = 0, t_plan = 0;
    bool order_cached = false;     // ordering + assembly tree came from the on-disk cache
    // multi-GPU partition (identical on every rank)
    int rank = 0, nranks = 1;
    std::vector<int> owner;          // per front
    std::vector<char> is_top;        // per front: part of the shared top of the tree
    std::vector<int> subtree_roots;  // roots of the rank-private subtrees
    std::vector<int> subtree_lo;     // per subtree root: first permuted index of the subtree
    std::vector<Xfer> xfers;         // forward sweep: update vectors that a member of a top front's group lacks
    std::vector<std::vector<int>> groups;   // distinct rank sets of the top fronts (sorted world ranks)
    std::vector<int> group_of;              // per front: index into groups (-1 for rank-private fronts)
    std::vector<TopFront> top_fronts;       // in postorder (children before parents), all ranks
    std::vector<TopBlock> top_blocks;
    std::vector<TopStep> top_steps;
    std::vector<TopBcast> top_child_bcasts;
    std::vector<PubRange> pub;              // who publishes which part of the solution
    double flops_local = 0;          // factor flops of the fronts owned by this rank (top fronts: 1 / group size)
    int64_t nnz_lu_local = 0;        // factor entries stored by this rank (top fronts: full copy on every member)
};

// analysis.cpp
// directory of the on-disk ordering cache; nullptr or "" disables it.  Until this is called the
// NKP_ANALYSIS_CACHE environment variable decides.
void set_analysis_cache_dir(const char* dir);
// coords may be null; otherwise coords[d] (d = 0,1,2) are per-unknown integer coordinates.
// rowmap may be null; otherwise row i of the CRS is row rowmap[i] of the matrix that is factored (static row
// permutation for a large diagonal, rowperm.cpp): Plan::perm then numbers the COLUMNS, row i goes to perm[rowmap[i]].
int analyse(int n, const int* rowptr, const int* colind, const int* const coords[3],
            const Options& opt, Plan& plan, const int* rowmap = nullptr);

inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// where does entry (a,b) of a front live?  returns heap offset
inline int64_t front_entry(int a, int b, int s, int m, int ld, int nb,
                           int64_t Loff, int64_t UToff, int64_t F22off) {
    if (b < s) {
        if (a >= s || a / nb >= b / nb) return Loff + a + (int64_t)b * ld;
        return UToff + b + (int64_t)a * ld;
    }
    if (a < s) return UToff + b + (int64_t)a * ld;
    return F22off + (a - s) + (int64_t)(b - s) * (m - s);
}

}  // namespace nkp
