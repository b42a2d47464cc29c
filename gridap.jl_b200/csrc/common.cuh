// common.cuh -- internal object model of libgridap_b200 (device buffers, context, plan).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/gridap_b200.h"

namespace gb {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

inline std::string fmt(const char *f, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, f);
  vsnprintf(buf, sizeof(buf), f, ap);
  va_end(ap);
  return buf;
}

#define GB_CUDA(expr)                                                                                      \
  do {                                                                                                     \
    cudaError_t e__ = (expr);                                                                              \
    if (e__ != cudaSuccess)                                                                                \
      throw gb::Error(GB200_ERR_CUDA, gb::fmt("CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, \
                                              __LINE__, cudaGetErrorString(e__)));                         \
  } while (0)

#define GB_REQUIRE(cond, code, ...) \
  do {                              \
    if (!(cond)) throw gb::Error(code, gb::fmt(__VA_ARGS__)); \
  } while (0)

// Device memory goes through a per-stream caching allocator (api.cu): freed blocks are kept and handed out again to later
// requests of a similar size on the SAME stream (stream order makes the reuse safe without synchronising), so repeated
// assemblies do not pay cudaMalloc / cudaFree for the multi-GB transients of the symbolic phase (milliseconds each, and
// cudaFree synchronises the device).  Every API call sets the stream of its context here (guarded() in api.cu).
// gb200_trim / an out-of-memory cudaMalloc return the cached blocks to the driver.
inline thread_local cudaStream_t g_alloc_stream = nullptr;
void *dev_alloc(size_t bytes);   // throws gb::Error(GB200_ERR_CUDA) when the device is out of memory
void dev_free(void *p) noexcept;
void dev_cache_trim(cudaStream_t stream_or_null);

// Owning device array.
template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf &operator=(DevBuf &&o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p) dev_free(p);
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count) {
    release();
    n = count;
    if (count) p = static_cast<T *>(dev_alloc(count * sizeof(T)));
  }
  void upload(const T *h, size_t count, cudaStream_t s) {
    if (n != count) alloc(count);
    if (count) GB_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void download(T *h, cudaStream_t s) const {
    if (n) GB_CUDA(cudaMemcpyAsync(h, p, n * sizeof(T), cudaMemcpyDeviceToHost, s));
  }
  void zero(cudaStream_t s) {
    if (n) GB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
  }
  size_t bytes() const { return n * sizeof(T); }
};

struct Timing {
  std::string name;
  float ms;
};
struct PendingTiming {
  std::string name;
  cudaEvent_t a, b;
};

}  // namespace gb

struct gb200_ctx_s {
  int device = 0;
  uint32_t flags = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // D2H of the pattern overlapped with the numeric phase (gb200_plan_get_pattern_async)
  bool copy_pending = false;
  std::vector<void *> copy_keep;       // device staging blocks of the pending copy (returned to the cache once it completed)
  // pattern download with the Int32 -> Int64 widening of the row indices on the HOST (half the PCIe bytes): page-locked staging
  // buffer for the Int32 rows, event recorded once they have arrived (the download of the values waits for it, so that the two
  // copies do not share the link and the widening overlaps the second one)
  void *host_stage = nullptr;
  size_t host_stage_bytes = 0;
  cudaEvent_t pattern_copied = nullptr;
  bool pattern_copied_pending = false;
  int num_sms = 148;
  std::string last_error;
  int64_t launches = 0;
  std::vector<gb::Timing> timings;
  std::vector<gb::PendingTiming> pending;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool deterministic() const { return flags & GB200_FLAG_DETERMINISTIC; }
};

struct gb200_mesh_s {
  gb200_ctx ctx;
  int D = 0, Dr = 0, nn = 0, celltype = 0;  // D: space dimension; Dr: dimension of the cell type (Dr = D - 1: boundary facets)
  int64_t nnodes = 0, ncells = 0;
  gb::DevBuf<double> X;            // [nnodes][D]
  gb::DevBuf<int32_t> cell_nodes;  // [ncells][nn], 0-based
  int affine = -1;                 // -1 unknown, 0/1
};

struct gb200_refel_s {
  gb200_ctx ctx;
  int D = 0, np = 0, nd = 0, ncomp = 1;
  std::vector<double> w, N, dN;  // host copies in kernel layout: N[p][a], dN[p][a][d]
};

struct gb200_space_s {
  gb200_ctx ctx;
  gb200_mesh mesh;
  gb200_refel refel;
  int nld = 0;  // local dofs per cell = nd*ncomp
  int64_t nfree = 0, ndir = 0;
  gb::DevBuf<int32_t> cell_dofs;  // [ncells][nld] signed, 1-based ids (as on the wire)
  std::vector<int32_t> h_cell_dofs;  // host copy (colouring)
};

namespace gb {

constexpr int MAX_FIELDS = 2;

// Reference-element data + per-field layout handed to the element kernels (lives in device global memory).
struct FieldDesc {
  int nds, ncomp, nld;   // scalar shape fns, components, local dofs
  int lofs;              // offset of this field's dofs in the concatenated local numbering
  int64_t row_off, col_off;
  const double *N;       // [np][nds]
  const double *dN;      // [np][nds][D]
  const int32_t *row_ids;  // test space cell dofs [ncells][nld]
  const int32_t *col_ids;  // trial space cell dofs
  const double *free_vals;  // state of the trial FE function (may be null)
  const double *dir_vals;
  const double *src_fq;     // source term of this field's rows at the quadrature points [ncells][np][ncomp] (null: constant src)
  double src[3];            // constant source per component (l((v,q)) = int(v.f + q*g): one source per field)
  const int32_t *state_ids; // cell dof ids the state u_h is gathered through: col_ids, or the unmasked ids of the global trial
                            // space when the plan's columns are masked / renumbered (owned-column plans, gb200_plan_set_state_space)
  int tab_ofs;           // offset (in doubles) of this field's physical gradients inside the team scratch
};

struct ElemDesc {
  int D, Dr, nn, np, nfields;  // Dr < D: embedded facets (measure sqrt(det(Jt J)), no gradients)
  int NL;                // total local dofs (all fields)
  const double *w;       // [np]
  const double *Ng;      // [np][nn]
  const double *dNg;     // [np][nn][D]
  const double *X;
  const int32_t *cell_nodes;
  int64_t ncells;
  FieldDesc f[MAX_FIELDS];
  unsigned char touched[MAX_FIELDS][MAX_FIELDS];  // [bi][bj]
  // facet-of-cell plans (gb200_plan_set_facets): local face of every "cell" (0-based) and the scaled reference normals; np is then
  // the number of points per facet and the tabulations hold one block of np points per local face
  const int32_t *lface;
  const double *nref;    // [nlf][D]
  // skeleton plans (gb200_plan_set_skeleton): field 0 lives on the plus cells (cell_nodes / lface), field 1 on the minus cells
  int skel;
  const double *X2;
  const int32_t *cell_nodes2, *lface2, *perm;   // perm[facet][np]: minus-side point that coincides with plus-side point p
};

}  // namespace gb

struct gb200_plan_s {
  gb200_ctx ctx;
  gb200_mesh mesh;
  gb200_refel geo;
  int nfields = 0;
  std::vector<gb200_space> test, trial;
  std::vector<uint8_t> touched;  // [bi + nf*bj]
  std::vector<int64_t> row_off, col_off;
  int64_t nrows = 0, ncols = 0, nnz = 0;
  int NL = 0;  // concatenated local dofs

  // pattern (device): 0-based colptr (int64), 0-based rowval (int32)
  gb::DevBuf<int64_t> colptr;
  gb::DevBuf<int32_t> rowval;
  // cell-centric slot map: rank of row li inside column of lj, [ncells][NL(lj)][NL(li)], 0xFFFF = not stored
  gb::DevBuf<uint16_t> rank;
  // results
  gb::DevBuf<double> nzval, bvec;
  // BlockMultiFieldStyle views of the pattern: per field block (bi + nf*bj) the start of its piece of every column and its colptr
  std::vector<gb::DevBuf<int64_t>> block_beg, block_ptr;
  std::vector<int64_t> block_nnz;
  // SparseMatrixCSR view of the pattern: rowptr, colval (ascending inside a row) and the CSC slot of every CSR position
  gb::DevBuf<int64_t> csr_rowptr, csr_src;
  gb::DevBuf<int32_t> csr_colval;
  // tabulation on device
  gb::DevBuf<double> tab;        // all tabulated arrays packed
  gb::DevBuf<double> state[gb::MAX_FIELDS][2];  // free / dirichlet values per field
  gb200_space state_space[gb::MAX_FIELDS] = {nullptr, nullptr};  // null: the trial space
  gb::ElemDesc ed;               // host copy of the descriptor (pointers are device pointers)
  gb::DevBuf<double> fq;         // source at quadrature points
  gb::DevBuf<int32_t> lface;     // facet-of-cell plans
  gb::DevBuf<int32_t> lface2, skel_perm;   // skeleton plans
  gb::DevBuf<double> nref;
  // colouring (deterministic generic path)
  int ncolors = 0;
  std::vector<int64_t> color_ptr;      // [ncolors+1]
  gb::DevBuf<int32_t> color_cells;     // cells sorted by colour
  // owner-computes gather plan for Q1 elements (column -> incident (cell, lj) list + packed ranks)
  bool has_gather = false;
  bool gather_plan_pending = false;
  bool adj_ready = false;            // adj_ptr / adj_cell / adj_rank already built (and sorted) by the fast symbolic phase  // built by the first numeric call that can use it (ensure_gather_plan)
  gb::DevBuf<int64_t> adj_ptr;    // [ncols+1]
  gb::DevBuf<int32_t> adj_cell;   // cell*8 + lj, ascending
  gb::DevBuf<uint64_t> adj_rank;  // 8 x u8 ranks of the rows of that cell inside the column (0xFF = none)
  // the same adjacency, blocked by 32 columns and transposed: row (blk_ptr[b] + q), lane l <-> column 32 b + l
  gb::DevBuf<int64_t> blk_ptr;    // [nblocks+1]
  gb::DevBuf<uint8_t> blk_flag;   // 1 = full 3x3x3 stencil block, 2 = stencil subset (col_mask), 0 = generic
  gb::DevBuf<int32_t> blk_base;   // flag bit 4: row q of the block is the run base[q] + 8*lane (no adjT loads needed)
  gb::DevBuf<uint32_t> col_mask;  // present stencil positions per column (flag-2 blocks)
  gb::DevBuf<int32_t> adjT_cell;  // -1 = no entry
  gb::DevBuf<uint64_t> adjT_rank;
  gb::DevBuf<double> cellG;       // per-cell geometric factors (affine path): 7 doubles, SoA [7][ncells]
  int gather_ok = -1;             // cached eligibility of the gather path (affine mesh, exact Q1 tabulation)
  int gather_diag = -1;           // cached: the metric of every cell is diagonal (3 factors per cell instead of 6)
  int gather_box = -1;            // cached: every cell is an axis-aligned box with bitwise equal parallel edges (factors from 4 nodes)
  int gather_ctas_per_sm[5] = {0, 0, 0, 0, 0};  // occupancy of the gather kernel instances (Laplacian, Laplacian diagonal, mass, staged)
  int64_t gather_span_max = 0;    // max nnz covered by one CTA of the gather kernel
  // the two launches of a headline step (cell_geom + gather) as one CUDA graph, re-instantiated when the arguments change
  cudaGraphExec_t gather_graph = nullptr;
  int gather_graph_form = -1, gather_graph_add = -1, gather_graph_calls = 0;
  double gather_graph_coef = 0.0;
  double *gather_graph_nzval = nullptr;
  const double *gather_graph_G = nullptr;
  ~gb200_plan_s() { if (gather_graph) cudaGraphExecDestroy(gather_graph); }
  // column-node gather for affine cells (affine_gather.cu): trial node -> incident (cell, local node) lists, reference tensors,
  // per-cell factors I = inv(Jt) and |det|
  int cng_ok = -1;
  bool cng_built = false;
  gb::DevBuf<int64_t> cng_adj, cng_unit_ptr;
  gb::DevBuf<char> cng_nodes;
  gb::DevBuf<double> cng_tab, cellF;
  int cng_oM = 0, cng_omass = 0, cng_oC = 0, cng_oMs = 0, cng_buf_len = 0;
  int64_t cng_nunits = 0, cng_nent = 0;
  // block-owner gather (affine_gather.cu): source lists and records of the stored node-pair blocks; 0 not built yet, 1 built, -1 n/a
  int bog_state = 0;
  int cng_diag = -1;   // inv(Jt) diagonal in every cell (axis-aligned boxes)
  gb::DevBuf<int64_t> bog_src;
  gb::DevBuf<double> ke_stage;   // staged mode: [ncells][pairs a <= b][9]
  gb::DevBuf<char> bog_blocks;
  int64_t bog_nblocks = 0;
  gb::DevBuf<uint32_t> bog_pairs;   // owned blocks and their mirror images: {t, t' | 0xFFFFFFFF} (symmetric forms: one evaluation, two writes)
  int64_t bog_npairs = 0;
  gb::DevBuf<int32_t> dir_cells;  // cells with a Dirichlet DoF (Q1 RHS lifting pass), built on first use
  int64_t n_dir_cells = -1;
  std::map<int, std::string> path;
  std::map<int, std::string> path_full;
  std::map<int, std::string> path_detail;  // e.g. "dmma" when the FP64 tensor-core instance of the vector kernel ran
};

namespace gb {
// Event timer: records two events on the context stream around the enclosed region, without any host
// synchronisation.  resolve_timings() (called after the final stream sync of an API call) turns them into ms.
struct ScopedTimer {
  gb200_ctx ctx;
  std::string name;
  cudaEvent_t a, b;
  ScopedTimer(gb200_ctx c, const char *n) : ctx(c), name(n) {
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, ctx->stream);
  }
  ~ScopedTimer() {
    cudaEventRecord(b, ctx->stream);
    ctx->pending.push_back({name, a, b});
  }
};
inline void resolve_timings(gb200_ctx ctx) {
  for (auto &p : ctx->pending) {
    float ms = 0;
    cudaEventSynchronize(p.b);
    cudaEventElapsedTime(&ms, p.a, p.b);
    ctx->timings.push_back({p.name, ms});
    cudaEventDestroy(p.a);
    cudaEventDestroy(p.b);
  }
  ctx->pending.clear();
}

inline void count_launch(gb200_ctx ctx, int n = 1) { ctx->launches += n; }
inline void check_launch(gb200_ctx ctx, const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw Error(GB200_ERR_CUDA, fmt("launch of %s failed: %s", what, cudaGetErrorString(e)));
  ctx->launches += 1;
}

// ---- implemented in symbolic.cu
int64_t ids_to_zero_based(gb200_ctx ctx, int32_t *ids, int64_t n, int64_t nmax);
int64_t count_ids_out_of_range(gb200_ctx ctx, const int32_t *ids, int64_t n, int64_t nfree, int64_t ndir);
void build_pattern(gb200_plan plan);
void build_gather_plan(gb200_plan plan);
void ensure_gather_plan(gb200_plan plan);
void add_matrix_from(gb200_plan dst, gb200_plan src);
void fold_constraints(gb200_plan dst, gb200_plan src, const int64_t *h_ptrs, const int32_t *h_mdofs, const double *h_coeffs, const double *h_dir,
                      int64_t ndir, bool with_matrix, bool with_vector);
void csr_to_host(gb200_plan plan, int64_t base, int64_t *rowptr, int64_t *colval, double *nzval);
int64_t block_layout(gb200_plan plan, int bi, int bj);
void block_to_host(gb200_plan plan, int bi, int bj, int64_t *colptr, int64_t *rowval, double *nzval);
void pattern_to_host(gb200_plan plan, int64_t *colptr, int64_t *rowval, bool async);
inline void sync_copies(gb200_ctx ctx) {
  if (ctx->copy_pending) {
    cudaError_t e = cudaStreamSynchronize(ctx->copy_stream);
    ctx->copy_pending = false;
    ctx->pattern_copied_pending = false;
    for (void *p : ctx->copy_keep) dev_free(p);
    ctx->copy_keep.clear();
    if (e != cudaSuccess) throw Error(GB200_ERR_CUDA, fmt("asynchronous pattern download failed: %s", cudaGetErrorString(e)));
  }
}
// ---- implemented in element_kernels.cu
struct NumericArgs {
  int form_mat = 0, form_vec = 0;
  double params[8] = {0};
  bool lift = false;
  const double *fq = nullptr;  // device
  const double *Ke_const = nullptr;  // device, column-major [NL][NL]
  bool skip_block00 = false;         // multi-field: block (0,0) already assembled by the specialised vector kernel
};
void launch_generic(gb200_plan plan, const NumericArgs &a, double *nzval, double *bvec);
void launch_quadrature_points(gb200_plan plan, double *xq_dev);
// ---- implemented in q1hex_rhs.cu
bool launch_q1hex_rhs(gb200_plan plan, int form_vec, int lift_form, const double *params, const double *fq, double *bvec);
// ---- implemented in vector_kernels.cu
bool launch_vector_kernel(gb200_plan plan, int form, int form_vec, const double *params, const double *fq, double *nzval, double *bvec,
                          double *ke_out = nullptr);
bool vector_kernel_pairs(gb200_plan plan, int &npair);   // node pairs a <= b of the plan's element when launch_vector_kernel has an instance
// ---- implemented in affine_gather.cu (returns false when the plan / form is outside its set: the caller uses the cell-centric kernels)
bool launch_affine_gather(gb200_plan plan, int form, const double *params, double *nzval, bool add);
// any geometry / state-dependent forms: cell-centric blocks staged in HBM (vector_kernels.cu), summed per stored block by the block-owner gather
bool launch_staged_gather(gb200_plan plan, int form, int form_vec, const double *params, const double *fq, double *nzval, double *bvec, bool add, bool zero_vec);
constexpr int FORM_STAGED = 100;   // (internal) instance of the block-owner gather that sums staged blocks
// ---- implemented in q1hex_gather.cu
bool gather_supported(gb200_plan plan, int form);
int gather_mode(gb200_plan plan, int form);
void launch_gather(gb200_plan plan, int form, const double *params, double *nzval, bool add);
bool cells_are_boxes(gb200_plan plan);   // every cell an axis-aligned box with bitwise equal parallel edges (checked once per plan)
// ---- implemented in mesh.cu
int mesh_check_affine(gb200_mesh mesh);
}  // namespace gb
