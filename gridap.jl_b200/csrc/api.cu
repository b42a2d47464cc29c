// api.cu -- the C ABI of libgridap_b200.so (see include/gridap_b200.h for the reference interface each call replaces).
#include <algorithm>
#include <mutex>
#include <sstream>
#include <unordered_map>

#include "common.cuh"

using namespace gb;

#define GB200_STR2(x) #x
#define GB200_STR(x) GB200_STR2(x)

static thread_local std::string g_last_error = "";

// ---------------------------------------------------------------------------------------------- device block cache
namespace gb {
namespace {
struct BlockCache {
  std::mutex mu;
  std::unordered_map<void *, std::pair<size_t, cudaStream_t>> live;        // block -> (size, stream it belongs to)
  std::map<cudaStream_t, std::multimap<size_t, void *>> free_blocks;      // per stream, by size
  size_t cached_bytes = 0;
};
BlockCache &cache() {
  static BlockCache *c = new BlockCache();  // leaked on purpose: DevBufs may be released during process teardown
  return *c;
}
size_t round_size(size_t bytes) {
  const size_t g = bytes >= (size_t(1) << 20) ? (size_t(2) << 20) : 512;  // 2 MB granules for large blocks
  return (bytes + g - 1) / g * g;
}
void release_cached(BlockCache &c, cudaStream_t only) {  // caller holds the lock
  for (auto &kv : c.free_blocks) {
    if (only && kv.first != only) continue;
    if (!kv.second.empty()) cudaStreamSynchronize(kv.first);
    for (auto &b : kv.second) {
      cudaFree(b.second);
      c.cached_bytes -= b.first;
    }
    kv.second.clear();
  }
}
}  // namespace

void *dev_alloc(size_t bytes) {
  BlockCache &c = cache();
  const size_t sz = round_size(bytes);
  const cudaStream_t s = g_alloc_stream;
  std::lock_guard<std::mutex> lock(c.mu);
  auto &fl = c.free_blocks[s];
  auto it = fl.lower_bound(sz);
  if (it != fl.end() && it->first <= sz + sz / 8) {  // close enough in size: reuse (stream order protects earlier users)
    void *p = it->second;
    c.cached_bytes -= it->first;
    c.live[p] = {it->first, s};
    fl.erase(it);
    return p;
  }
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, sz);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    release_cached(c, nullptr);
    e = cudaMalloc(&p, sz);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    throw Error(GB200_ERR_CUDA, fmt("cudaMalloc of %zu bytes failed: %s", sz, cudaGetErrorString(e)));
  }
  c.live[p] = {sz, s};
  return p;
}

void dev_free(void *p) noexcept {
  if (!p) return;
  BlockCache &c = cache();
  std::lock_guard<std::mutex> lock(c.mu);
  auto it = c.live.find(p);
  if (it == c.live.end()) { cudaFree(p); return; }
  const size_t sz = it->second.first;
  const cudaStream_t s = it->second.second;
  c.live.erase(it);
  static const bool no_cache = getenv("GB200_NO_DEVICE_CACHE") != nullptr;
  if (no_cache) { cudaFree(p); return; }
  c.free_blocks[s].insert({sz, p});
  c.cached_bytes += sz;
}

void dev_cache_trim(cudaStream_t only) {
  BlockCache &c = cache();
  std::lock_guard<std::mutex> lock(c.mu);
  release_cached(c, only);
}
}  // namespace gb

template <class F>
static int32_t guarded(gb200_ctx ctx, F &&f) {
  try {
    if (ctx) {
      GB_CUDA(cudaSetDevice(ctx->device));
      gb::g_alloc_stream = ctx->stream;
    }
    f();
    return GB200_OK;
  } catch (const gb::Error &e) {
    g_last_error = e.what();
    if (ctx) ctx->last_error = e.what();
    cudaGetLastError();
    return e.code;
  } catch (const std::exception &e) {
    g_last_error = e.what();
    if (ctx) ctx->last_error = e.what();
    return GB200_ERR_INVALID;
  }
}

static int nodes_of(int celltype) {
  switch (celltype) {
    case GB200_QUAD4: return 4;
    case GB200_HEX8: return 8;
    case GB200_TRI3: return 3;
    case GB200_TET4: return 4;
    case GB200_SEG2: return 2;
  }
  return 0;
}
// Table{Int32}.ptrs of a table whose rows all have `len` entries: first row that does not, or -1.  The common (regular)
// case is one branch-free, vectorisable pass.
static int64_t first_irregular_row(const int32_t *ptrs, int64_t nrows, int len) {
  int32_t diff = 0;
  for (int64_t c = 0; c < nrows; c++) diff |= (ptrs[c + 1] - ptrs[c]) ^ len;
  if (!diff) return -1;
  for (int64_t c = 0; c < nrows; c++)
    if (ptrs[c + 1] - ptrs[c] != len) return c;
  return -1;
}
static int dim_of(int celltype) { return celltype == GB200_SEG2 ? 1 : (celltype == GB200_QUAD4 || celltype == GB200_TRI3) ? 2 : 3; }

extern "C" {

const char *gb200_version(void) { return "gridap_b200 0.1.0 (sm_100a; CUDA " GB200_STR(CUDART_VERSION) ")"; }

const char *gb200_last_error(gb200_ctx ctx) { return ctx ? ctx->last_error.c_str() : g_last_error.c_str(); }

int32_t gb200_init(int32_t device, uint32_t flags, gb200_ctx *out) {
  if (!out) return GB200_ERR_INVALID;
  *out = nullptr;
  return guarded(nullptr, [&] {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    GB_REQUIRE(e == cudaSuccess && n > 0, GB200_ERR_CUDA, "no CUDA device available (%s); libgridap_b200 has no CPU path",
               e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    GB_REQUIRE(device >= 0 && device < n, GB200_ERR_INVALID, "device %d out of range (found %d)", device, n);
    GB_CUDA(cudaSetDevice(device));
    auto *ctx = new gb200_ctx_s();
    ctx->device = device;
    ctx->flags = flags;
    GB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    GB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    cudaDeviceProp prop;
    GB_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->num_sms = prop.multiProcessorCount;
    *out = ctx;
  });
}

int32_t gb200_finalize(gb200_ctx ctx) {
  if (!ctx) return GB200_ERR_INVALID;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->copy_stream);
  for (void *p : ctx->copy_keep) gb::dev_free(p);
  ctx->copy_keep.clear();
  if (ctx->host_stage) cudaFreeHost(ctx->host_stage);
  if (ctx->pattern_copied) cudaEventDestroy(ctx->pattern_copied);
  gb::dev_cache_trim(ctx->stream);
  cudaStreamDestroy(ctx->copy_stream);
  cudaStreamDestroy(ctx->stream);
  if (gb::g_alloc_stream == ctx->stream) gb::g_alloc_stream = nullptr;
  delete ctx;
  return GB200_OK;
}

int32_t gb200_get_timings(gb200_ctx ctx, char *buf, size_t len) {
  if (!ctx || !buf || !len) return GB200_ERR_INVALID;
  // mean device time per region name over all calls since the previous gb200_get_timings, then reset
  resolve_timings(ctx);
  std::vector<std::string> names;
  std::map<std::string, std::pair<double, int>> acc;
  for (auto &t : ctx->timings) {
    if (!acc.count(t.name)) names.push_back(t.name);
    acc[t.name].first += t.ms;
    acc[t.name].second += 1;
  }
  std::ostringstream os;
  os << "{";
  for (size_t i = 0; i < names.size(); i++) os << (i ? "," : "") << "\"" << names[i] << "\":" << acc[names[i]].first / acc[names[i]].second;
  os << "}";
  snprintf(buf, len, "%s", os.str().c_str());
  ctx->timings.clear();
  return GB200_OK;
}

int32_t gb200_host_alloc(gb200_ctx ctx, size_t bytes, void **p) {
  if (!ctx || !p) return GB200_ERR_INVALID;
  return guarded(ctx, [&] { GB_CUDA(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault)); });
}
int32_t gb200_host_free(gb200_ctx ctx, void *p) {
  if (!ctx) return GB200_ERR_INVALID;
  return guarded(ctx, [&] { GB_CUDA(cudaFreeHost(p)); });
}
int32_t gb200_host_register(gb200_ctx ctx, void *p, size_t bytes) {
  if (!ctx || !p) return GB200_ERR_INVALID;
  return guarded(ctx, [&] {
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return; }
    GB_CUDA(e);
  });
}
int32_t gb200_host_unregister(gb200_ctx ctx, void *p) {
  if (!ctx || !p) return GB200_ERR_INVALID;
  return guarded(ctx, [&] {
    cudaError_t e = cudaHostUnregister(p);
    if (e == cudaErrorHostMemoryNotRegistered) { cudaGetLastError(); return; }
    GB_CUDA(e);
  });
}

int32_t gb200_trim(gb200_ctx ctx) {
  if (!ctx) return GB200_ERR_INVALID;
  return guarded(ctx, [&] {
    GB_CUDA(cudaStreamSynchronize(ctx->stream));
    sync_copies(ctx);
    dev_cache_trim(ctx->stream);
  });
}

int64_t gb200_launch_count(gb200_ctx ctx) { return ctx ? ctx->launches : 0; }
void *gb200_stream(gb200_ctx ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int32_t gb200_synchronize(gb200_ctx ctx) {
  return guarded(ctx, [&] {
    GB_CUDA(cudaStreamSynchronize(ctx->stream));
    sync_copies(ctx);
  });
}

// ---------------------------------------------------------------------------------------------- mesh
int32_t gb200_mesh_create(gb200_ctx ctx, int32_t D, int64_t nnodes, const double *coords, int64_t ncells, const int32_t *cell_node_data,
                          const int32_t *cell_node_ptrs, int32_t celltype, gb200_mesh *out) {
  if (!ctx || !out) return GB200_ERR_INVALID;
  *out = nullptr;
  return guarded(ctx, [&] {
    int nn = nodes_of(celltype);
    GB_REQUIRE(nn > 0, GB200_ERR_UNSUPPORTED, "cell type %d is not supported (QUAD4, HEX8, TRI3, TET4, SEG2)", celltype);
    GB_REQUIRE((D == dim_of(celltype) && celltype != GB200_SEG2) || (D == dim_of(celltype) + 1 && D <= 3), GB200_ERR_INVALID,
               "cell type %d lives in %dD (or, as boundary facets, in %dD), got D=%d", celltype, dim_of(celltype), dim_of(celltype) + 1, D);
    GB_REQUIRE(coords && cell_node_data && cell_node_ptrs && nnodes > 0 && ncells >= 0, GB200_ERR_INVALID, "null / empty mesh arrays");
    GB_REQUIRE(ncells * nn < (int64_t)1 << 31, GB200_ERR_UNSUPPORTED, "more than 2^31 cell-node entries");
    GB_REQUIRE(cell_node_ptrs[0] == 1, GB200_ERR_INVALID, "cell_node_ptrs must start at 1");
    GB_REQUIRE((int64_t)cell_node_ptrs[ncells] - 1 == ncells * nn, GB200_ERR_UNSUPPORTED,
               "cell_node_data holds %lld entries, expected %lld (%d nodes per cell of the declared type)",
               (long long)cell_node_ptrs[ncells] - 1, (long long)(ncells * nn), nn);
    auto *m = new gb200_mesh_s();
    m->ctx = ctx;
    m->D = D;
    m->Dr = dim_of(celltype);
    m->nn = nn;
    m->celltype = celltype;
    m->nnodes = nnodes;
    m->ncells = ncells;
    m->X.upload(coords, (size_t)nnodes * D, ctx->stream);
    m->cell_nodes.upload(cell_node_data, (size_t)ncells * nn, ctx->stream);
    // the row-length check of the Table runs on the host while the copies are in flight
    int64_t bad_cell = first_irregular_row(cell_node_ptrs, ncells, nn);
    if (bad_cell >= 0) {
      cudaStreamSynchronize(ctx->stream);
      delete m;
      throw gb::Error(GB200_ERR_UNSUPPORTED, fmt("cell %lld has %d nodes; all cells must be of the declared type (%d nodes)",
                                                 (long long)bad_cell + 1, cell_node_ptrs[bad_cell + 1] - cell_node_ptrs[bad_cell], nn));
    }
    // 1-based -> 0-based and range check on the device (no host pass over the connectivity)
    int64_t bad = ids_to_zero_based(ctx, m->cell_nodes.p, ncells * nn, nnodes);
    if (bad) {
      delete m;
      throw gb::Error(GB200_ERR_INVALID, fmt("%lld node ids are outside 1..%lld", (long long)bad, (long long)nnodes));
    }
    *out = m;
  });
}
int32_t gb200_mesh_destroy(gb200_mesh m) {
  if (!m) return GB200_ERR_INVALID;
  cudaSetDevice(m->ctx->device);
  gb::g_alloc_stream = m->ctx->stream;
  delete m;
  return GB200_OK;
}
int32_t gb200_mesh_is_affine(gb200_mesh m, int32_t *is_affine) {
  if (!m || !is_affine) return GB200_ERR_INVALID;
  return guarded(m->ctx, [&] { *is_affine = mesh_check_affine(m); });
}

// ---------------------------------------------------------------------------------------------- refel
int32_t gb200_refel_create(gb200_ctx ctx, int32_t D, int32_t np, int32_t nd, int32_t ncomp, const double *w, const double *N,
                           const double *dN, gb200_refel *out) {
  if (!ctx || !out) return GB200_ERR_INVALID;
  *out = nullptr;
  return guarded(ctx, [&] {
    GB_REQUIRE(D >= 1 && D <= 3, GB200_ERR_UNSUPPORTED, "D=%d", D);
    GB_REQUIRE(np > 0 && np <= 64 && nd > 0 && ncomp >= 1 && ncomp <= 3, GB200_ERR_UNSUPPORTED,
               "reference element out of range (np=%d nd=%d ncomp=%d)", np, nd, ncomp);
    GB_REQUIRE(w && N && dN, GB200_ERR_INVALID, "null tabulation arrays");
    auto *r = new gb200_refel_s();
    r->ctx = ctx;
    r->D = D;
    r->np = np;
    r->nd = nd;
    r->ncomp = ncomp;
    r->w.assign(w, w + np);
    r->N.resize((size_t)np * nd);
    r->dN.resize((size_t)np * nd * D);
    for (int p = 0; p < np; p++)
      for (int a = 0; a < nd; a++) {
        r->N[p * nd + a] = N[p + np * a];  // Julia Matrix [np,nd] -> [p][a]
        for (int d = 0; d < D; d++) r->dN[(p * nd + a) * D + d] = dN[d + D * (p + np * a)];
      }
    *out = r;
  });
}
int32_t gb200_refel_destroy(gb200_refel r) {
  if (!r) return GB200_ERR_INVALID;
  delete r;
  return GB200_OK;
}

// ---------------------------------------------------------------------------------------------- space
int32_t gb200_space_create(gb200_ctx ctx, gb200_mesh mesh, gb200_refel refel, const int32_t *cell_dof_data, const int32_t *cell_dof_ptrs,
                           int64_t nfree, int64_t ndir, gb200_space *out) {
  if (!ctx || !out) return GB200_ERR_INVALID;
  *out = nullptr;
  return guarded(ctx, [&] {
    GB_REQUIRE(mesh && refel && cell_dof_data && cell_dof_ptrs, GB200_ERR_INVALID, "null argument");
    GB_REQUIRE(refel->D == mesh->Dr, GB200_ERR_INVALID, "reference element and mesh (cell type) dimensions differ");
    int nld = refel->nd * refel->ncomp;
    auto *s = new gb200_space_s();
    s->ctx = ctx;
    s->mesh = mesh;
    s->refel = refel;
    s->nld = nld;
    s->nfree = nfree;
    s->ndir = ndir;
    GB_REQUIRE(cell_dof_ptrs[0] == 1, GB200_ERR_INVALID, "cell_dof_ptrs must start at 1");
    if ((int64_t)cell_dof_ptrs[mesh->ncells] - 1 != mesh->ncells * nld) {
      delete s;
      // spaces with a varying number of DoFs per cell / constraints are outside the supported set
      throw gb::Error(GB200_ERR_UNSUPPORTED, fmt("cell_dof_data holds %lld entries, expected %lld (%d DoFs per cell)",
                                                 (long long)cell_dof_ptrs[mesh->ncells] - 1, (long long)(mesh->ncells * nld), nld));
    }
    s->cell_dofs.upload(cell_dof_data, (size_t)mesh->ncells * nld, ctx->stream);
    int64_t c = first_irregular_row(cell_dof_ptrs, mesh->ncells, nld);  // on the host while the copy is in flight
    if (c >= 0) {
      cudaStreamSynchronize(ctx->stream);
      delete s;
      throw gb::Error(GB200_ERR_UNSUPPORTED, fmt("cell %lld has %d DoFs, expected %d", (long long)c + 1,
                                                 cell_dof_ptrs[c + 1] - cell_dof_ptrs[c], nld));
    }
    int64_t bad = count_ids_out_of_range(ctx, s->cell_dofs.p, mesh->ncells * nld, nfree, ndir);
    if (bad) {
      delete s;
      throw gb::Error(GB200_ERR_INVALID, fmt("%lld DoF ids are out of range (nfree=%lld ndirichlet=%lld)", (long long)bad,
                                             (long long)nfree, (long long)ndir));
    }
    if (ctx->deterministic()) s->h_cell_dofs.assign(cell_dof_data, cell_dof_data + (size_t)mesh->ncells * nld);  // host colouring
    *out = s;
  });
}
int32_t gb200_space_destroy(gb200_space s) {
  if (!s) return GB200_ERR_INVALID;
  cudaSetDevice(s->ctx->device);
  gb::g_alloc_stream = s->ctx->stream;
  delete s;
  return GB200_OK;
}

// ---------------------------------------------------------------------------------------------- plan
// Greedy cell colouring on the host (deterministic generic path): two cells of one colour share no row / column id.
static void color_cells(gb200_plan plan) {
  const int64_t nc = plan->mesh->ncells;
  std::vector<uint64_t> row_mask((size_t)plan->nrows, 0), col_mask((size_t)plan->ncols, 0);
  std::vector<uint8_t> color((size_t)nc, 0);
  int ncolors = 0;
  for (int64_t c = 0; c < nc; c++) {
    uint64_t used = 0;
    for (int f = 0; f < plan->nfields; f++) {
      const auto *t = plan->test[f];
      const auto *u = plan->trial[f];
      for (int k = 0; k < t->nld; k++) {
        int32_t r = t->h_cell_dofs[c * t->nld + k], cc = u->h_cell_dofs[c * u->nld + k];
        if (r > 0) used |= row_mask[r - 1 + plan->row_off[f]];
        if (cc > 0) used |= col_mask[cc - 1 + plan->col_off[f]];
      }
    }
    int col = 0;
    while (col < 64 && ((used >> col) & 1)) col++;
    GB_REQUIRE(col < 64, GB200_ERR_UNSUPPORTED, "mesh needs more than 64 colours for the deterministic scatter");
    color[c] = (uint8_t)col;
    ncolors = std::max(ncolors, col + 1);
    for (int f = 0; f < plan->nfields; f++) {
      const auto *t = plan->test[f];
      const auto *u = plan->trial[f];
      for (int k = 0; k < t->nld; k++) {
        int32_t r = t->h_cell_dofs[c * t->nld + k], cc = u->h_cell_dofs[c * u->nld + k];
        if (r > 0) row_mask[r - 1 + plan->row_off[f]] |= 1ull << col;
        if (cc > 0) col_mask[cc - 1 + plan->col_off[f]] |= 1ull << col;
      }
    }
  }
  plan->ncolors = ncolors;
  plan->color_ptr.assign(ncolors + 1, 0);
  for (int64_t c = 0; c < nc; c++) plan->color_ptr[color[c] + 1]++;
  for (int k = 0; k < ncolors; k++) plan->color_ptr[k + 1] += plan->color_ptr[k];
  std::vector<int64_t> cur(plan->color_ptr.begin(), plan->color_ptr.end() - 1);
  std::vector<int32_t> order((size_t)nc);
  for (int64_t c = 0; c < nc; c++) order[cur[color[c]]++] = (int32_t)c;
  plan->color_cells.upload(order.data(), order.size(), plan->ctx->stream);
  GB_CUDA(cudaStreamSynchronize(plan->ctx->stream));
}

int32_t gb200_plan_create(gb200_ctx ctx, gb200_mesh mesh, gb200_refel geo, int32_t ntest, const gb200_space *test_spaces, int32_t ntrial,
                          const gb200_space *trial_spaces, const uint8_t *touched, const int64_t *row_offsets, const int64_t *col_offsets,
                          int64_t nrows, int64_t ncols, gb200_plan *out) {
  if (!ctx || !out) return GB200_ERR_INVALID;
  *out = nullptr;
  gb200_plan_s *plan = nullptr;
  int32_t rc = guarded(ctx, [&] {
    GB_REQUIRE(mesh && geo && test_spaces && trial_spaces, GB200_ERR_INVALID, "null argument");
    GB_REQUIRE(ntest == ntrial, GB200_ERR_UNSUPPORTED, "ntest (%d) != ntrial (%d): only Galerkin pairs of fields are supported", ntest, ntrial);
    GB_REQUIRE(ntest >= 1 && ntest <= MAX_FIELDS, GB200_ERR_UNSUPPORTED, "%d fields (max %d)", ntest, MAX_FIELDS);
    GB_REQUIRE(geo->nd == mesh->nn && geo->ncomp == 1 && geo->D == mesh->Dr, GB200_ERR_INVALID,
               "geometry reference element does not match the cell type");
    GB_REQUIRE(nrows > 0 && ncols > 0 && nrows < ((int64_t)1 << 31) && ncols < ((int64_t)1 << 31), GB200_ERR_UNSUPPORTED,
               "system size out of range");
    plan = new gb200_plan_s();
    plan->ctx = ctx;
    plan->mesh = mesh;
    plan->geo = geo;
    plan->nfields = ntest;
    plan->nrows = nrows;
    plan->ncols = ncols;
    int NL = 0;
    for (int f = 0; f < ntest; f++) {
      gb200_space t = test_spaces[f], u = trial_spaces[f];
      // (field 1 of a skeleton plan lives on the mesh of the minus cells: same cell type, one cell per facet)
      GB_REQUIRE(t && u && t->mesh == u->mesh && (t->mesh == mesh || (f == 1 && ntest == 2 && t->mesh->ncells == mesh->ncells && t->mesh->celltype == mesh->celltype && t->mesh->D == mesh->D)),
                 GB200_ERR_INVALID, "space %d lives on another mesh", f);
      GB_REQUIRE(t->refel->nd == u->refel->nd && t->refel->ncomp == u->refel->ncomp && t->refel->np == geo->np && u->refel->np == geo->np,
                 GB200_ERR_UNSUPPORTED, "test/trial reference elements of field %d differ (or use another quadrature)", f);
      plan->test.push_back(t);
      plan->trial.push_back(u);
      plan->row_off.push_back(row_offsets ? row_offsets[f] : 0);
      plan->col_off.push_back(col_offsets ? col_offsets[f] : 0);
      NL += t->nld;
    }
    plan->NL = NL;
    plan->touched.assign((size_t)ntest * ntest, 1);
    if (touched) plan->touched.assign(touched, touched + ntest * ntest);

    // pack the tabulation on the device: w | Ng | dNg | per field N | dN
    std::vector<double> pack;
    auto push = [&](const std::vector<double> &v) { size_t o = pack.size(); pack.insert(pack.end(), v.begin(), v.end()); return o; };
    size_t o_w = push(geo->w), o_Ng = push(geo->N), o_dNg = push(geo->dN);
    size_t o_N[MAX_FIELDS], o_dN[MAX_FIELDS];
    for (int f = 0; f < ntest; f++) { o_N[f] = push(plan->test[f]->refel->N); o_dN[f] = push(plan->test[f]->refel->dN); }
    plan->tab.upload(pack.data(), pack.size(), ctx->stream);
    ElemDesc &ed = plan->ed;
    memset(&ed, 0, sizeof(ed));
    ed.D = mesh->D; ed.Dr = mesh->Dr; ed.nn = mesh->nn; ed.np = geo->np; ed.nfields = ntest; ed.NL = NL;
    ed.w = plan->tab.p + o_w; ed.Ng = plan->tab.p + o_Ng; ed.dNg = plan->tab.p + o_dNg;
    ed.X = mesh->X.p; ed.cell_nodes = mesh->cell_nodes.p; ed.ncells = mesh->ncells;
    int lofs = 0, tofs = 0;
    for (int f = 0; f < ntest; f++) {
      FieldDesc &fd = ed.f[f];
      fd.nds = plan->test[f]->refel->nd; fd.ncomp = plan->test[f]->refel->ncomp; fd.nld = plan->test[f]->nld;
      fd.lofs = lofs; lofs += fd.nld;
      fd.row_off = plan->row_off[f]; fd.col_off = plan->col_off[f];
      fd.N = plan->tab.p + o_N[f]; fd.dN = plan->tab.p + o_dN[f];
      fd.row_ids = plan->test[f]->cell_dofs.p; fd.col_ids = plan->trial[f]->cell_dofs.p;
      fd.free_vals = nullptr; fd.dir_vals = nullptr;
      fd.state_ids = fd.col_ids;
      fd.tab_ofs = tofs; tofs += ed.np * fd.nds * ed.D;
      for (int g = 0; g < ntest; g++) ed.touched[f][g] = plan->touched[f + ntest * g];
    }
    resolve_timings(ctx);
    ctx->timings.clear();
    if (mesh->ncells == 0) {
      // empty triangulation (e.g. Triangulation(model, Int[]) in the reference's tests): empty pattern, zero results
      plan->colptr.alloc(ncols + 1);
      plan->colptr.zero(ctx->stream);
      plan->nnz = 0;
    } else {
      build_pattern(plan);
      // owner-computes gather plan (Q1 hexahedra): built by the first numeric call, so that an asynchronous download of the
      // pattern (gb200_plan_get_pattern_async) overlaps it
      if (ntest == 1 && mesh->celltype == GB200_HEX8 && NL == 8) plan->gather_plan_pending = true;
      else if (plan->adj_ready) {  // adjacency left by the fast symbolic phase, not needed without a gather plan
        plan->adj_ptr.release();
        plan->adj_cell.release();
        plan->adj_rank.release();
        plan->adj_ready = false;
      }
      if (ctx->deterministic()) color_cells(plan);
    }
    plan->nzval.alloc((size_t)std::max<int64_t>(plan->nnz, 1));
    plan->nzval.zero(ctx->stream);
    plan->bvec.alloc((size_t)nrows);
    plan->bvec.zero(ctx->stream);
    GB_CUDA(cudaStreamSynchronize(ctx->stream));
  });
  if (rc != GB200_OK) { delete plan; return rc; }
  *out = plan;
  return GB200_OK;
}

int32_t gb200_plan_set_facets(gb200_plan plan, const int32_t *lface, int32_t nlfaces, const double *nref) {
  if (!plan || !lface || !nref) return GB200_ERR_INVALID;
  return guarded(plan->ctx, [&] {
    ElemDesc &ed = plan->ed;
    GB_REQUIRE(!ed.lface, GB200_ERR_STATE, "the plan already is a facet-of-cell plan");
    GB_REQUIRE(ed.Dr == ed.D, GB200_ERR_INVALID, "facet-of-cell plans live on the cells adjacent to the facets (cell type of dimension D)");
    GB_REQUIRE(nlfaces >= 1 && ed.np % nlfaces == 0, GB200_ERR_INVALID,
               "the tabulations hold %d points: not a multiple of the %d local faces", ed.np, nlfaces);
    const int64_t nc = plan->mesh->ncells;
    std::vector<int32_t> lf((size_t)nc);
    for (int64_t c = 0; c < nc; c++) {
      GB_REQUIRE(lface[c] >= 1 && lface[c] <= nlfaces, GB200_ERR_INVALID, "local face %d of facet %lld out of range 1..%d", lface[c], (long long)c + 1, nlfaces);
      lf[(size_t)c] = lface[c] - 1;
    }
    cudaStream_t s = plan->ctx->stream;
    plan->lface.upload(lf.data(), lf.size(), s);
    plan->nref.upload(nref, (size_t)nlfaces * ed.D, s);   // nref[d + D*lf]: already [lf][d]
    GB_CUDA(cudaStreamSynchronize(s));
    ed.np /= nlfaces;
    ed.lface = plan->lface.p;
    ed.nref = plan->nref.p;
    int tofs = 0;
    for (int f = 0; f < ed.nfields; f++) { ed.f[f].tab_ofs = tofs; tofs += ed.np * ed.f[f].nds * ed.D; }
  });
}

int32_t gb200_plan_set_skeleton(gb200_plan plan, const int32_t *lface_plus, const int32_t *lface_minus, const int32_t *perm, int32_t nlfaces,
                                const double *nref) {
  if (!plan || !lface_plus || !lface_minus || !perm || !nref) return GB200_ERR_INVALID;
  return guarded(plan->ctx, [&] {
    ElemDesc &ed = plan->ed;
    GB_REQUIRE(!ed.lface, GB200_ERR_STATE, "the plan already is a facet-of-cell / skeleton plan");
    GB_REQUIRE(ed.Dr == ed.D && plan->nfields == 2, GB200_ERR_INVALID, "skeleton plans have two fields: the space on the plus and on the minus cells");
    GB_REQUIRE(ed.f[0].nds == ed.f[1].nds && ed.f[0].ncomp == ed.f[1].ncomp, GB200_ERR_INVALID, "plus and minus sides must carry the same reference FE");
    GB_REQUIRE(nlfaces >= 1 && ed.np % nlfaces == 0, GB200_ERR_INVALID, "the tabulations hold %d points: not a multiple of the %d local faces", ed.np, nlfaces);
    const int64_t nc = plan->mesh->ncells;
    const int npf = ed.np / nlfaces;
    std::vector<int32_t> lf((size_t)nc), lf2((size_t)nc);
    for (int64_t c = 0; c < nc; c++) {
      GB_REQUIRE(lface_plus[c] >= 1 && lface_plus[c] <= nlfaces && lface_minus[c] >= 1 && lface_minus[c] <= nlfaces, GB200_ERR_INVALID,
                 "local faces (%d, %d) of facet %lld out of range 1..%d", lface_plus[c], lface_minus[c], (long long)c + 1, nlfaces);
      lf[(size_t)c] = lface_plus[c] - 1;
      lf2[(size_t)c] = lface_minus[c] - 1;
      for (int p = 0; p < npf; p++)
        GB_REQUIRE(perm[c * npf + p] >= 0 && perm[c * npf + p] < npf, GB200_ERR_INVALID, "point permutation of facet %lld out of range", (long long)c + 1);
    }
    cudaStream_t s = plan->ctx->stream;
    plan->lface.upload(lf.data(), lf.size(), s);
    plan->lface2.upload(lf2.data(), lf2.size(), s);
    plan->skel_perm.upload(perm, (size_t)nc * npf, s);
    plan->nref.upload(nref, (size_t)nlfaces * ed.D, s);
    GB_CUDA(cudaStreamSynchronize(s));
    ed.np = npf;
    ed.lface = plan->lface.p;
    ed.lface2 = plan->lface2.p;
    ed.perm = plan->skel_perm.p;
    ed.nref = plan->nref.p;
    ed.skel = 1;
    ed.X2 = plan->test[1]->mesh->X.p;
    ed.cell_nodes2 = plan->test[1]->mesh->cell_nodes.p;
    int tofs = 0;
    for (int f = 0; f < ed.nfields; f++) { ed.f[f].tab_ofs = tofs; tofs += ed.np * ed.f[f].nds * ed.D; }
  });
}

int32_t gb200_plan_destroy(gb200_plan plan) {
  if (!plan) return GB200_ERR_INVALID;
  cudaSetDevice(plan->ctx->device);
  gb::g_alloc_stream = plan->ctx->stream;
  delete plan;
  return GB200_OK;
}
int32_t gb200_plan_nnz(gb200_plan plan, int64_t *nnz) {
  if (!plan || !nnz) return GB200_ERR_INVALID;
  *nnz = plan->nnz;
  return GB200_OK;
}
int32_t gb200_plan_get_pattern(gb200_plan plan, int64_t *colptr, int64_t *rowval) {
  if (!plan || !colptr || (!rowval && plan->nnz)) return GB200_ERR_INVALID;
  return guarded(plan->ctx, [&] { pattern_to_host(plan, colptr, rowval, false); });
}
int32_t gb200_plan_get_pattern_async(gb200_plan plan, int64_t *colptr, int64_t *rowval) {
  if (!plan || !colptr || (!rowval && plan->nnz)) return GB200_ERR_INVALID;
  return guarded(plan->ctx, [&] { pattern_to_host(plan, colptr, rowval, true); });
}
int32_t gb200_plan_set_state(gb200_plan plan, int32_t field, const double *free_values, const double *dirichlet_values) {
  if (!plan) return GB200_ERR_INVALID;
  return guarded(plan->ctx, [&] {
    GB_REQUIRE(field >= 0 && field < plan->nfields, GB200_ERR_INVALID, "field %d out of range", field);
    gb200_space u = plan->state_space[field] ? plan->state_space[field] : plan->trial[field];
    FieldDesc &fd = plan->ed.f[field];
    if (free_values && u->nfree) { plan->state[field][0].upload(free_values, (size_t)u->nfree, plan->ctx->stream); fd.free_vals = plan->state[field][0].p; }
    else fd.free_vals = nullptr;
    if (dirichlet_values && u->ndir) { plan->state[field][1].upload(dirichlet_values, (size_t)u->ndir, plan->ctx->stream); fd.dir_vals = plan->state[field][1].p; }
    else fd.dir_vals = nullptr;
    GB_CUDA(cudaStreamSynchronize(plan->ctx->stream));
  });
}

int32_t gb200_plan_set_state_device(gb200_plan plan, int32_t field, const double *d_free, const double *d_dir) {
  if (!plan) return GB200_ERR_INVALID;
  return guarded(plan->ctx, [&] {
    GB_REQUIRE(field >= 0 && field < plan->nfields, GB200_ERR_INVALID, "field %d out of range", field);
    gb200_space u = plan->state_space[field] ? plan->state_space[field] : plan->trial[field];
    FieldDesc &fd = plan->ed.f[field];
    cudaStream_t s = plan->ctx->stream;
    if (d_free && u->nfree) {
      if (plan->state[field][0].n != (size_t)u->nfree) plan->state[field][0].alloc((size_t)u->nfree);
      GB_CUDA(cudaMemcpyAsync(plan->state[field][0].p, d_free, (size_t)u->nfree * 8, cudaMemcpyDeviceToDevice, s));
      fd.free_vals = plan->state[field][0].p;
    }   // a null argument leaves that vector as it is (documented: the Dirichlet values of a Newton loop are set once)
    if (d_dir && u->ndir) {
      if (plan->state[field][1].n != (size_t)u->ndir) plan->state[field][1].alloc((size_t)u->ndir);
      GB_CUDA(cudaMemcpyAsync(plan->state[field][1].p, d_dir, (size_t)u->ndir * 8, cudaMemcpyDeviceToDevice, s));
      fd.dir_vals = plan->state[field][1].p;
    }
  });
}

int32_t gb200_plan_set_state_space(gb200_plan plan, int32_t field, gb200_space space) {
  if (!plan) return GB200_ERR_INVALID;
  return guarded(plan->ctx, [&] {
    GB_REQUIRE(field >= 0 && field < plan->nfields, GB200_ERR_INVALID, "field %d out of range", field);
    gb200_space u = plan->trial[field];
    if (space) {
      GB_REQUIRE(space->mesh == plan->mesh && space->nld == u->nld, GB200_ERR_INVALID,
                 "the state space must live on the plan's mesh with the trial space's local DoF layout");
    }
    plan->state_space[field] = space;
    FieldDesc &fd = plan->ed.f[field];
    fd.state_ids = space ? space->cell_dofs.p : fd.col_ids;
    fd.free_vals = nullptr;
    fd.dir_vals = nullptr;
  });
}

// ---------------------------------------------------------------------------------------------- numeric
static void check_matrix_form(gb200_plan plan, int form) {
  const ElemDesc &ed = plan->ed;
  if (form == GB200_FORM_SKELETON) {
    GB_REQUIRE(ed.skel, GB200_ERR_UNSUPPORTED, "jump / mean terms need a skeleton plan (gb200_plan_set_skeleton)");
    return;
  }
  GB_REQUIRE(!ed.skel, GB200_ERR_UNSUPPORTED, "matrix integrand %d on a skeleton plan: only GB200_FORM_SKELETON is evaluated there", form);
  if (form == GB200_FORM_FACET) {
    GB_REQUIRE(ed.lface && plan->nfields == 1, GB200_ERR_UNSUPPORTED, "normal-derivative / Nitsche terms need a single-field facet-of-cell plan (gb200_plan_set_facets)");
    return;
  }
  GB_REQUIRE(!ed.lface, GB200_ERR_UNSUPPORTED, "matrix integrand %d on a facet-of-cell plan: only GB200_FORM_FACET is evaluated there", form);
  GB_REQUIRE(ed.Dr == ed.D || form == GB200_FORM_MASS, GB200_ERR_UNSUPPORTED,
             "matrix integrand %d on boundary facets: only the mass (Robin) term is supported there", form);
  switch (form) {
    case GB200_FORM_MASS:
    case GB200_FORM_LAPLACIAN:
      GB_REQUIRE(plan->nfields == 1, GB200_ERR_UNSUPPORTED, "mass / Laplacian are single-field forms");
      return;
    case GB200_FORM_ELASTICITY:
    case GB200_FORM_NEOHOOKEAN_JAC:
      GB_REQUIRE(plan->nfields == 1 && ed.f[0].ncomp == ed.D, GB200_ERR_UNSUPPORTED, "form %d needs one vector-valued field with D components", form);
      return;
    case GB200_FORM_STOKES:
      GB_REQUIRE(plan->nfields == 2 && ed.f[0].ncomp == ed.D && ed.f[1].ncomp == 1, GB200_ERR_UNSUPPORTED,
                 "Stokes needs fields (velocity with D components, scalar pressure)");
      GB_REQUIRE(ed.touched[0][0] && ed.touched[0][1] && ed.touched[1][0] && !ed.touched[1][1], GB200_ERR_INVALID,
                 "Stokes touches blocks (v,u), (v,p), (q,u) and not (q,p)");
      return;
  }
  throw gb::Error(GB200_ERR_UNSUPPORTED,
                  fmt("matrix integrand %d is not in the supported set {mass, laplacian, elasticity, stokes, neo-Hookean Jacobian}; "
                      "there is no CPU fallback", form));
}
static void check_vector_form(gb200_plan plan, int form) {
  GB_REQUIRE(!plan->ed.skel, GB200_ERR_UNSUPPORTED, "vector integrands on a skeleton plan are not supported");
  if (form == GB200_FORM_FACET_VEC) {
    GB_REQUIRE(plan->ed.lface && plan->nfields == 1, GB200_ERR_UNSUPPORTED, "normal-derivative / Nitsche terms need a single-field facet-of-cell plan (gb200_plan_set_facets)");
    return;
  }
  GB_REQUIRE(!plan->ed.lface, GB200_ERR_UNSUPPORTED, "vector integrand %d on a facet-of-cell plan: only GB200_FORM_FACET_VEC is evaluated there", form);
  if (form == GB200_FORM_SOURCE) return;   // per-field sources of a multi-field plan: gb200_plan_set_source (params / fq per field)
  GB_REQUIRE(plan->ed.Dr == plan->ed.D, GB200_ERR_UNSUPPORTED, "vector integrand %d on boundary facets: only source (Neumann) terms are supported there", form);
  if (form == GB200_FORM_NEOHOOKEAN_RES) {
    GB_REQUIRE(plan->nfields == 1 && plan->ed.f[0].ncomp == plan->ed.D, GB200_ERR_UNSUPPORTED, "neo-Hookean residual needs one vector field");
    return;
  }
  throw gb::Error(GB200_ERR_UNSUPPORTED, fmt("vector integrand %d is not in the supported set {source, neo-Hookean residual}", form));
}

static void set_params(NumericArgs &a, int form_mat, const double *mp, int nm, int form_vec, const double *vp, int nv) {
  for (double &p : a.params) p = 0.0;
  if (form_mat == GB200_FORM_MASS || form_mat == GB200_FORM_LAPLACIAN) a.params[0] = 1.0;
  for (int i = 0; i < nm && i < 4; i++) a.params[i] = mp[i];
  if (form_vec == GB200_FORM_NEOHOOKEAN_RES)
    for (int i = 0; i < nv && i < 4; i++) a.params[i] = vp[i];
  else
    for (int i = 0; i < nv && i < 4; i++) a.params[4 + i] = vp[i];
  if (form_mat == GB200_FORM_SKELETON) {   // {coef, T kind, w+, w-, U kind, z+, z-}: seven parameters, no vector form on skeleton plans
    GB_REQUIRE(nm >= 7 && form_vec == 0, GB200_ERR_INVALID, "skeleton terms need params {coef, T kind, w+, w-, U kind, z+, z-} and no vector form");
    for (int i = 0; i < 7; i++) a.params[i] = mp[i];
  }
  if (form_mat == GB200_FORM_ELASTICITY || form_mat == GB200_FORM_NEOHOOKEAN_JAC)
    GB_REQUIRE(nm >= 2, GB200_ERR_INVALID, "form %d needs params {lambda, mu}", form_mat);
  if (form_vec == GB200_FORM_NEOHOOKEAN_RES && !form_mat) GB_REQUIRE(nv >= 2, GB200_ERR_INVALID, "neo-Hookean residual needs params {lambda, mu}");
}

static void run_numeric(gb200_plan plan, int form_mat, const double *mp, int nm, int form_vec, const double *vp, int nv, const double *fq,
                        const double *Ke, bool lift, double *nzval, double *b, bool want_mat, bool want_vec, int add_flag) {
  gb200_ctx ctx = plan->ctx;
  cudaStream_t s = ctx->stream;
  if (ctx->pending.size() > 8192) resolve_timings(ctx);  // bound the number of live events in long device-resident loops
  if (want_mat && !Ke && plan->mesh->ncells > 0) ensure_gather_plan(plan);
  NumericArgs a;
  set_params(a, form_mat, mp, nm, form_vec, vp, nv);
  a.form_mat = form_mat;
  a.form_vec = form_vec;
  a.lift = lift;
  DevBuf<double> d_Ke;
  if (Ke) { d_Ke.upload(Ke, (size_t)plan->NL * plan->NL, s); a.Ke_const = d_Ke.p; }
  if (form_vec == GB200_FORM_SOURCE || form_vec == GB200_FORM_FACET_VEC) {
    // one source per field, field after field: constants vp[0..ncomp_0), vp[ncomp_0..) ...; fq likewise [field][cell][p][comp]
    size_t total = 0;
    for (int f = 0; f < plan->nfields; f++) total += (size_t)plan->mesh->ncells * plan->ed.np * plan->ed.f[f].ncomp;
    if (fq) {
      ScopedTimer t(ctx, "h2d_fq");
      plan->fq.upload(fq, total, s);
      a.fq = plan->fq.p;
    }
    int po = 0;
    size_t fo = 0;
    for (int f = 0; f < plan->nfields; f++) {
      FieldDesc &fd = plan->ed.f[f];
      for (int c = 0; c < 3; c++) fd.src[c] = (form_vec == GB200_FORM_SOURCE && c < fd.ncomp && po + c < nv) ? vp[po + c] : 0.0;
      fd.src_fq = fq ? plan->fq.p + fo : nullptr;
      po += fd.ncomp;
      fo += (size_t)plan->mesh->ncells * plan->ed.np * fd.ncomp;
    }
  }
  if (add_flag) {
    // assemble_*_add!: the caller's current values are the starting point
    ScopedTimer t(ctx, "h2d_add");
    if (want_mat && nzval && plan->nnz) GB_CUDA(cudaMemcpyAsync(plan->nzval.p, nzval, plan->nnz * 8, cudaMemcpyHostToDevice, s));
    if (want_vec && b) GB_CUDA(cudaMemcpyAsync(plan->bvec.p, b, plan->nrows * 8, cudaMemcpyHostToDevice, s));
  }
  {
    ScopedTimer t(ctx, "kernels");
    bool gather = want_mat && !Ke && plan->mesh->ncells > 0 && gather_supported(plan, form_mat);
    if (plan->mesh->ncells == 0) {
      if (!add_flag && want_vec) plan->bvec.zero(s);
    } else if (gather) {
      plan->path[form_mat] = gather_mode(plan, form_mat) == 2 ? "q1hex_gather_general" : "q1hex_gather_affine";
      launch_gather(plan, form_mat, a.params, plan->nzval.p, add_flag != 0);
      if (want_vec) {
        if (!add_flag) plan->bvec.zero(s);
        // local vector + lifting: specialised cell kernel (q1hex_rhs.cu), else the generic kernel
        if (!launch_q1hex_rhs(plan, form_vec, lift ? form_mat : 0, a.params, a.fq, plan->bvec.p)) {
          NumericArgs v = a;
          launch_generic(plan, v, nullptr, plan->bvec.p);
        }
      }
    } else if (want_mat && !Ke && launch_affine_gather(plan, form_mat, a.params, plan->nzval.p, add_flag != 0)) {
      // affine cells: owner-computes column-node gather (no atomics, no zero-fill, deterministic); the local vector and the
      // lifting b_e -= K_e u_e by the cell-centric kernels (without a matrix target the generic kernel evaluates only the
      // entries of Dirichlet columns on the cells that touch a Dirichlet DoF)
      plan->path[form_mat] = "affine_gather";   // (+blocks: thread per stored node-pair block; +columns: warp per column node)
      if (want_vec) {
        if (!add_flag) { plan->bvec.zero(s); count_launch(ctx, 1); }
        if (lift || !launch_vector_kernel(plan, 0, form_vec, a.params, a.fq, nullptr, plan->bvec.p)) launch_generic(plan, a, nullptr, plan->bvec.p);
      }
    } else if (want_mat && !Ke && !(want_vec && lift) &&
               launch_staged_gather(plan, form_mat, want_vec ? form_vec : 0, a.params, a.fq, plan->nzval.p, want_vec ? plan->bvec.p : nullptr, add_flag != 0,
                                    want_vec && !add_flag)) {
      // one vector-valued field, any geometry, state-dependent integrands: cell-centric node-pair blocks staged in HBM, summed per stored
      // block by the block-owner gather (no atomics on the matrix, no zero-fill); residual_and_jacobian stays one fused cell kernel
      plan->path[form_mat] = "staged_gather";
    } else {
      if (want_mat) plan->path[form_mat] = ctx->deterministic() ? "generic_coloured" : "generic_atomic";
      if (!add_flag) {
        if (want_mat) plan->nzval.zero(s);
        if (want_vec) plan->bvec.zero(s);
        count_launch(ctx, (want_mat ? 1 : 0) + (want_vec ? 1 : 0));
      }
      // one vector-valued field in 3D: specialised node-pair kernels (vector_kernels.cu) for the matrix and/or the local
      // vector; a fused Dirichlet lifting (needs K_e and b_e together) stays on the generic kernel.  Stokes: the velocity
      // block goes through the vector-Laplacian instance, the coupling blocks through the generic kernel.
      bool fast = false;
      if (!want_mat && want_vec && !Ke) fast = launch_q1hex_rhs(plan, form_vec, 0, a.params, a.fq, plan->bvec.p);
      if (!fast && !Ke && want_mat && want_vec && lift) {
        // AffineFEOperator on a vector-valued / multi-field space: the matrix by the specialised kernel, then the local vector
        // and the lifting b_e -= K_e u_e by the generic kernel, which without a matrix target evaluates only the entries of
        // Dirichlet columns on the cells that touch a Dirichlet DoF
        if (form_mat == GB200_FORM_STOKES) {
          double lap[8] = {1.0, 0, 0, 0, 0, 0, 0, 0};
          fast = launch_vector_kernel(plan, GB200_FORM_LAPLACIAN, 0, lap, nullptr, plan->nzval.p, nullptr);
        } else {
          fast = launch_vector_kernel(plan, form_mat, 0, a.params, nullptr, plan->nzval.p, nullptr);
        }
        if (fast) {
          plan->path[form_mat] = ctx->deterministic() ? "vector_coloured" : "vector_atomic";
          launch_generic(plan, a, nullptr, plan->bvec.p);
        }
      } else if (!fast && !Ke && !(want_vec && lift)) {
        if (form_mat == GB200_FORM_STOKES && want_mat && !want_vec) {
          double lap[8] = {1.0, 0, 0, 0, 0, 0, 0, 0};
          fast = launch_vector_kernel(plan, GB200_FORM_LAPLACIAN, 0, lap, nullptr, plan->nzval.p, nullptr);
          if (fast) plan->path[form_mat] = ctx->deterministic() ? "vector_coloured" : "vector_atomic";  // all three blocks in one launch
        } else {
          fast = launch_vector_kernel(plan, want_mat ? form_mat : 0, want_vec ? form_vec : 0, a.params, a.fq, want_mat ? plan->nzval.p : nullptr,
                                      want_vec ? plan->bvec.p : nullptr);
          if (!fast && want_mat && want_vec) {  // no fused instance: matrix and vector separately
            bool m = launch_vector_kernel(plan, form_mat, 0, a.params, nullptr, plan->nzval.p, nullptr);
            if (m) {
              NumericArgs v = a;
              v.form_mat = 0;
              if (!launch_vector_kernel(plan, 0, form_vec, a.params, a.fq, nullptr, plan->bvec.p)) launch_generic(plan, v, nullptr, plan->bvec.p);
              fast = true;
            }
          }
          if (fast && want_mat) plan->path[form_mat] = ctx->deterministic() ? "vector_coloured" : "vector_atomic";
        }
      }
      if (fast) {
      } else {
        launch_generic(plan, a, want_mat ? plan->nzval.p : nullptr, want_vec ? plan->bvec.p : nullptr);
      }
    }
  }
  if ((want_mat && nzval && plan->nnz) || (want_vec && b)) {   // (device-resident calls: nothing to copy, no event records either)
    if (ctx->copy_pending && ctx->pattern_copied_pending) {
      // an asynchronous pattern download with host-side widening is in flight: the values follow the Int32 rows over the link
      // instead of sharing it with them, so that the host threads widen while the values arrive
      GB_CUDA(cudaStreamWaitEvent(s, ctx->pattern_copied, 0));
    }
    ScopedTimer t(ctx, "d2h");
    if (want_mat && nzval && plan->nnz) GB_CUDA(cudaMemcpyAsync(nzval, plan->nzval.p, plan->nnz * 8, cudaMemcpyDeviceToHost, s));
    if (want_vec && b) GB_CUDA(cudaMemcpyAsync(b, plan->bvec.p, plan->nrows * 8, cudaMemcpyDeviceToHost, s));
  }
  // Device-resident calls (no host array involved) are asynchronous: consecutive re-assemblies queue back to back on the
  // context stream.  gb200_synchronize / gb200_get_timings / any call with a host array synchronises.
  if ((want_mat && nzval) || (want_vec && b) || fq || Ke || (add_flag && (nzval || b))) {
    GB_CUDA(cudaStreamSynchronize(s));
    sync_copies(ctx);
  }
}

int32_t gb200_assemble_matrix(gb200_plan plan, int32_t form, const double *params, int32_t nparams, double *nzval, int32_t add_flag) {
  if (!plan) return GB200_ERR_INVALID;
  return guarded(plan->ctx, [&] {
    check_matrix_form(plan, form);
    run_numeric(plan, form, params, nparams, 0, nullptr, 0, nullptr, nullptr, false, nzval, nullptr, true, false, add_flag);
  });
}

int32_t gb200_assemble_matrix_const(gb200_plan plan, const double *Ke, double *nzval, int32_t add_flag) {
  if (!plan || !Ke) return GB200_ERR_INVALID;
  return guarded(plan->ctx, [&] {
    run_numeric(plan, 0, nullptr, 0, 0, nullptr, 0, nullptr, Ke, false, nzval, nullptr, true, false, add_flag);
  });
}

int32_t gb200_assemble_vector(gb200_plan plan, int32_t form, const double *params, int32_t nparams, const double *fq, double *b, int32_t add_flag) {
  if (!plan) return GB200_ERR_INVALID;
  return guarded(plan->ctx, [&] {
    check_vector_form(plan, form);
    run_numeric(plan, 0, nullptr, 0, form, params, nparams, fq, nullptr, false, nullptr, b, false, true, add_flag);
  });
}

int32_t gb200_assemble_matrix_and_vector(gb200_plan plan, int32_t form_mat, const double *mat_params, int32_t nmat, int32_t form_vec,
                                         const double *vec_params, int32_t nvec, const double *fq, double *nzval, double *b, int32_t add_flag) {
  if (!plan) return GB200_ERR_INVALID;
  return guarded(plan->ctx, [&] {
    check_matrix_form(plan, form_mat);
    check_vector_form(plan, form_vec);
    // Dirichlet lifting b_e -= K_e u_e belongs to the affine pair (a, l) of AffineFEOperator; a residual already carries the
    // Dirichlet values in u_h (residual_and_jacobian!, src/FESpaces/FEOperatorsFromWeakForm.jl:85-103)
    // (facet-of-cell plans: a Nitsche matrix term is lifted the same way, paired with a zero GB200_FORM_FACET_VEC vector)
    const bool lift = form_vec == GB200_FORM_SOURCE || form_vec == GB200_FORM_FACET_VEC;
    run_numeric(plan, form_mat, mat_params, nmat, form_vec, vec_params, nvec, fq, nullptr, lift, nzval, b, true, true, add_flag);
  });
}

int32_t gb200_quadrature_points(gb200_plan plan, double *xq) {
  if (!plan || !xq) return GB200_ERR_INVALID;
  return guarded(plan->ctx, [&] {
    DevBuf<double> d;
    d.alloc((size_t)plan->mesh->ncells * plan->ed.np * plan->ed.D);
    launch_quadrature_points(plan, d.p);
    d.download(xq, plan->ctx->stream);
    GB_CUDA(cudaStreamSynchronize(plan->ctx->stream));
  });
}

int32_t gb200_plan_device_nzval(gb200_plan plan, void **dptr, int64_t *nnz) {
  if (!plan || !dptr) return GB200_ERR_INVALID;
  *dptr = plan->nzval.p;
  if (nnz) *nnz = plan->nnz;
  return GB200_OK;
}
int32_t gb200_plan_device_pattern(gb200_plan plan, void **colptr, void **rowval) {
  if (!plan || !colptr || !rowval) return GB200_ERR_INVALID;
  *colptr = plan->colptr.p;
  *rowval = plan->rowval.p;
  return GB200_OK;
}
int32_t gb200_plan_device_vector(gb200_plan plan, void **dptr, int64_t *nrows) {
  if (!plan || !dptr) return GB200_ERR_INVALID;
  *dptr = plan->bvec.p;
  if (nrows) *nrows = plan->nrows;
  return GB200_OK;
}
int32_t gb200_plan_download(gb200_plan plan, double *nzval, double *b) {
  if (!plan) return GB200_ERR_INVALID;
  return guarded(plan->ctx, [&] {
    cudaStream_t s = plan->ctx->stream;
    if (nzval && plan->nnz) GB_CUDA(cudaMemcpyAsync(nzval, plan->nzval.p, plan->nnz * 8, cudaMemcpyDeviceToHost, s));
    if (b) GB_CUDA(cudaMemcpyAsync(b, plan->bvec.p, plan->nrows * 8, cudaMemcpyDeviceToHost, s));
    GB_CUDA(cudaStreamSynchronize(s));
    sync_copies(plan->ctx);
  });
}
static void check_block(gb200_plan plan, int bi, int bj) {
  GB_REQUIRE(bi >= 0 && bi < plan->nfields && bj >= 0 && bj < plan->nfields, GB200_ERR_INVALID, "block (%d,%d) out of range (%d fields)", bi, bj, plan->nfields);
}
int32_t gb200_plan_block_nnz(gb200_plan plan, int32_t bi, int32_t bj, int64_t *nnz) {
  if (!plan || !nnz) return GB200_ERR_INVALID;
  return guarded(plan->ctx, [&] {
    check_block(plan, bi, bj);
    *nnz = plan->nnz ? block_layout(plan, bi, bj) : 0;
  });
}
int32_t gb200_plan_get_block_pattern(gb200_plan plan, int32_t bi, int32_t bj, int64_t *colptr, int64_t *rowval) {
  if (!plan || !colptr) return GB200_ERR_INVALID;
  return guarded(plan->ctx, [&] {
    check_block(plan, bi, bj);
    block_to_host(plan, bi, bj, colptr, rowval, nullptr);
  });
}
int32_t gb200_plan_download_block(gb200_plan plan, int32_t bi, int32_t bj, double *nzval) {
  if (!plan || !nzval) return GB200_ERR_INVALID;
  return guarded(plan->ctx, [&] {
    check_block(plan, bi, bj);
    block_to_host(plan, bi, bj, nullptr, nullptr, nzval);
  });
}
int32_t gb200_plan_get_csr_pattern(gb200_plan plan, int32_t index_base, int64_t *rowptr, int64_t *colval) {
  if (!plan || !rowptr || (!colval && plan->nnz)) return GB200_ERR_INVALID;
  return guarded(plan->ctx, [&] {
    GB_REQUIRE(index_base == 0 || index_base == 1, GB200_ERR_INVALID, "SparseMatrixCSR{Bi}: Bi must be 0 or 1, got %d", index_base);
    csr_to_host(plan, index_base, rowptr, colval, nullptr);
  });
}
int32_t gb200_plan_download_csr(gb200_plan plan, double *nzval) {
  if (!plan || (!nzval && plan->nnz)) return GB200_ERR_INVALID;
  return guarded(plan->ctx, [&] { csr_to_host(plan, 0, nullptr, nullptr, nzval); });
}
int32_t gb200_plan_add_matrix_from(gb200_plan dst, gb200_plan src) {
  if (!dst || !src) return GB200_ERR_INVALID;
  return guarded(dst->ctx, [&] {
    GB_REQUIRE(dst->ctx == src->ctx && dst->nrows == src->nrows && dst->ncols == src->ncols, GB200_ERR_INVALID,
               "plans of different contexts / global systems");
    add_matrix_from(dst, src);
  });
}
int32_t gb200_plan_upload_vector(gb200_plan plan, const double *b) {
  if (!plan || !b) return GB200_ERR_INVALID;
  return guarded(plan->ctx, [&] {
    GB_CUDA(cudaMemcpyAsync(plan->bvec.p, b, (size_t)plan->nrows * 8, cudaMemcpyHostToDevice, plan->ctx->stream));
    GB_CUDA(cudaStreamSynchronize(plan->ctx->stream));
  });
}
int32_t gb200_plan_fold_constraints(gb200_plan dst, gb200_plan src, const int64_t *dof_ptrs, const int32_t *dof_mdofs, const double *dof_coeffs,
                                    const double *dirichlet_master_values, int64_t ndirichlet_masters, int32_t with_matrix, int32_t with_vector) {
  if (!dst || !src || !dof_ptrs || !dof_mdofs || !dof_coeffs) return GB200_ERR_INVALID;
  return guarded(dst->ctx, [&] {
    GB_REQUIRE(dst->ctx == src->ctx && src->nrows == src->ncols, GB200_ERR_INVALID,
               "the source plan must be the square system of the unconstrained space (free and Dirichlet DoFs in one numbering)");
    GB_REQUIRE(dof_ptrs[0] == 1, GB200_ERR_INVALID, "dof_ptrs must start at 1");
    const int64_t nd = dof_ptrs[src->ncols] - 1;
    for (int64_t q = 0; q < nd; q++)
      GB_REQUIRE(dof_mdofs[q] != 0 && dof_mdofs[q] <= dst->ncols && dof_mdofs[q] <= dst->nrows && -(int64_t)dof_mdofs[q] <= std::max<int64_t>(ndirichlet_masters, 0),
                 GB200_ERR_INVALID, "master DoF %d out of range", dof_mdofs[q]);
    fold_constraints(dst, src, dof_ptrs, dof_mdofs, dof_coeffs, dirichlet_master_values, ndirichlet_masters, with_matrix != 0, with_vector != 0);
  });
}
int32_t gb200_owned_column_ids(const int32_t *ids, int64_t n, const uint8_t *owned, int64_t nfree, int32_t *out, int64_t *n_owned,
                               int64_t *owned_ids) {
  if (!ids || !owned || !out || n < 0 || nfree < 0) return GB200_ERR_INVALID;
  return guarded(nullptr, [&] {
    std::vector<int32_t> local((size_t)nfree + 1, 0);
    int64_t cnt = 0;
    for (int64_t j = 0; j < nfree; j++)
      if (owned[j]) {
        local[(size_t)j + 1] = (int32_t)(++cnt);
        if (owned_ids) owned_ids[cnt - 1] = j + 1;
      }
    for (int64_t t = 0; t < n; t++) {
      const int32_t id = ids[t];
      GB_REQUIRE(id <= nfree, GB200_ERR_INVALID, "DoF id %d exceeds the number of free DoFs %lld", id, (long long)nfree);
      out[t] = id > 0 ? local[(size_t)id] : id;
    }
    if (n_owned) *n_owned = cnt;
  });
}

const char *gb200_plan_kernel_path(gb200_plan plan, int32_t form) {
  if (!plan) return "";
  auto it = plan->path.find(form);
  if (it == plan->path.end()) return "";
  auto d = plan->path_detail.find(form);
  if (d == plan->path_detail.end()) return it->second.c_str();
  plan->path_full[form] = it->second + "+" + d->second;  // e.g. "vector_atomic+dmma"
  return plan->path_full[form].c_str();
}

}  // extern "C"
