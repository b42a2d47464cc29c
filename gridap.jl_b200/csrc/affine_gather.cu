// affine_gather.cu -- owner-computes, atomic-free assembly for Lagrangian elements of any order on AFFINE cells in 3D
// (simplices always; hexahedra whose vertices form a parallelepiped, e.g. every Cartesian mesh): vector- / scalar-valued
// single fields (mass, Laplacian, linear elasticity) and the Stokes Taylor-Hood pair (BASELINE.json configs 3 and 4).
//
// Reference work being replaced per cell (a4-a8, a13 of SURVEY.md section 8): Jt / inv / det at every quadrature point, physical
// gradients, the integrand broadcast aq[p,i,j] and the IntegrationMap contraction (src/Fields/FieldArrays.jl:675-757,
// src/Fields/FieldsInterfaces.jl:737-760), then ni*nj binary-search insertions (src/Algebra/SparseMatrixCSC.jl:124-150).
//
// On an affine cell inv(Jt) =: I and |det Jt| are constant, so the quadrature collapses onto reference tensors that are
// tabulated once per plan from the quadrature the caller passed (exact restatement, not an approximation):
//     A_ab[i][j] = sum_p dV_p d_i phi_a d_j phi_b = |det| sum_mn I[i][m] I[j][n] M^{mn}_ab ,  M^{mn}_ab = sum_p w_p d_m N_a d_n N_b
//     mass_ab    = |det| sum_p w_p N_a N_b ,      T_aq[c] = sum_p dV_p d_c phi_a psi_q = |det| sum_m I[c][m] C^m_aq
// and every stored entry is a few FMAs of (I, |det|) of its cell -- the generalisation of the closed form of the headline path.
//
// Design (B200): no atomics, no zero-fill, every nnz slot written exactly once with coalesced stores.
//   kernel 1 (cell-parallel):  I (9 doubles) and |det| per cell from the node coordinates.
//   kernel 2 (column-NODE-parallel): a warp owns the columns of one trial node (its <= 3 components: the 3x3 component block of a
//     node pair shares B = |det| I M I^T).  The node's incident (cell, local node) list comes from the plan (ascending cells);
//     for each incident cell the lanes run over the cell's row nodes, evaluate the <= 3x3 block in closed form and add it into
//     the column buffers in shared memory at the in-column ranks of the plan's slot map (distinct rows per lane: no conflicts;
//     cells one after the other: per-slot summation order = ascending cell order = the reference's, bitwise reproducible).
//     The finished columns leave as contiguous, coalesced stores.  Work is handed out in chunks of columns through one atomic
//     counter (vertex nodes of a P2 / Q2 mesh have up to 24 / 8 incident cells, interior nodes one).
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace gb {

namespace {

constexpr int CNG_THREADS = 128;   // 4 warps per CTA
constexpr int CNG_F = 16;          // doubles per cell: I = inv(Jt) (9), |det| (1), G = |det| I^T I (6: 00 11 22 01 02 12)
constexpr int CNG_CHUNK = 64;      // incident (cell, node) entries per work unit (units are dealt round-robin to the persistent warps)

// One record per trial node (leader = its first free component's column): where its incident entries start, how many, which
// columns (free components) it owns.  64 bytes: one coalesced load per node.
struct CngNode {
  int64_t adj_begin;
  int32_t nent, field;
  int64_t cbase[3];     // nzval offset of the column of component c, -1: not stored (Dirichlet / masked)
  int32_t clen[3];
  int32_t pad[3];
};
static_assert(sizeof(CngNode) == 64, "CngNode must be 64 bytes");

struct CngArgs {
  int nfields, NL, nn_tot;            // concatenated local dofs / local nodes of all fields
  int nds[MAX_FIELDS], ncomp[MAX_FIELDS], lofs[MAX_FIELDS], nofs[MAX_FIELDS];
  const int32_t *row_ids[MAX_FIELDS], *col_ids[MAX_FIELDS];
  int64_t col_off[MAX_FIELDS];
  int64_t ncells, ncols;
  const int64_t *colptr;
  const uint16_t *rank;
  const CngNode *nodes;     // [nnodes], ascending leader column
  const int64_t *unit_ptr;  // [nunits + 1]: first node of every work unit
  int64_t nunits, nent;
  const int64_t *adj;       // (cell << 6) | local node, ascending inside a node's list
  const double *F;          // [ncells][CNG_F]: I = inv(Jt) row-major, |det Jt|, G = |det| I^T I
  const double *tab;        // reference tensors
  int o_M, o_mass, o_C;     // offsets into tab: M[(b*9 + mn)*nds0 + a], mass[b*nds0 + a], C[(q*3 + m)*nds0 + a]
  int o_Ms;                 // Ms[(b*6 + s)*nds0 + a]: M00, M11, M22, M01 + M10, M02 + M20, M12 + M21 (what G : M needs)
  double p0, p1;
  double *nzval;
  int add, buf_len;
  const double *Ke;         // staged blocks [ncells][pairs a <= b][9] (FORM_STAGED)
};

__device__ __forceinline__ int field_of_node(const CngArgs &k, int ln) { return (k.nfields > 1 && ln >= k.nofs[1]) ? 1 : 0; }

// ---- plan: trial node -> incident (cell, local node) lists ---------------------------------------------------------------
__device__ __forceinline__ int64_t leader_column(const CngArgs &k, int64_t cell, int ln) {
  const int f = field_of_node(k, ln), a = ln - k.nofs[f];
  const int32_t *ids = k.col_ids[f] + cell * (int64_t)(k.nds[f] * k.ncomp[f]);
  for (int c = 0; c < k.ncomp[f]; c++) {
    const int32_t id = ids[a + k.nds[f] * c];
    if (id > 0) return (int64_t)id - 1 + k.col_off[f];
  }
  return -1;
}

__global__ void cng_count_kernel(CngArgs k, unsigned long long *cnt) {
  const int64_t total = k.ncells * k.nn_tot;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t J = leader_column(k, t / k.nn_tot, (int)(t % k.nn_tot));
    if (J >= 0) atomicAdd(&cnt[J], 1ull);
  }
}

__global__ void cng_flag_kernel(const int64_t *cnt, int64_t ncols, int64_t *flag) {
  for (int64_t J = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; J <= ncols; J += (int64_t)gridDim.x * blockDim.x) flag[J] = (J < ncols && cnt[J] > 0) ? 1 : 0;
}

__global__ void cng_fill_kernel(CngArgs k, const int64_t *ptr, unsigned long long *cursor, int64_t *adj) {
  const int64_t total = k.ncells * k.nn_tot;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t cell = t / k.nn_tot;
    const int ln = (int)(t % k.nn_tot);
    const int64_t J = leader_column(k, cell, ln);
    if (J >= 0) adj[ptr[J] + (int64_t)atomicAdd(&cursor[J], 1ull)] = (cell << 6) | ln;
  }
}

// ascending cells inside every list (= the reference's summation order), the node records, the first node of every work unit
// and the largest column buffer a warp needs
__global__ void cng_sort_kernel(CngArgs k, const int64_t *ptr, const int64_t *node_index, int64_t *adj, CngNode *nodes, int64_t *unit_ptr,
                                unsigned long long *buf_max) {
  for (int64_t J = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; J < k.ncols; J += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = ptr[J], e = ptr[J + 1];
    if (e == b) continue;
    for (int64_t q = b + 1; q < e; q++) {
      const int64_t v = adj[q];
      int64_t m = q - 1;
      while (m >= b && adj[m] > v) { adj[m + 1] = adj[m]; m--; }
      adj[m + 1] = v;
    }
    const int64_t cell = adj[b] >> 6;
    const int ln = (int)(adj[b] & 63);
    const int f = field_of_node(k, ln), a = ln - k.nofs[f];
    const int32_t *ids = k.col_ids[f] + cell * (int64_t)(k.nds[f] * k.ncomp[f]);
    CngNode nd;
    nd.adj_begin = b;
    nd.nent = (int32_t)(e - b);
    nd.field = f;
    unsigned long long len = 0;
    for (int c = 0; c < 3; c++) {
      nd.cbase[c] = -1;
      nd.clen[c] = 0;
      nd.pad[c] = 0;
      if (c >= k.ncomp[f]) continue;
      const int32_t id = ids[a + k.nds[f] * c];
      if (id > 0) {
        const int64_t col = (int64_t)id - 1 + k.col_off[f];
        nd.cbase[c] = k.colptr[col];
        nd.clen[c] = (int32_t)(k.colptr[col + 1] - k.colptr[col]);
        len += (unsigned long long)nd.clen[c];
      }
    }
    const int64_t idx = node_index[J];
    nodes[idx] = nd;
    // work unit u = the nodes whose list begins in [u CNG_CHUNK, (u + 1) CNG_CHUNK): unit_ptr[u] = first node beginning at or after the cut
    for (int64_t u = (b + CNG_CHUNK - 1) / CNG_CHUNK; u * CNG_CHUNK < e; u++) unit_ptr[u] = (u * CNG_CHUNK == b) ? idx : idx + 1;
    atomicMax(buf_max, len);
  }
}

// ---- kernel 1: affine cell factors ----------------------------------------------------------------------------------------
// Jt = sum_a dNg_a(q_0) (x) x_a (constant over an affine cell), I = inv(Jt), |det|
__global__ void cng_factors_kernel(const double *__restrict__ X, const int32_t *__restrict__ cell_nodes, int nn, const double *__restrict__ dNg,
                                   int64_t ncells, double *__restrict__ F, int *not_diagonal) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  double J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (nn == 8) {   // Jt[i][:] = mean of the four edges parallel to axis i (exact zeros off the diagonal for axis-aligned boxes)
    double x[8][3];
    for (int a = 0; a < 8; a++)
      for (int d = 0; d < 3; d++) x[a][d] = X[(int64_t)cell_nodes[c * 8 + a] * 3 + d];
    for (int d = 0; d < 3; d++) {
      J[0 + d] = 0.25 * ((x[1][d] - x[0][d]) + (x[3][d] - x[2][d]) + (x[5][d] - x[4][d]) + (x[7][d] - x[6][d]));
      J[3 + d] = 0.25 * ((x[2][d] - x[0][d]) + (x[3][d] - x[1][d]) + (x[6][d] - x[4][d]) + (x[7][d] - x[5][d]));
      J[6 + d] = 0.25 * ((x[4][d] - x[0][d]) + (x[5][d] - x[1][d]) + (x[6][d] - x[2][d]) + (x[7][d] - x[3][d]));
    }
  } else if (nn == 4) {   // Jt[i][:] = x_{i+1} - x_0
    const double *x0 = X + (int64_t)cell_nodes[c * 4] * 3;
    for (int i = 0; i < 3; i++) {
      const double *xi = X + (int64_t)cell_nodes[c * 4 + i + 1] * 3;
      for (int d = 0; d < 3; d++) J[i * 3 + d] = xi[d] - x0[d];
    }
  } else {
    for (int a = 0; a < nn; a++) {
      const double *x = X + (int64_t)cell_nodes[c * nn + a] * 3;
      const double *dn = dNg + a * 3;   // first quadrature point
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) J[i * 3 + j] += dn[i] * x[j];
    }
  }
  const double det = J[0] * J[4] * J[8] + J[1] * J[5] * J[6] + J[2] * J[3] * J[7] - (J[0] * J[5] * J[7] + J[1] * J[3] * J[8] + J[2] * J[4] * J[6]);
  const double ci = 1.0 / det;
  double *o = F + c * CNG_F;
  o[0] = (J[4] * J[8] - J[5] * J[7]) * ci;
  o[1] = -(J[1] * J[8] - J[2] * J[7]) * ci;
  o[2] = (J[1] * J[5] - J[2] * J[4]) * ci;
  o[3] = -(J[3] * J[8] - J[5] * J[6]) * ci;
  o[4] = (J[0] * J[8] - J[2] * J[6]) * ci;
  o[5] = -(J[0] * J[5] - J[2] * J[3]) * ci;
  o[6] = (J[3] * J[7] - J[4] * J[6]) * ci;
  o[7] = -(J[0] * J[7] - J[1] * J[6]) * ci;
  o[8] = (J[0] * J[4] - J[1] * J[3]) * ci;
  const double ad = fabs(det);
  o[9] = ad;
  // tr(|det| I M I^T) = sum_mn G_mn M_mn : the Laplacian-type blocks need only G
  o[10] = ad * (o[0] * o[0] + o[3] * o[3] + o[6] * o[6]);
  o[11] = ad * (o[1] * o[1] + o[4] * o[4] + o[7] * o[7]);
  o[12] = ad * (o[2] * o[2] + o[5] * o[5] + o[8] * o[8]);
  o[13] = ad * (o[0] * o[1] + o[3] * o[4] + o[6] * o[7]);
  o[14] = ad * (o[0] * o[2] + o[3] * o[5] + o[6] * o[8]);
  o[15] = ad * (o[1] * o[2] + o[4] * o[5] + o[7] * o[8]);
  if (not_diagonal && (o[1] != 0.0 || o[2] != 0.0 || o[3] != 0.0 || o[5] != 0.0 || o[6] != 0.0 || o[7] != 0.0)) *not_diagonal = 1;   // (benign race: all writers store 1)
}

// ---- kernel 2: column-node gather -------------------------------------------------------------------------------------------
// Element sizes are template parameters (N0 nodes x C0 components on field 0, N1 scalar nodes on field 1 or 0): every index
// computation folds at compile time (the run-time-sized first version spent 750 instructions per incident entry on them).
// A warp evaluates EPW = 32 / LPE incident entries of its node at a time (LPE lanes per entry = row nodes rounded up to a power of
// two: 1 entry for Q2, 2 for the 14 nodes of P2/P1, 4 for Q1) and adds them to the column buffers one entry after the other, so the
// per-slot summation order stays the ascending cell order.
template <int N0, int C0, int N1>
struct CngShape {
  static constexpr int NN = N0 + N1, NL = N0 * C0 + N1;
  static constexpr int LPE = NN <= 4 ? 4 : NN <= 8 ? 8 : NN <= 16 ? 16 : 32;
  static constexpr int EPW = 32 / LPE;
};

// the <= 3x3 component block of (row node r, column node b of field FC) of one cell
template <bool SMEM>
__device__ __forceinline__ double tab_load(const double *p) { return SMEM ? *p : __ldg(p); }

template <int FORM, int N0, int C0, int N1, int FC, bool SMEM = false>
__device__ __forceinline__ void node_block(const CngArgs &k, const double *tab, const double *__restrict__ I, double det, int r, int b, double *K) {
#pragma unroll
  for (int q = 0; q < 9; q++) K[q] = 0.0;
  const bool row0 = N1 == 0 || r < N0;
  const int a = row0 ? r : r - N0;
  if (row0 && FC == 0) {
    if (FORM == GB200_FORM_MASS) {
      const double m = k.p0 * det * tab_load<SMEM>(tab + k.o_mass + b * N0 + a);
      K[0] = K[4] = K[8] = m;
    } else {
      double m[9];
      const double *M = tab + k.o_M + b * 9 * N0 + a;
#pragma unroll
      for (int q = 0; q < 9; q++) m[q] = tab_load<SMEM>(M + q * N0);
      double T[9], B[9];   // B = |det| I M I^T : A_ab[i][j] of the file header
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int n = 0; n < 3; n++) T[i * 3 + n] = I[i * 3 + 0] * m[0 * 3 + n] + I[i * 3 + 1] * m[1 * 3 + n] + I[i * 3 + 2] * m[2 * 3 + n];
      if (FORM == GB200_FORM_ELASTICITY) {
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 3; j++) B[i * 3 + j] = det * (T[i * 3 + 0] * I[j * 3 + 0] + T[i * 3 + 1] * I[j * 3 + 1] + T[i * 3 + 2] * I[j * 3 + 2]);
        const double mtr = k.p1 * (B[0] + B[4] + B[8]);
#pragma unroll
        for (int ci = 0; ci < 3; ci++)
#pragma unroll
          for (int cj = 0; cj < 3; cj++) K[ci * 3 + cj] = k.p0 * B[ci * 3 + cj] + k.p1 * B[cj * 3 + ci] + (ci == cj ? mtr : 0.0);
      } else {   // Laplacian / Stokes velocity block: equal components only, tr(B)
        double tr = 0.0;
#pragma unroll
        for (int i = 0; i < 3; i++) tr += T[i * 3 + 0] * I[i * 3 + 0] + T[i * 3 + 1] * I[i * 3 + 1] + T[i * 3 + 2] * I[i * 3 + 2];
        const double v = (FORM == GB200_FORM_STOKES ? 1.0 : k.p0) * det * tr;
        K[0] = K[4] = K[8] = v;
      }
    }
  } else if (FORM == GB200_FORM_STOKES && row0 && FC == 1) {   // (v_a,ci | p_b)
    const double *C = tab + k.o_C + b * 3 * N0 + a;
    const double c0 = tab_load<SMEM>(C), c1 = tab_load<SMEM>(C + N0), c2 = tab_load<SMEM>(C + 2 * N0);
#pragma unroll
    for (int ci = 0; ci < 3; ci++) K[ci * 3 + 0] = -det * (I[ci * 3 + 0] * c0 + I[ci * 3 + 1] * c1 + I[ci * 3 + 2] * c2);
  } else if (FORM == GB200_FORM_STOKES && !row0 && FC == 0) {  // (q_a | u_b,cj)
    const double *C = tab + k.o_C + a * 3 * N0 + b;
    const double c0 = tab_load<SMEM>(C), c1 = tab_load<SMEM>(C + N0), c2 = tab_load<SMEM>(C + 2 * N0);
#pragma unroll
    for (int cj = 0; cj < 3; cj++) K[0 * 3 + cj] = det * (I[cj * 3 + 0] * c0 + I[cj * 3 + 1] * c1 + I[cj * 3 + 2] * c2);
  }
}

// what a lane needs of one incident (cell, column node) entry: the cell's factors and the in-column ranks of its row node
struct CngItem {
  double F[10];
  unsigned rr[9];   // [cj][ci]
  int b;
};

template <int N0, int C0, int N1>
__device__ __forceinline__ void load_item(const CngArgs &k, int64_t ent, int fc, int r, bool valid, CngItem &it) {
  using S = CngShape<N0, C0, N1>;
  const int64_t cell = ent >> 6;
  const int ln = (int)(ent & 63);
  it.b = fc ? ln - N0 : ln;
#pragma unroll
  for (int q = 0; q < 9; q++) it.rr[q] = 0xFFFFu;
  if (!valid) return;
  const double2 *Fc = reinterpret_cast<const double2 *>(k.F + cell * CNG_F);   // I and |det|: the first 80 of the cell's 128 bytes
#pragma unroll
  for (int q = 0; q < 5; q++) {
    const double2 v = __ldg(Fc + q);
    it.F[2 * q] = v.x;
    it.F[2 * q + 1] = v.y;
  }
  if (r < S::NN) {
    const bool row0 = N1 == 0 || r < N0;
    const int lofs_r = row0 ? 0 : N0 * C0, a = row0 ? r : r - N0, n_r = row0 ? N0 : N1, c_r = row0 ? C0 : 1;
    const int lofs_c = fc ? N0 * C0 : 0, n_c = fc ? N1 : N0, c_c = fc ? 1 : C0;
    const uint16_t *rk = k.rank + cell * (int64_t)(S::NL * S::NL) + (lofs_c + it.b) * S::NL + lofs_r + a;
#pragma unroll
    for (int cj = 0; cj < 3; cj++)
#pragma unroll
      for (int ci = 0; ci < 3; ci++)
        if (cj < c_c && ci < c_r) it.rr[cj * 3 + ci] = (unsigned)__ldg(rk + n_c * cj * S::NL + n_r * ci);
  }
}

template <int FORM, int N0, int C0, int N1>
__global__ void __launch_bounds__(CNG_THREADS, 4) cng_gather_kernel(CngArgs k) {
  using S = CngShape<N0, C0, N1>;
  constexpr int LPE = S::LPE, EPW = S::EPW;
  constexpr bool DIAG = FORM != GB200_FORM_ELASTICITY;   // the (field 0, field 0) block couples equal components only
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, sub = lane / LPE, r = lane % LPE;
  double *buf = smem + (size_t)(threadIdx.x >> 5) * k.buf_len;
  const int64_t wstride = (int64_t)gridDim.x * (CNG_THREADS / 32);
  for (int64_t u = (int64_t)blockIdx.x * (CNG_THREADS / 32) + (threadIdx.x >> 5); u < k.nunits; u += wstride) {
    const int64_t n0 = k.unit_ptr[u], n1 = k.unit_ptr[u + 1];
    if (n1 <= n0) continue;
    // software pipeline over the steps (EPW incident entries each) of the unit's nodes: while a step is evaluated, the factors /
    // ranks of the next step (of this node or the next one) and the record of the next node are already on their way
    const long long *recp = reinterpret_cast<const long long *>(k.nodes + n0);
    long long w = lane < 8 ? __ldg(recp + lane) : 0;   // node record: 8 x 8 bytes, one coalesced load
    int64_t e = __shfl_sync(0xffffffffu, w, 0);
    CngItem nxt;
    {
      const long long w1 = __shfl_sync(0xffffffffu, w, 1);
      const int nent0 = (int)(w1 & 0xffffffffll), fc0 = (int)(w1 >> 32);
      const bool valid = sub < nent0;
      load_item<N0, C0, N1>(k, valid ? __ldg(k.adj + e + sub) : 0, fc0, r, valid, nxt);
    }
    for (int64_t n = n0; n < n1; n++) {
      const long long wcur = w;
      if (n + 1 < n1) w = lane < 8 ? __ldg(recp + (n + 1 - n0) * 8 + lane) : 0;
      const long long w1 = __shfl_sync(0xffffffffu, wcur, 1);
      const int nent = (int)(w1 & 0xffffffffll), fc = (int)(w1 >> 32);
      const int64_t cbase[3] = {__shfl_sync(0xffffffffu, wcur, 2), __shfl_sync(0xffffffffu, wcur, 3), __shfl_sync(0xffffffffu, wcur, 4)};
      const long long w5 = __shfl_sync(0xffffffffu, wcur, 5), w6 = __shfl_sync(0xffffffffu, wcur, 6);
      const int clen[3] = {(int)(w5 & 0xffffffffll), (int)(w5 >> 32), (int)(w6 & 0xffffffffll)};
      const int cofs[3] = {0, clen[0], clen[0] + clen[1]};
      const int total = clen[0] + clen[1] + clen[2];
      int nent_after = 0, fc_after = 0;
      if (n + 1 < n1) {
        const long long a1 = __shfl_sync(0xffffffffu, w, 1);
        nent_after = (int)(a1 & 0xffffffffll);
        fc_after = (int)(a1 >> 32);
      }
      for (int q = lane; q < total; q += 32) buf[q] = 0.0;
      __syncwarp();
      for (int i = 0; i < nent; i += EPW) {
        const CngItem cur = nxt;
        const bool cur_valid = i + sub < nent;
        {   // next step: the following entries of this node, or the first ones of the next node
          const bool same = i + EPW < nent;
          const int64_t eb = same ? e + i + EPW : e + nent;
          const bool valid = same ? (i + EPW + sub < nent) : (sub < nent_after);
          load_item<N0, C0, N1>(k, valid ? __ldg(k.adj + eb + sub) : 0, same ? fc : fc_after, r, valid, nxt);
        }
        double K[9];
        if (cur_valid && r < S::NN) {
          if (N1 > 0 && fc == 1) node_block<FORM, N0, C0, N1, 1>(k, k.tab, cur.F, cur.F[9], r, cur.b, K);
          else node_block<FORM, N0, C0, N1, 0>(k, k.tab, cur.F, cur.F[9], r, cur.b, K);
        }
        // add at the in-column ranks of the slot map, one entry after the other (ascending cells); a structural zero of the local
        // matrix leaves the (zeroed) slot untouched
        const bool row0 = N1 == 0 || r < N0;
        const bool diag = DIAG && row0 && fc == 0;
#pragma unroll
        for (int sidx = 0; sidx < EPW; sidx++) {
          if (sub == sidx && cur_valid && r < S::NN) {
            // the (at most 9) slots of a lane are distinct: all loads first, then the adds and stores (no serialised read-modify-write chain)
            double old[9];
            bool on[9];
#pragma unroll
            for (int cj = 0; cj < 3; cj++)
#pragma unroll
              for (int ci = 0; ci < 3; ci++) {
                const unsigned q = cur.rr[cj * 3 + ci];
                on[cj * 3 + ci] = clen[cj] != 0 && !(diag && ci != cj) && q != 0xFFFFu;
                old[cj * 3 + ci] = on[cj * 3 + ci] ? buf[cofs[cj] + q] : 0.0;
              }
#pragma unroll
            for (int cj = 0; cj < 3; cj++)
#pragma unroll
              for (int ci = 0; ci < 3; ci++)
                if (on[cj * 3 + ci]) buf[cofs[cj] + cur.rr[cj * 3 + ci]] = old[cj * 3 + ci] + K[ci * 3 + cj];
          }
          __syncwarp();
        }
      }
      e += nent;
      // the finished columns: contiguous, coalesced stores (every slot exactly once)
#pragma unroll
      for (int c = 0; c < 3; c++) {
        if (clen[c] == 0) continue;
        double *out = k.nzval + cbase[c];
        const double *src = buf + cofs[c];
        if (k.add)
          for (int q = lane; q < clen[c]; q += 32) out[q] += src[q];
        else
          for (int q = lane; q < clen[c]; q += 32) out[q] = src[q];
      }
      __syncwarp();
    }
  }
}

// ---- block-owner gather ------------------------------------------------------------------------------------------------------
// The same closed form with one THREAD per stored node-pair block (row node x column node, <= 3x3 entries): the thread loops over
// the block's source cells (ascending), accumulates in registers and writes its entries straight to nzval -- no shared-memory
// accumulation, no synchronisation, all 32 lanes busy whatever the element (the column-node kernel above spends ~500 warp
// instructions per incident entry on 14 - 27 active lanes).  The plan adds, per block, its source list {cell, row node, column node}
// (8 bytes per source) and a 16-byte record; it needs every block to be "regular": the stored components of the row node occupy
// consecutive slots at the same in-column rank in every column of the column node (true for the node-blocked numbering of
// Lagrangian spaces; an AssemblyStrategy that permutes rows falls back to the column-node kernel).
struct BogBlock {
  uint32_t node;        // index of the column node (CngNode)
  uint16_t r0;          // in-column rank of the first stored component of the row node
  uint8_t info;         // bits 0-1: first stored row component, bits 2-3: number of stored row components, bit 4: row node on field 1
  uint8_t nsrc;
  int64_t src_begin;
};
static_assert(sizeof(BogBlock) == 16, "BogBlock must be 16 bytes");
constexpr int BOG_MAXKEYS = 1024;   // incident cells x row nodes of one column node
constexpr int BOG_THREADS = 256;

// keys of one column node in shared memory: (r0 << 16) | (entry << 6) | row node, invalid = 0xFFFFFFFF; returns the padded length
template <int N0, int C0, int N1>
__device__ __forceinline__ int bog_sorted_keys(const CngArgs &k, int64_t e0, int nent, int fc, int cj0, uint32_t *keys, int lane) {
  using S = CngShape<N0, C0, N1>;
  const int nitems = nent * S::NN;
  int P = 32;
  while (P < nitems) P <<= 1;
  for (int idx = lane; idx < P; idx += 32) {
    uint32_t key = 0xFFFFFFFFu;
    if (idx < nitems) {
      const int i = idx / S::NN, la = idx - i * S::NN;
      const int64_t ent = k.adj[e0 + i];
      const int64_t cell = ent >> 6;
      const int b = (int)(ent & 63) - (fc ? N0 : 0);
      const bool row0 = N1 == 0 || la < N0;
      const int lofs_r = row0 ? 0 : N0 * C0, a = row0 ? la : la - N0, n_r = row0 ? N0 : N1, c_r = row0 ? C0 : 1;
      const int lj = (fc ? N0 * C0 : 0) + b + (fc ? N1 : N0) * cj0;
      const uint16_t *rk = k.rank + cell * (int64_t)(S::NL * S::NL) + lj * S::NL + lofs_r + a;
      for (int ci = 0; ci < c_r; ci++) {
        const unsigned q = rk[n_r * ci];
        if (q != 0xFFFFu) { key = (q << 16) | ((unsigned)i << 6) | (unsigned)la; break; }
      }
    }
    keys[idx] = key;
  }
  __syncwarp();
  for (int size = 2; size <= P; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int q = lane; q < (P >> 1); q += 32) {
        const int lo = 2 * q - (q & (stride - 1)), hi = lo + stride;
        const bool up = (lo & size) == 0;
        const uint32_t x = keys[lo], y = keys[hi];
        if ((x > y) == up) { keys[lo] = y; keys[hi] = x; }
      }
      __syncwarp();
    }
  return P;
}

// pass 1 (blocks == nullptr): sources / blocks per node;  pass 2: the source lists and block records
template <int N0, int C0, int N1>
__global__ void __launch_bounds__(128) bog_plan_kernel(CngArgs k, int64_t nnodes, int64_t *nsrc_node, int64_t *nblk_node, const int64_t *src_ptr,
                                                       const int64_t *blk_ptr, int64_t *src, BogBlock *blocks, int *error) {
  using S = CngShape<N0, C0, N1>;
  __shared__ uint32_t s_keys[4][BOG_MAXKEYS];
  __shared__ uint16_t s_start[4][BOG_MAXKEYS + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t *keys = s_keys[warp];
  uint16_t *starts = s_start[warp];
  for (int64_t n = blockIdx.x * 4ll + warp; n < nnodes; n += (int64_t)gridDim.x * 4) {
    const CngNode nd = k.nodes[n];
    const int nent = nd.nent, fc = nd.field;
    if (nent * S::NN > BOG_MAXKEYS || nent > 1023) {
      if (lane == 0) atomicExch(error, 1);
      if (!blocks && lane == 0) { nsrc_node[n] = 0; nblk_node[n] = 0; }
      continue;
    }
    int cj0 = 0;
    while (cj0 < 2 && nd.clen[cj0] == 0) cj0++;
    const int P = bog_sorted_keys<N0, C0, N1>(k, nd.adj_begin, nent, fc, cj0, keys, lane);
    // block starts: positions where r0 changes (valid keys only)
    int nvalid = 0, nblk = 0;
    for (int q0 = 0; q0 < P; q0 += 32) {
      const int q = q0 + lane;
      const uint32_t v = keys[q];
      const bool valid = v != 0xFFFFFFFFu;
      const bool start = valid && (q == 0 || (keys[q - 1] >> 16) != (v >> 16));
      const unsigned mv = __ballot_sync(0xffffffffu, valid), ms = __ballot_sync(0xffffffffu, start);
      if (start) starts[nblk + __popc(ms & ((1u << lane) - 1u))] = (uint16_t)q;
      nvalid += __popc(mv);
      nblk += __popc(ms);
    }
    if (lane == 0) starts[nblk] = (uint16_t)nvalid;
    __syncwarp();
    if (!blocks) {
      if (lane == 0) { nsrc_node[n] = nvalid; nblk_node[n] = nblk; }
      __syncwarp();
      continue;
    }
    const int64_t sb = src_ptr[n], bb = blk_ptr[n];
    for (int q = lane; q < nvalid; q += 32) {
      const uint32_t v = keys[q];
      const int i = (v >> 6) & 1023, la = v & 63;
      const int64_t ent = k.adj[nd.adj_begin + i];
      src[sb + q] = ((ent >> 6) << 12) | ((int64_t)la << 6) | (ent & 63);
    }
    for (int j = lane; j < nblk; j += 32) {
      const int q = starts[j], cnt = starts[j + 1] - q;
      const uint32_t v = keys[q];
      const unsigned r0 = v >> 16;
      const int i = (v >> 6) & 1023, la = v & 63;
      const int64_t ent = k.adj[nd.adj_begin + i];
      const int64_t cell = ent >> 6;
      const int b = (int)(ent & 63) - (fc ? N0 : 0);
      const bool row0 = N1 == 0 || la < N0;
      const int lofs_r = row0 ? 0 : N0 * C0, a = row0 ? la : la - N0, n_r = row0 ? N0 : N1, c_r = row0 ? C0 : 1;
      const int n_c = fc ? N1 : N0, c_c = fc ? 1 : C0;
      // regular: the stored row components are contiguous and sit at r0, r0 + 1, ... in every stored column of the node
      int ci_first = -1, nci = 0;
      bool ok = cnt <= 255;
      for (int cj = 0; cj < c_c; cj++) {
        if (nd.clen[cj] == 0) continue;
        const int lj = (fc ? N0 * C0 : 0) + b + n_c * cj;
        const uint16_t *rk = k.rank + cell * (int64_t)(S::NL * S::NL) + lj * S::NL + lofs_r + a;
        int f = -1;
        unsigned mask = 0;
        for (int ci = 0; ci < c_r; ci++) {
          const unsigned q2 = rk[n_r * ci];
          if (q2 == 0xFFFFu) continue;
          if (f < 0) f = ci;
          mask |= 1u << ci;
          if (q2 != r0 + (unsigned)(ci - f)) ok = false;
        }
        const int cntc = __popc(mask);
        if (f < 0 || mask != (((1u << cntc) - 1u) << f)) ok = false;
        if (ci_first < 0) { ci_first = f; nci = cntc; }
        else if (f != ci_first || cntc != nci) ok = false;
      }
      if (!ok || ci_first < 0) atomicExch(error, 1);
      BogBlock blk;
      blk.node = (uint32_t)n;
      blk.r0 = (uint16_t)r0;
      blk.info = (uint8_t)((ci_first & 3) | ((nci & 3) << 2) | (row0 ? 0 : 16));
      blk.nsrc = (uint8_t)cnt;
      blk.src_begin = sb + q;
      blocks[bb + j] = blk;
    }
    __syncwarp();
  }
}

// ---- mirror pairs ------------------------------------------------------------------------------------------------------------
// All forms of this file are symmetric up to a sign: the block (row node R | column node C) and the block (C | R) hold
// K and +-K^T of the same node pair, summed over the same cells.  When rows and columns share their numbering the plan pairs every
// block t with its mirror image t': the lower index owns the pair, evaluates the sources once and writes both blocks.
constexpr uint32_t BOG_NONE = 0xFFFFFFFFu;

__device__ __forceinline__ int first_stored(const CngNode &nd) {
  int c = 0;
  while (c < 2 && nd.clen[c] == 0) c++;
  return c;
}

// column index of the column that starts at nzval offset `base` (a column with at least one stored entry)
__device__ __forceinline__ int64_t column_at(const int64_t *colptr, int64_t ncols, int64_t base) {
  int64_t lo = 0, hi = ncols;   // largest J with colptr[J] <= base
  while (lo < hi) {
    const int64_t mid = (lo + hi + 1) >> 1;
    if (colptr[mid] <= base) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void bog_col_node_kernel(const CngNode *nodes, int64_t nnodes, const int64_t *colptr, int64_t ncols, int32_t *col_node) {
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < nnodes; n += (int64_t)gridDim.x * blockDim.x) {
    const CngNode nd = nodes[n];
    const int c = first_stored(nd);
    if (nd.clen[c] > 0) col_node[column_at(colptr, ncols, nd.cbase[c])] = (int32_t)n;   // the node's first stored column
  }
}

__global__ void bog_mirror_kernel(const BogBlock *blocks, int64_t nblocks, const CngNode *nodes, const int64_t *blk_ptr, const int64_t *colptr,
                                  const int32_t *rowval, int64_t ncols, const int32_t *col_node, uint32_t *mirror) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nblocks; t += (int64_t)gridDim.x * blockDim.x) {
    const BogBlock b = blocks[t];
    const CngNode nd = nodes[b.node];
    const int c0 = first_stored(nd);
    uint32_t res = BOG_NONE;
    const int64_t g = rowval[nd.cbase[c0] + b.r0];        // row of the block's first stored row component = column of the same DoF
    const int32_t m = (g >= 0 && g < ncols) ? col_node[g] : -1;
    if (m >= 0) {
      // the mirror block sits in the columns of node m at the rank of this node's first stored column (seen as a row)
      const int64_t jn = column_at(colptr, ncols, nd.cbase[c0]);
      int64_t lo = colptr[g], hi = colptr[g + 1];
      const int64_t cb = lo, ce = hi;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (rowval[mid] < jn) lo = mid + 1; else hi = mid;
      }
      if (lo < ce && rowval[lo] == jn) {
        const unsigned r0m = (unsigned)(lo - cb);
        int64_t a = blk_ptr[m], e = blk_ptr[m + 1];
        while (a < e) {
          const int64_t mid = (a + e) >> 1;
          if (blocks[mid].r0 < r0m) a = mid + 1; else e = mid;
        }
        if (a < blk_ptr[m + 1] && blocks[a].r0 == r0m && blocks[a].nsrc == b.nsrc) res = (uint32_t)a;
      }
    }
    mirror[t] = res;
  }
}

// a pairing is usable only if it is an involution: mirror[mirror[t]] == t
__global__ void bog_owner_flag_kernel(const uint32_t *mirror, int64_t nblocks, int64_t *flag) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t <= nblocks; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t f = 0;
    if (t < nblocks) {
      const uint32_t m = mirror[t];
      const bool paired = m != BOG_NONE && mirror[m] == (uint32_t)t;
      f = (!paired || (uint32_t)t <= m) ? 1 : 0;
    }
    flag[t] = f;
  }
}

__global__ void bog_pairs_kernel(const uint32_t *mirror, int64_t nblocks, const int64_t *pos, uint32_t *pairs) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nblocks; t += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t m = mirror[t];
    const bool paired = m != BOG_NONE && mirror[m] == (uint32_t)t;
    if (!paired || (uint32_t)t <= m) {
      pairs[2 * pos[t]] = (uint32_t)t;
      pairs[2 * pos[t] + 1] = (paired && m != (uint32_t)t) ? m : BOG_NONE;
    }
  }
}

// one source (cell, row node la, column node lb) of a block of type bt, added to what the block accumulates (LINEAR in the local
// matrix; ascending cells = the reference's summation order):
//   type 0 (field 0 rows and columns): mass: |det| m_ab;  Laplacian / Stokes: G : M_ab;  elasticity: A = |det| I M_ab I^T (9 entries)
//   type 1 (field 0 rows | pressure column), type 2 (pressure row | field 0 columns): T[c] = |det| (I C_aq)[c]
// DJ: inv(Jt) of every cell is diagonal (axis-aligned boxes; checked on the factors): A_ij = (|det| I_ii I_jj) M_ij
template <int FORM, int N0, int C0, int N1, bool DJ>
__device__ __forceinline__ void bog_source(const CngArgs &k, const double *s_tab, int64_t e, int bt, double *acc) {
  const int64_t cell = e >> 12;
  const int la = (int)(e >> 6) & 63, lb = (int)e & 63;
  if (FORM == FORM_STAGED) {   // the block of the pair (min, max), transposed when the row node is the larger one
    const int a = min(la, lb), b = max(la, lb);
    const double *Kp = k.Ke + (cell * (N0 * (N0 + 1) / 2) + (b * (b + 1) / 2 + a)) * 9;
    double v[9];
#pragma unroll
    for (int q = 0; q < 9; q++) v[q] = __ldg(Kp + q);
    if (la <= lb) {
#pragma unroll
      for (int q = 0; q < 9; q++) acc[q] += v[q];
    } else {
#pragma unroll
      for (int ci = 0; ci < 3; ci++)
#pragma unroll
        for (int cj = 0; cj < 3; cj++) acc[ci * 3 + cj] += v[cj * 3 + ci];
    }
    return;
  }
  const double *Fc = k.F + cell * CNG_F;
  if (bt == 0) {
    const int a = la, b = lb;
    if (FORM == GB200_FORM_MASS) {
      acc[0] += __ldg(Fc + 9) * s_tab[k.o_mass + b * N0 + a];
    } else if (FORM == GB200_FORM_ELASTICITY) {
      const double *M = s_tab + k.o_M + b * 9 * N0 + a;
      if (DJ) {
        const double det = __ldg(Fc + 9), i0 = __ldg(Fc), i1 = __ldg(Fc + 4), i2 = __ldg(Fc + 8);
        const double d0 = det * i0, d1 = det * i1, d2 = det * i2;
        const double s00 = d0 * i0, s01 = d0 * i1, s02 = d0 * i2, s11 = d1 * i1, s12 = d1 * i2, s22 = d2 * i2;
        acc[0] += s00 * M[0 * N0]; acc[1] += s01 * M[1 * N0]; acc[2] += s02 * M[2 * N0];
        acc[3] += s01 * M[3 * N0]; acc[4] += s11 * M[4 * N0]; acc[5] += s12 * M[5 * N0];
        acc[6] += s02 * M[6 * N0]; acc[7] += s12 * M[7 * N0]; acc[8] += s22 * M[8 * N0];
      } else {
        double I[9], m[9], T[9];
        const double2 *F2 = reinterpret_cast<const double2 *>(Fc);
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const double2 v = __ldg(F2 + q);
          I[2 * q] = v.x;
          I[2 * q + 1] = v.y;
        }
        const double2 v4 = __ldg(F2 + 4);
        I[8] = v4.x;
        const double det = v4.y;
#pragma unroll
        for (int q = 0; q < 9; q++) m[q] = M[q * N0];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int n = 0; n < 3; n++) T[i * 3 + n] = det * (I[i * 3 + 0] * m[0 * 3 + n] + I[i * 3 + 1] * m[1 * 3 + n] + I[i * 3 + 2] * m[2 * 3 + n]);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 3; j++) acc[i * 3 + j] += T[i * 3 + 0] * I[j * 3 + 0] + T[i * 3 + 1] * I[j * 3 + 1] + T[i * 3 + 2] * I[j * 3 + 2];
      }
    } else {   // Laplacian / Stokes velocity block: G : M
      const double *M = s_tab + k.o_Ms + b * 6 * N0 + a;
      const double2 *G2 = reinterpret_cast<const double2 *>(Fc + 10);
      const double2 g0 = __ldg(G2), g1 = __ldg(G2 + 1), g2 = __ldg(G2 + 2);   // 00 11 | 22 01 | 02 12
      acc[0] += g0.x * M[0 * N0] + g0.y * M[1 * N0] + g1.x * M[2 * N0] + g1.y * M[3 * N0] + g2.x * M[4 * N0] + g2.y * M[5 * N0];
    }
  } else {
    // coupling blocks of Stokes: velocity node av, pressure node q
    const int av = bt == 1 ? la : lb, q = bt == 1 ? lb - N0 : la - N0;
    const double *C = s_tab + k.o_C + q * 3 * N0 + av;
    const double c0 = C[0], c1 = C[N0], c2 = C[2 * N0];
    double I[9];
    const double2 *F2 = reinterpret_cast<const double2 *>(Fc);
#pragma unroll
    for (int w = 0; w < 4; w++) {
      const double2 v = __ldg(F2 + w);
      I[2 * w] = v.x;
      I[2 * w + 1] = v.y;
    }
    const double2 v4 = __ldg(F2 + 4);
    I[8] = v4.x;
    const double det = v4.y;
#pragma unroll
    for (int c = 0; c < 3; c++) acc[c] += det * (I[c * 3 + 0] * c0 + I[c * 3 + 1] * c1 + I[c * 3 + 2] * c2);
  }
}

// A thread owns a block; its first `cap` sources it evaluates itself, the remainder of a long list (the diagonal block of a P2
// vertex has 24 sources next to neighbours with 2 - 6) is spread over the lanes of the warp and reduced in a fixed shuffle tree,
// so that the warp does not idle behind one lane.  Deterministic (fixed order), identical from run to run.
template <int FORM, int N0, int C0, int N1, bool DJ>
__global__ void __launch_bounds__(BOG_THREADS, FORM == GB200_FORM_ELASTICITY ? (DJ ? 3 : 2) : 4) bog_gather_kernel(CngArgs k, const BogBlock *__restrict__ blocks, const int64_t *__restrict__ src, int64_t nblocks,
                                                                  int tab_off, int tab_len, int cap, const uint32_t *__restrict__ pairs) {
  // the part of the reference tensors this form reads ([tab_off, tab_off + tab_len) of the plan's table) in shared memory
  extern __shared__ double s_stage[];
  for (int i = threadIdx.x; i < tab_len; i += BOG_THREADS) s_stage[i] = k.tab[tab_off + i];
  __syncthreads();
  const double *s_tab = s_stage - tab_off;
  const int lane = threadIdx.x & 31;
  constexpr int NRED = (FORM == GB200_FORM_ELASTICITY || FORM == FORM_STAGED) ? 9 : (N1 > 0 ? 3 : 1);   // accumulators in use
  for (int64_t t0 = blockIdx.x * (int64_t)BOG_THREADS + (threadIdx.x & ~31); t0 < nblocks; t0 += (int64_t)gridDim.x * BOG_THREADS) {
    const int64_t t = t0 + lane;
    const bool live = t < nblocks;   // (nblocks = number of owned pairs when `pairs` is given)
    int4 raw = make_int4(0, 0, 0, 0);
    uint32_t tmirror = BOG_NONE;
    if (live) {
      int64_t tb = t;
      if (pairs) {
        const uint2 pr = __ldg(reinterpret_cast<const uint2 *>(pairs) + t);
        tb = pr.x;
        tmirror = pr.y;
      }
      raw = __ldg(reinterpret_cast<const int4 *>(blocks + tb));
    }
    const uint32_t node = (uint32_t)raw.x;
    const unsigned r0 = (unsigned)raw.y & 0xFFFFu, info = ((unsigned)raw.y >> 16) & 0xFFu, nsrc = ((unsigned)raw.y >> 24) & 0xFFu;
    const int64_t sb = ((int64_t)(uint32_t)raw.z) | ((int64_t)raw.w << 32);
    const CngNode *nd = k.nodes + node;
    const int fc = (N1 > 0 && live) ? __ldg(&nd->field) : 0;
    // block type: 0 (field 0 rows, field 0 columns), 1 (field 0 rows | pressure column), 2 (pressure row | field 0 columns)
    const int bt = N1 == 0 ? 0 : (fc == 1 ? 1 : ((info & 16) ? 2 : 0));
    double acc[9];
#pragma unroll
    for (int q = 0; q < 9; q++) acc[q] = 0.0;
    const unsigned nown = min(nsrc, (unsigned)cap);
    int64_t e = nown ? __ldg(src + sb) : 0;
    for (unsigned s = 0; s < nown; s++) {
      const int64_t ecur = e;
      if (s + 1 < nown) e = __ldg(src + sb + s + 1);
      bog_source<FORM, N0, C0, N1, DJ>(k, s_tab, ecur, bt, acc);
    }
    unsigned longmask = __ballot_sync(0xffffffffu, nsrc > (unsigned)cap);
    while (longmask) {
      const int owner = __ffs(longmask) - 1;
      longmask &= longmask - 1;
      const int64_t sb_o = __shfl_sync(0xffffffffu, sb, owner);
      const unsigned n_o = __shfl_sync(0xffffffffu, nsrc, owner);
      const int bt_o = __shfl_sync(0xffffffffu, bt, owner);
      double part[9];
#pragma unroll
      for (int q = 0; q < 9; q++) part[q] = 0.0;
      for (unsigned s = cap + lane; s < n_o; s += 32) bog_source<FORM, N0, C0, N1, DJ>(k, s_tab, __ldg(src + sb_o + s), bt_o, part);
#pragma unroll
      for (int q = 0; q < NRED; q++) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) part[q] += __shfl_xor_sync(0xffffffffu, part[q], off);
        if (lane == owner) acc[q] += part[q];
      }
    }
    if (!live) continue;
    // finalise: the <= 3x3 entries K[ci][cj] of the block
    double K[9];
#pragma unroll
    for (int q = 0; q < 9; q++) K[q] = 0.0;
    if (FORM == FORM_STAGED) {
#pragma unroll
      for (int q = 0; q < 9; q++) K[q] = acc[q];
    } else if (bt == 0) {
      if (FORM == GB200_FORM_ELASTICITY) {
        const double mtr = k.p1 * (acc[0] + acc[4] + acc[8]);
#pragma unroll
        for (int ci = 0; ci < 3; ci++)
#pragma unroll
          for (int cj = 0; cj < 3; cj++) K[ci * 3 + cj] = k.p0 * acc[ci * 3 + cj] + k.p1 * acc[cj * 3 + ci] + (ci == cj ? mtr : 0.0);
      } else {
        const double v = (FORM == GB200_FORM_STOKES ? 1.0 : k.p0) * acc[0];
        K[0] = K[4] = K[8] = v;   // equal components only; the other entries of the block are structural zeros (stored, = 0)
      }
    } else if (bt == 1) {
#pragma unroll
      for (int ci = 0; ci < 3; ci++) K[ci * 3 + 0] = -acc[ci];   // (v_a,ci | p_q)
    } else {
#pragma unroll
      for (int cj = 0; cj < 3; cj++) K[0 * 3 + cj] = acc[cj];    // (q | u_b,cj)
    }
    // the block's entries: row components ci_first .. ci_first + nci - 1 at the ranks r0, r0 + 1, ... of every stored column
    const int ci_first = info & 3, nci = (info >> 2) & 3;
#pragma unroll
    for (int cj = 0; cj < 3; cj++) {
      if (cj >= (N1 > 0 && fc == 1 ? 1 : C0)) continue;
      if (__ldg(&nd->clen[cj]) == 0) continue;
      double *out = k.nzval + __ldg(&nd->cbase[cj]) + r0;
#pragma unroll
      for (int ci = 0; ci < 3; ci++) {
        const int o = ci - ci_first;
        if (o < 0 || o >= nci) continue;
        if (k.add) out[o] += K[ci * 3 + cj];
        else out[o] = K[ci * 3 + cj];
      }
    }
    if (tmirror != BOG_NONE) {
      // the mirror image (column node | row node): +K^T, -K^T for the Stokes coupling blocks
      const int4 mr = __ldg(reinterpret_cast<const int4 *>(blocks + tmirror));
      const CngNode *md = k.nodes + (uint32_t)mr.x;
      const unsigned mr0 = (unsigned)mr.y & 0xFFFFu, minfo = ((unsigned)mr.y >> 16) & 0xFFu;
      const int mfc = N1 > 0 ? __ldg(&md->field) : 0;
      const int mci_first = minfo & 3, mnci = (minfo >> 2) & 3;
      const double sgn = bt == 0 ? 1.0 : -1.0;
#pragma unroll
      for (int cj = 0; cj < 3; cj++) {
        if (cj >= (N1 > 0 && mfc == 1 ? 1 : C0)) continue;
        if (__ldg(&md->clen[cj]) == 0) continue;
        double *out = k.nzval + __ldg(&md->cbase[cj]) + mr0;
#pragma unroll
        for (int ci = 0; ci < 3; ci++) {
          const int o = ci - mci_first;
          if (o < 0 || o >= mnci) continue;
          const double v = sgn * K[cj * 3 + ci];
          if (k.add) out[o] += v;
          else out[o] = v;
        }
      }
    }
  }
}

typedef void (*bog_plan_kernel_t)(CngArgs, int64_t, int64_t *, int64_t *, const int64_t *, const int64_t *, int64_t *, BogBlock *, int *);
typedef void (*bog_kernel_t)(CngArgs, const BogBlock *, const int64_t *, int64_t, int, int, int, const uint32_t *);

template <int N0, int C0, int N1>
void bog_kernels_for(int form, bool dj, bog_plan_kernel_t &pk, bog_kernel_t &gk) {
  pk = bog_plan_kernel<N0, C0, N1>;
  gk = nullptr;
  if constexpr (N1 > 0) {
    if (form == GB200_FORM_STOKES) gk = bog_gather_kernel<GB200_FORM_STOKES, N0, C0, N1, false>;
  } else {
    if (form == GB200_FORM_MASS) gk = bog_gather_kernel<GB200_FORM_MASS, N0, C0, N1, false>;
    if (form == GB200_FORM_LAPLACIAN) gk = bog_gather_kernel<GB200_FORM_LAPLACIAN, N0, C0, N1, false>;
    if constexpr (C0 == 3) {
      if (form == GB200_FORM_ELASTICITY) gk = dj ? bog_gather_kernel<GB200_FORM_ELASTICITY, N0, C0, N1, true> : bog_gather_kernel<GB200_FORM_ELASTICITY, N0, C0, N1, false>;
      if (form == FORM_STAGED) gk = bog_gather_kernel<FORM_STAGED, N0, C0, N1, false>;
    }
  }
}

bool bog_select(int n0, int c0, int n1, int form, bool dj, bog_plan_kernel_t &pk, bog_kernel_t &gk) {
  pk = nullptr;
  gk = nullptr;
  if (n1 == 0 && c0 == 3) {
    if (n0 == 27) bog_kernels_for<27, 3, 0>(form, dj, pk, gk);
    if (n0 == 8) bog_kernels_for<8, 3, 0>(form, dj, pk, gk);
    if (n0 == 10) bog_kernels_for<10, 3, 0>(form, dj, pk, gk);
    if (n0 == 4) bog_kernels_for<4, 3, 0>(form, dj, pk, gk);
  }
  if (n1 == 0 && c0 == 1) {
    if (n0 == 27) bog_kernels_for<27, 1, 0>(form, dj, pk, gk);
    if (n0 == 10) bog_kernels_for<10, 1, 0>(form, dj, pk, gk);
    if (n0 == 4) bog_kernels_for<4, 1, 0>(form, dj, pk, gk);
    if (n0 == 8) bog_kernels_for<8, 1, 0>(form, dj, pk, gk);
  }
  if (n1 == 4 && n0 == 10 && c0 == 3) bog_kernels_for<10, 3, 4>(form, dj, pk, gk);
  return pk && gk;
}

typedef void (*cng_kernel_t)(CngArgs);

template <int N0, int C0, int N1>
cng_kernel_t cng_kernel_for(int form) {
  if constexpr (N1 > 0) {
    return form == GB200_FORM_STOKES ? cng_gather_kernel<GB200_FORM_STOKES, N0, C0, N1> : nullptr;
  } else {
    if (form == GB200_FORM_MASS) return cng_gather_kernel<GB200_FORM_MASS, N0, C0, N1>;
    if (form == GB200_FORM_LAPLACIAN) return cng_gather_kernel<GB200_FORM_LAPLACIAN, N0, C0, N1>;
    if constexpr (C0 == 3) {
      if (form == GB200_FORM_ELASTICITY) return cng_gather_kernel<GB200_FORM_ELASTICITY, N0, C0, N1>;
    }
    return nullptr;
  }
}

// the instance for (element of field 0, components, element of field 1), nullptr: no instance
cng_kernel_t cng_select(int n0, int c0, int n1, int form) {
  if (n1 == 0 && c0 == 3) {
    if (n0 == 27) return cng_kernel_for<27, 3, 0>(form);   // Q2 hexahedra
    if (n0 == 8) return cng_kernel_for<8, 3, 0>(form);     // Q1 hexahedra
    if (n0 == 10) return cng_kernel_for<10, 3, 0>(form);   // P2 tetrahedra
    if (n0 == 4) return cng_kernel_for<4, 3, 0>(form);     // P1 tetrahedra
  }
  if (n1 == 0 && c0 == 1) {
    if (n0 == 27) return cng_kernel_for<27, 1, 0>(form);
    if (n0 == 10) return cng_kernel_for<10, 1, 0>(form);
    if (n0 == 4) return cng_kernel_for<4, 1, 0>(form);
    if (n0 == 8) return cng_kernel_for<8, 1, 0>(form);     // (scalar Q1 hexahedra: the headline path of q1hex_gather.cu comes first)
  }
  if (n1 == 4 && n0 == 10 && c0 == 3) return cng_kernel_for<10, 3, 4>(form);   // Taylor-Hood P2/P1
  return nullptr;
}

}  // namespace

// reference tensors of the plan's tabulation: M^{mn}_ab, mass_ab on field 0 and C^m_aq between fields 0 and 1
static void cng_tables(gb200_plan plan, std::vector<double> &tab, int &o_M, int &o_mass, int &o_C, int &o_Ms) {
  const gb200_refel_s *r0 = plan->test[0]->refel;
  const int np = r0->np, n0 = r0->nd;
  const int n1 = plan->nfields > 1 ? plan->test[1]->refel->nd : 0;
  // layout: M (elasticity) | Ms, C (Laplacian-type, Stokes) | mass -- a form stages one contiguous part in shared memory
  o_M = 0;
  o_Ms = 9 * n0 * n0;
  o_C = o_Ms + 6 * n0 * n0;
  o_mass = o_C + 3 * n0 * n1;
  tab.assign((size_t)o_mass + n0 * n0, 0.0);
  const std::vector<double> &w = plan->geo->w;
  for (int b = 0; b < n0; b++)
    for (int a = 0; a < n0; a++) {
      double mass = 0.0;
      for (int p = 0; p < np; p++) {
        mass += w[p] * r0->N[p * n0 + a] * r0->N[p * n0 + b];
        for (int m = 0; m < 3; m++)
          for (int n = 0; n < 3; n++) tab[o_M + (size_t)(b * 9 + m * 3 + n) * n0 + a] += w[p] * r0->dN[(p * n0 + a) * 3 + m] * r0->dN[(p * n0 + b) * 3 + n];
      }
      tab[o_mass + b * n0 + a] = mass;
    }
  for (int b = 0; b < n0; b++)
    for (int a = 0; a < n0; a++) {
      auto M = [&](int m, int n) { return tab[o_M + (size_t)(b * 9 + m * 3 + n) * n0 + a]; };
      const double ms[6] = {M(0, 0), M(1, 1), M(2, 2), M(0, 1) + M(1, 0), M(0, 2) + M(2, 0), M(1, 2) + M(2, 1)};
      for (int q = 0; q < 6; q++) tab[o_Ms + (size_t)(b * 6 + q) * n0 + a] = ms[q];
    }
  if (n1) {
    const gb200_refel_s *r1 = plan->test[1]->refel;
    for (int q = 0; q < n1; q++)
      for (int m = 0; m < 3; m++)
        for (int a = 0; a < n0; a++) {
          double s = 0.0;
          for (int p = 0; p < np; p++) s += w[p] * r0->dN[(p * n0 + a) * 3 + m] * r1->N[p * n1 + q];
          tab[o_C + (size_t)(q * 3 + m) * n0 + a] = s;
        }
  }
}

static void cng_fill_args(gb200_plan plan, CngArgs &k) {
  memset(&k, 0, sizeof(k));
  const ElemDesc &ed = plan->ed;
  k.nfields = plan->nfields;
  k.NL = plan->NL;
  int nofs = 0;
  for (int f = 0; f < plan->nfields; f++) {
    k.nds[f] = ed.f[f].nds; k.ncomp[f] = ed.f[f].ncomp; k.lofs[f] = ed.f[f].lofs; k.nofs[f] = nofs;
    nofs += ed.f[f].nds;
    k.row_ids[f] = ed.f[f].row_ids; k.col_ids[f] = ed.f[f].col_ids; k.col_off[f] = ed.f[f].col_off;
  }
  k.nn_tot = nofs;
  k.ncells = plan->mesh->ncells;
  k.ncols = plan->ncols;
  k.colptr = plan->colptr.p;
  k.rank = plan->rank.p;
}

// 1: the plan can be assembled by the column-node gather (checked once per plan), 0: not
bool affine_gather_supported(gb200_plan plan, int form) {
  if (getenv("GB200_NO_AFFINE_GATHER") != nullptr) return false;   // (tests: keep the cell-centric kernels reachable on affine meshes)
  const ElemDesc &ed = plan->ed;
  if (ed.D != 3 || ed.Dr != 3 || plan->mesh->ncells == 0 || ed.lface) return false;
  if (form == GB200_FORM_STOKES) {
    if (plan->nfields != 2 || ed.f[0].ncomp != 3 || ed.f[1].ncomp != 1) return false;
  } else if (form == GB200_FORM_MASS || form == GB200_FORM_LAPLACIAN) {
    if (plan->nfields != 1 || (ed.f[0].ncomp != 1 && ed.f[0].ncomp != 3)) return false;
  } else if (form == GB200_FORM_ELASTICITY) {
    if (plan->nfields != 1 || ed.f[0].ncomp != 3) return false;
  } else {
    return false;
  }
  if (!cng_select(ed.f[0].nds, ed.f[0].ncomp, plan->nfields > 1 ? ed.f[1].nds : 0, form)) return false;   // no instance for this element
  if (plan->cng_ok < 0) plan->cng_ok = mesh_check_affine(plan->mesh) ? 1 : 0;
  return plan->cng_ok == 1;
}

static void cng_build_plan(gb200_plan plan) {
  gb200_ctx ctx = plan->ctx;
  cudaStream_t s = ctx->stream;
  ScopedTimer timer(ctx, "affine_gather_plan");
  CngArgs k;
  cng_fill_args(plan, k);
  const int64_t total = k.ncells * k.nn_tot, ncols = plan->ncols;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)ctx->num_sms * 32));
  const int gcol = (int)std::max<int64_t>(1, std::min<int64_t>((ncols + 256) / 256, (int64_t)ctx->num_sms * 32));
  DevBuf<int64_t> cnt, cursor, ptr, flag, node_index;
  cnt.alloc(ncols + 1);
  cnt.zero(s);
  cng_count_kernel<<<grid, 256, 0, s>>>(k, (unsigned long long *)cnt.p);
  check_launch(ctx, "cng_count_kernel");
  ptr.alloc(ncols + 1);
  flag.alloc(ncols + 1);
  node_index.alloc(ncols + 1);
  cng_flag_kernel<<<gcol, 256, 0, s>>>(cnt.p, ncols, flag.p);
  check_launch(ctx, "cng_flag_kernel");
  {
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt.p, ptr.p, ncols + 1, s);
    DevBuf<char> tmp;
    tmp.alloc(tmp_bytes);
    cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, cnt.p, ptr.p, ncols + 1, s);            // cnt[ncols] = 0: ptr[ncols] = number of entries
    cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, flag.p, node_index.p, ncols + 1, s);    // node_index[ncols] = number of nodes
    count_launch(ctx, 2);
  }
  int64_t nent = 0, nnodes = 0;
  GB_CUDA(cudaMemcpyAsync(&nent, ptr.p + ncols, 8, cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaMemcpyAsync(&nnodes, node_index.p + ncols, 8, cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaStreamSynchronize(s));
  flag.release();
  plan->cng_adj.alloc((size_t)std::max<int64_t>(nent, 1));
  cursor.alloc(ncols + 1);
  cursor.zero(s);
  cng_fill_kernel<<<grid, 256, 0, s>>>(k, ptr.p, (unsigned long long *)cursor.p, plan->cng_adj.p);
  check_launch(ctx, "cng_fill_kernel");
  const int64_t nunits = (nent + CNG_CHUNK - 1) / CNG_CHUNK;
  plan->cng_nodes.alloc((size_t)std::max<int64_t>(nnodes, 1) * sizeof(CngNode));
  plan->cng_unit_ptr.alloc((size_t)nunits + 1);
  DevBuf<int64_t> bufmax;
  bufmax.alloc(1);
  bufmax.zero(s);
  cng_sort_kernel<<<gcol, 256, 0, s>>>(k, ptr.p, node_index.p, plan->cng_adj.p, reinterpret_cast<CngNode *>(plan->cng_nodes.p), plan->cng_unit_ptr.p,
                                       (unsigned long long *)bufmax.p);
  check_launch(ctx, "cng_sort_kernel");
  GB_CUDA(cudaMemcpyAsync(plan->cng_unit_ptr.p + nunits, &nnodes, 8, cudaMemcpyHostToDevice, s));
  int64_t h = 0;
  bufmax.download(&h, s);
  std::vector<double> tab;
  cng_tables(plan, tab, plan->cng_oM, plan->cng_omass, plan->cng_oC, plan->cng_oMs);
  plan->cng_tab.upload(tab.data(), tab.size(), s);
  GB_CUDA(cudaStreamSynchronize(s));
  plan->cng_buf_len = (int)h;
  plan->cng_nunits = nunits;
  plan->cng_nent = nent;
  plan->cng_built = true;
}

// block plan (source lists per stored node-pair block); plan->bog_state: 1 built, -1 not available (irregular blocks, too many
// incident cells, not enough device memory): the column-node kernel is used instead
static void bog_build_plan(gb200_plan plan, int form) {
  gb200_ctx ctx = plan->ctx;
  cudaStream_t s = ctx->stream;
  plan->bog_state = -1;
  if (getenv("GB200_NO_BLOCK_GATHER") != nullptr) return;
  const ElemDesc &ed = plan->ed;
  bog_plan_kernel_t pk;
  bog_kernel_t gk;
  if (!bog_select(ed.f[0].nds, ed.f[0].ncomp, plan->nfields > 1 ? ed.f[1].nds : 0, form, false, pk, gk)) return;
  ScopedTimer timer(ctx, "affine_block_plan");
  CngArgs k;
  cng_fill_args(plan, k);
  k.nodes = reinterpret_cast<const CngNode *>(plan->cng_nodes.p);
  k.adj = plan->cng_adj.p;
  const int64_t nnodes = (int64_t)(plan->cng_nodes.n / sizeof(CngNode));
  try {
    DevBuf<int64_t> nsrc, nblk, src_ptr, blk_ptr;
    DevBuf<int> err;
    nsrc.alloc(nnodes + 1);
    nblk.alloc(nnodes + 1);
    nsrc.zero(s);
    nblk.zero(s);
    err.alloc(1);
    err.zero(s);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nnodes + 3) / 4, (int64_t)ctx->num_sms * 16));
    pk<<<grid, 128, 0, s>>>(k, nnodes, nsrc.p, nblk.p, nullptr, nullptr, nullptr, nullptr, err.p);
    check_launch(ctx, "bog_plan_kernel");
    src_ptr.alloc(nnodes + 1);
    blk_ptr.alloc(nnodes + 1);
    {
      size_t tmp_bytes = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, nsrc.p, src_ptr.p, nnodes + 1, s);
      DevBuf<char> tmp;
      tmp.alloc(tmp_bytes);
      cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, nsrc.p, src_ptr.p, nnodes + 1, s);
      cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, nblk.p, blk_ptr.p, nnodes + 1, s);
      count_launch(ctx, 2);
    }
    int64_t tot_src = 0, tot_blk = 0;
    int herr = 0;
    GB_CUDA(cudaMemcpyAsync(&tot_src, src_ptr.p + nnodes, 8, cudaMemcpyDeviceToHost, s));
    GB_CUDA(cudaMemcpyAsync(&tot_blk, blk_ptr.p + nnodes, 8, cudaMemcpyDeviceToHost, s));
    err.download(&herr, s);
    GB_CUDA(cudaStreamSynchronize(s));
    if (herr) return;
    size_t free_b = 0, total_b = 0;
    GB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    if ((size_t)tot_src * 8 + (size_t)tot_blk * 16 + ((size_t)2 << 30) > free_b) return;   // keep 2 GB of head room: column-node kernel instead
    plan->bog_src.alloc((size_t)std::max<int64_t>(tot_src, 1));
    plan->bog_blocks.alloc((size_t)std::max<int64_t>(tot_blk, 1) * sizeof(BogBlock));
    pk<<<grid, 128, 0, s>>>(k, nnodes, nullptr, nullptr, src_ptr.p, blk_ptr.p, plan->bog_src.p, reinterpret_cast<BogBlock *>(plan->bog_blocks.p), err.p);
    check_launch(ctx, "bog_plan_kernel");
    err.download(&herr, s);
    GB_CUDA(cudaStreamSynchronize(s));
    if (herr) {
      plan->bog_src.release();
      plan->bog_blocks.release();
      return;
    }
    plan->bog_nblocks = tot_blk;
    plan->bog_state = 1;
    // mirror pairs: rows and columns must share their numbering (same DoF tables and offsets), at most 2^32 - 2 blocks
    bool same = tot_blk < (int64_t)0xFFFFFFF0ll && getenv("GB200_MIRROR") != nullptr;   // opt-in: measured slower (the mirror writes are scattered 24-byte pieces)
    for (int f = 0; f < plan->nfields; f++)
      same = same && plan->test[f]->cell_dofs.p == plan->trial[f]->cell_dofs.p && plan->row_off[f] == plan->col_off[f];
    plan->bog_npairs = 0;
    if (same && plan->nrows == plan->ncols) {
      ScopedTimer tm(ctx, "affine_mirror_plan");
      DevBuf<int32_t> col_node;
      DevBuf<uint32_t> mirror;
      DevBuf<int64_t> flag, pos;
      col_node.alloc((size_t)plan->ncols);
      GB_CUDA(cudaMemsetAsync(col_node.p, 0xFF, (size_t)plan->ncols * 4, s));
      mirror.alloc((size_t)tot_blk);
      flag.alloc((size_t)tot_blk + 1);
      pos.alloc((size_t)tot_blk + 1);
      const int gn = (int)std::max<int64_t>(1, std::min<int64_t>((nnodes + 255) / 256, (int64_t)ctx->num_sms * 32));
      const int gb_ = (int)std::max<int64_t>(1, std::min<int64_t>((tot_blk + 256) / 256, (int64_t)ctx->num_sms * 32));
      bog_col_node_kernel<<<gn, 256, 0, s>>>(k.nodes, nnodes, plan->colptr.p, plan->ncols, col_node.p);
      check_launch(ctx, "bog_col_node_kernel");
      bog_mirror_kernel<<<gb_, 256, 0, s>>>(reinterpret_cast<const BogBlock *>(plan->bog_blocks.p), tot_blk, k.nodes, blk_ptr.p, plan->colptr.p,
                                           plan->rowval.p, plan->ncols, col_node.p, mirror.p);
      check_launch(ctx, "bog_mirror_kernel");
      bog_owner_flag_kernel<<<gb_, 256, 0, s>>>(mirror.p, tot_blk, flag.p);
      check_launch(ctx, "bog_owner_flag_kernel");
      {
        size_t tmp_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, flag.p, pos.p, tot_blk + 1, s);
        DevBuf<char> tmp;
        tmp.alloc(tmp_bytes);
        cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, flag.p, pos.p, tot_blk + 1, s);
        count_launch(ctx, 1);
      }
      int64_t npairs = 0;
      GB_CUDA(cudaMemcpyAsync(&npairs, pos.p + tot_blk, 8, cudaMemcpyDeviceToHost, s));
      GB_CUDA(cudaStreamSynchronize(s));
      if (npairs < tot_blk) {   // something pairs up: worth the indirection
        plan->bog_pairs.alloc((size_t)std::max<int64_t>(npairs, 1) * 2);
        bog_pairs_kernel<<<gb_, 256, 0, s>>>(mirror.p, tot_blk, pos.p, plan->bog_pairs.p);
        check_launch(ctx, "bog_pairs_kernel");
        GB_CUDA(cudaStreamSynchronize(s));
        plan->bog_npairs = npairs;
      }
    }
  } catch (const gb::Error &) {
    cudaGetLastError();
    plan->bog_src.release();
    plan->bog_blocks.release();
  }
}

bool launch_affine_gather(gb200_plan plan, int form, const double *params, double *nzval, bool add) {
  if (!affine_gather_supported(plan, form)) return false;
  gb200_ctx ctx = plan->ctx;
  cudaStream_t s = ctx->stream;
  if (!plan->cng_built) cng_build_plan(plan);
  if (plan->bog_state == 0) bog_build_plan(plan, form);
  if (plan->bog_state == 1) {
    const ElemDesc &ed = plan->ed;
    bog_plan_kernel_t pk;
    bog_kernel_t gk;
    const int64_t nc = plan->mesh->ncells;
    if (plan->cellF.n != (size_t)(CNG_F * nc)) plan->cellF.alloc((size_t)(CNG_F * nc));
    if (plan->cng_diag < 0) {   // once per plan: is inv(Jt) diagonal in every cell (exact zeros)?
      DevBuf<int> nd;
      nd.alloc(1);
      nd.zero(s);
      cng_factors_kernel<<<(int)((nc + 127) / 128), 128, 0, s>>>(plan->mesh->X.p, plan->mesh->cell_nodes.p, plan->mesh->nn, plan->ed.dNg, nc, plan->cellF.p, nd.p);
      check_launch(ctx, "cng_factors_kernel");
      int h = 0;
      nd.download(&h, s);
      GB_CUDA(cudaStreamSynchronize(s));
      plan->cng_diag = h ? 0 : 1;
    }
    if (bog_select(ed.f[0].nds, ed.f[0].ncomp, plan->nfields > 1 ? ed.f[1].nds : 0, form, plan->cng_diag == 1 && !getenv("GB200_NO_DIAG_J"), pk, gk)) {
      {
        ScopedTimer t(ctx, "k:affine_factors");
        cng_factors_kernel<<<(int)((nc + 127) / 128), 128, 0, s>>>(plan->mesh->X.p, plan->mesh->cell_nodes.p, plan->mesh->nn, plan->ed.dNg, nc, plan->cellF.p, nullptr);
        check_launch(ctx, "cng_factors_kernel");
      }
      ScopedTimer t(ctx, "k:affine_gather");
      CngArgs k;
      cng_fill_args(plan, k);
      k.nodes = reinterpret_cast<const CngNode *>(plan->cng_nodes.p);
      k.F = plan->cellF.p;
      k.tab = plan->cng_tab.p;
      k.o_M = plan->cng_oM; k.o_mass = plan->cng_omass; k.o_C = plan->cng_oC; k.o_Ms = plan->cng_oMs;
      k.p0 = params[0]; k.p1 = params[1];
      k.nzval = nzval;
      k.add = add ? 1 : 0;
      // the contiguous part of the table the form reads: M | Ms, C | mass
      const int n0 = ed.f[0].nds;
      int tab_off = plan->cng_oMs, tab_len = plan->cng_omass - plan->cng_oMs;
      if (form == GB200_FORM_ELASTICITY) { tab_off = plan->cng_oM; tab_len = 9 * n0 * n0; }
      if (form == GB200_FORM_MASS) { tab_off = plan->cng_omass; tab_len = n0 * n0; }
      const size_t smem = (size_t)tab_len * sizeof(double);
      static std::map<std::pair<const void *, int>, bool> opted;
      auto key = std::make_pair(reinterpret_cast<const void *>(gk), ctx->device);
      if (!opted.count(key)) {
        GB_CUDA(cudaFuncSetAttribute(gk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        opted[key] = true;
      }
      int cps = 0;
      GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cps, gk, BOG_THREADS, smem));
      cps = std::max(cps, 1);
      const bool mirrored = plan->bog_npairs > 0;
      const int64_t nwork = mirrored ? plan->bog_npairs : plan->bog_nblocks;
      const int64_t want = (nwork + BOG_THREADS - 1) / BOG_THREADS;
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)ctx->num_sms * cps * 4));
      static const int cap_env = getenv("GB200_BLOCK_CAP") ? atoi(getenv("GB200_BLOCK_CAP")) : 0;
      const int cap = cap_env > 0 ? cap_env : (plan->mesh->nn == 4 ? 6 : 4);   // sources a thread evaluates itself (tets: 2 - 6 per off-diagonal block)
      gk<<<grid, BOG_THREADS, smem, s>>>(k, reinterpret_cast<const BogBlock *>(plan->bog_blocks.p), plan->bog_src.p, nwork, tab_off, tab_len, cap,
                                         mirrored ? plan->bog_pairs.p : nullptr);
      check_launch(ctx, "bog_gather_kernel");
      plan->path_detail[form] = std::string(mirrored ? "pairs" : "blocks") + ((form == GB200_FORM_ELASTICITY && plan->cng_diag == 1 && !getenv("GB200_NO_DIAG_J")) ? "+diagJ" : "");
      return true;
    }
  }
  plan->path_detail[form] = "columns";
  const size_t smem = (size_t)(CNG_THREADS / 32) * plan->cng_buf_len * sizeof(double);
  if (smem > 200 * 1024) return false;   // (columns too long for the shared-memory buffers: cell-centric kernels)
  const int64_t nc = plan->mesh->ncells;
  if (plan->cellF.n != (size_t)(CNG_F * nc)) plan->cellF.alloc((size_t)(CNG_F * nc));
  {
    ScopedTimer t(ctx, "k:affine_factors");
    cng_factors_kernel<<<(int)((nc + 127) / 128), 128, 0, s>>>(plan->mesh->X.p, plan->mesh->cell_nodes.p, plan->mesh->nn, plan->ed.dNg, nc, plan->cellF.p, nullptr);
    check_launch(ctx, "cng_factors_kernel");
  }
  ScopedTimer t(ctx, "k:affine_gather");
  CngArgs k;
  cng_fill_args(plan, k);
  k.nodes = reinterpret_cast<const CngNode *>(plan->cng_nodes.p);
  k.unit_ptr = plan->cng_unit_ptr.p;
  k.nunits = plan->cng_nunits;
  k.nent = plan->cng_nent;
  k.adj = plan->cng_adj.p;
  k.F = plan->cellF.p;
  k.tab = plan->cng_tab.p;
  k.o_M = plan->cng_oM; k.o_mass = plan->cng_omass; k.o_C = plan->cng_oC; k.o_Ms = plan->cng_oMs;
  k.p0 = params[0]; k.p1 = params[1];
  k.nzval = nzval;
  k.add = add ? 1 : 0;
  k.buf_len = plan->cng_buf_len;
  const ElemDesc &ed = plan->ed;
  cng_kernel_t kern = cng_select(ed.f[0].nds, ed.f[0].ncomp, plan->nfields > 1 ? ed.f[1].nds : 0, form);
  static std::map<std::pair<const void *, int>, bool> opted;   // the opt-in belongs to (function, device)
  auto key = std::make_pair(reinterpret_cast<const void *>(kern), ctx->device);
  if (!opted.count(key)) {
    GB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    opted[key] = true;
  }
  int cps = 0;
  GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cps, kern, CNG_THREADS, smem));
  cps = std::max(cps, 1);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((plan->cng_nunits + 3) / 4, (int64_t)ctx->num_sms * cps));
  kern<<<grid, CNG_THREADS, smem, s>>>(k);
  check_launch(ctx, "cng_gather_kernel");
  return true;
}

// Staged gather: the cell-centric kernel of vector_kernels.cu writes the 3x3 blocks of the node pairs a <= b of every cell to HBM
// (any geometry, state-dependent integrands: neo-Hookean), the block-owner gather then sums, per stored node-pair block, its source
// blocks in ascending cell order and writes every nnz slot exactly once: no atomics on the matrix, no zero-fill, deterministic.
// The local vector (source / neo-Hookean residual) stays fused in the cell kernel.
bool launch_staged_gather(gb200_plan plan, int form, int form_vec, const double *params, const double *fq, double *nzval, double *bvec, bool add, bool zero_vec) {
  if (getenv("GB200_NO_STAGED_GATHER") != nullptr) return false;
  if (form != GB200_FORM_MASS && form != GB200_FORM_LAPLACIAN && form != GB200_FORM_ELASTICITY && form != GB200_FORM_NEOHOOKEAN_JAC) return false;
  if (form_vec != 0 && !(form_vec == GB200_FORM_NEOHOOKEAN_RES && form == GB200_FORM_NEOHOOKEAN_JAC)) return false;   // fused: residual + Jacobian only
  int npair = 0;
  if (!vector_kernel_pairs(plan, npair) || plan->mesh->ncells == 0) return false;
  // Q2 elasticity keeps the FP64 tensor-core kernel (27 KB of staged blocks per cell would have to go through HBM twice)
  if (form == GB200_FORM_ELASTICITY && plan->ed.f[0].nds == 27 && getenv("GB200_STAGED_Q2") == nullptr) return false;
  gb200_ctx ctx = plan->ctx;
  cudaStream_t s = ctx->stream;
  const ElemDesc &ed = plan->ed;
  bog_plan_kernel_t pk;
  bog_kernel_t gk;
  if (!bog_select(ed.f[0].nds, 3, 0, FORM_STAGED, false, pk, gk)) return false;
  if (!plan->cng_built) cng_build_plan(plan);
  if (plan->bog_state == 0) bog_build_plan(plan, FORM_STAGED);
  if (plan->bog_state != 1) return false;
  const size_t need = (size_t)plan->mesh->ncells * npair * 9;
  if (plan->ke_stage.n != need) {
    size_t free_b = 0, total_b = 0;
    GB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    if (need * 8 + ((size_t)2 << 30) > free_b + plan->ke_stage.n * 8) return false;   // not enough device memory: cell-centric scatter
    plan->ke_stage.alloc(need);
  }
  const bool fused_vec = form_vec != 0 && bvec != nullptr;
  if (fused_vec && zero_vec) {
    GB_CUDA(cudaMemsetAsync(bvec, 0, (size_t)plan->nrows * 8, s));
    count_launch(ctx, 1);
  }
  if (!launch_vector_kernel(plan, form, fused_vec ? form_vec : 0, params, fq, nullptr, fused_vec ? bvec : nullptr, plan->ke_stage.p)) return false;
  ScopedTimer t(ctx, "k:staged_gather");
  CngArgs k;
  cng_fill_args(plan, k);
  k.nodes = reinterpret_cast<const CngNode *>(plan->cng_nodes.p);
  k.Ke = plan->ke_stage.p;
  k.tab = plan->cng_tab.p;
  k.nzval = nzval;
  k.add = add ? 1 : 0;
  static std::map<std::pair<const void *, int>, bool> opted;
  auto key = std::make_pair(reinterpret_cast<const void *>(gk), ctx->device);
  if (!opted.count(key)) {
    GB_CUDA(cudaFuncSetAttribute(gk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    opted[key] = true;
  }
  int cps = 0;
  GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cps, gk, BOG_THREADS, 0));
  cps = std::max(cps, 1);
  const bool mirrored = plan->bog_npairs > 0;
  const int64_t nwork = mirrored ? plan->bog_npairs : plan->bog_nblocks;
  const int64_t want = (nwork + BOG_THREADS - 1) / BOG_THREADS;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)ctx->num_sms * cps * 4));
  gk<<<grid, BOG_THREADS, 0, s>>>(k, reinterpret_cast<const BogBlock *>(plan->bog_blocks.p), plan->bog_src.p, nwork, 0, 0, 8,
                                  mirrored ? plan->bog_pairs.p : nullptr);
  check_launch(ctx, "bog_gather_kernel");
  plan->path_detail[form] = mirrored ? "pairs" : "blocks";
  return true;
}

}  // namespace gb
