// q1hex_gather.cu -- owner-computes assembly for scalar Q1 hexahedra on affine cells (the headline path).
//
// Reference work being replaced (per cell: a4-a8, a13): Jt at the quadrature points, inv/det, physical
// gradients, aq[p,i,j], IntegrationMap, then 64 binary-search insertions into the CSC
// (src/Fields/FieldsInterfaces.jl:737-760, src/Algebra/SparseMatrixCSC.jl:124-150).
//
// B200 design: no atomics, no zero-fill, every nnz slot written exactly once, fully coalesced:
//   kernel 1 (cell-parallel):   G_c = |det Jt| inv(Jt)^T inv(Jt) (6 doubles) and |det Jt| from the node
//                               coordinates; for an affine cell Jt is constant, so hoisting it out of the
//                               quadrature loop is exact up to round-off.
//   kernel 2 (column-parallel): a thread owns one CSC column j, a warp 32 consecutive columns (= one contiguous
//                               nzval range).  For each incident (cell, lj) the thread evaluates the 8 entries
//                               K_e[:,lj] = sum_kl G_kl M^{kl}[:,lj] in closed form (M^{kl}_{ab} = sum_q w_q d_kN_a d_lN_b,
//                               exact for the 2x2x2 Gauss rule) and adds them at their in-column ranks:
//                                 * canonical blocks (3x3x3 stencil, detected in the plan): 27 register accumulators
//                                   with compile-time indices, no rank loads;
//                                 * any other block: accumulators in shared memory, ranks from the plan.
//                               The warp's nzval range is staged in shared memory (laid out exactly like nzval) and
//                               streamed out with coalesced stores.  Per-slot summation order = ascending cell order,
//                               the reference's own order (SparseMatrixAssemblers.jl:242-247) => deterministic, and both
//                               branches give bitwise identical results.
#include "common.cuh"

namespace gb {

namespace {

constexpr int GATHER_THREADS = 128;

__global__ void __launch_bounds__(256) cell_geom_kernel(const double *__restrict__ X, const int32_t *__restrict__ cell_nodes,
                                                        int64_t ncells, double *__restrict__ G, int want_det) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  const int4 *cn = reinterpret_cast<const int4 *>(cell_nodes + c * 8);
  int4 n0 = cn[0], n1 = cn[1];
  int ids[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
  double x[8][3];
#pragma unroll
  for (int a = 0; a < 8; a++) {
    const double *p = X + (int64_t)ids[a] * 3;
    x[a][0] = p[0]; x[a][1] = p[1]; x[a][2] = p[2];
  }
  // Jt[i][:] = dx/dxi_i at the cell centre = mean of the four edges parallel to axis i
  double J[9];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    J[0 + d] = 0.25 * ((x[1][d] - x[0][d]) + (x[3][d] - x[2][d]) + (x[5][d] - x[4][d]) + (x[7][d] - x[6][d]));
    J[3 + d] = 0.25 * ((x[2][d] - x[0][d]) + (x[3][d] - x[1][d]) + (x[6][d] - x[4][d]) + (x[7][d] - x[5][d]));
    J[6 + d] = 0.25 * ((x[4][d] - x[0][d]) + (x[5][d] - x[1][d]) + (x[6][d] - x[2][d]) + (x[7][d] - x[3][d]));
  }
  double det = J[0] * J[4] * J[8] + J[1] * J[5] * J[6] + J[2] * J[3] * J[7] - (J[0] * J[5] * J[7] + J[1] * J[3] * J[8] + J[2] * J[4] * J[6]);
  double ci = 1.0 / det;
  double I[9];
  I[0] = (J[4] * J[8] - J[5] * J[7]) * ci;
  I[1] = -(J[1] * J[8] - J[2] * J[7]) * ci;
  I[2] = (J[1] * J[5] - J[2] * J[4]) * ci;
  I[3] = -(J[3] * J[8] - J[5] * J[6]) * ci;
  I[4] = (J[0] * J[8] - J[2] * J[6]) * ci;
  I[5] = -(J[0] * J[5] - J[2] * J[3]) * ci;
  I[6] = (J[3] * J[7] - J[4] * J[6]) * ci;
  I[7] = -(J[0] * J[7] - J[1] * J[6]) * ci;
  I[8] = (J[0] * J[4] - J[1] * J[3]) * ci;
  double ad = fabs(det);
  // grad(phi) = I . grad(N)  =>  grad(phi_a).grad(phi_b) = gN_a^T (I^T I) gN_b ;  Gm[k][l] = |det| sum_i I[i][k] I[i][l]
  // SoA layout [7][ncells]: lanes of a warp own consecutive cells, so every load/store is one 256-byte request
  G[c] = ad * (I[0] * I[0] + I[3] * I[3] + I[6] * I[6]);
  G[ncells + c] = ad * (I[1] * I[1] + I[4] * I[4] + I[7] * I[7]);
  G[2 * ncells + c] = ad * (I[2] * I[2] + I[5] * I[5] + I[8] * I[8]);
  G[3 * ncells + c] = ad * (I[0] * I[1] + I[3] * I[4] + I[6] * I[7]);
  G[4 * ncells + c] = ad * (I[0] * I[2] + I[3] * I[5] + I[6] * I[8]);
  G[5 * ncells + c] = ad * (I[1] * I[2] + I[4] * I[5] + I[7] * I[8]);
  if (want_det) G[6 * ncells + c] = ad;
}

// K_e[a][b] for the Laplacian on an affine Q1 hex, as a function of t_d = +1 if a_d == b_d else -1:
//   m_d = 1/4 + t_d/12 (1-D mass), 1-D stiffness = t_d, mixed terms carry tau_k tau_l (folded into o_kl by the caller)
template <int T0, int T1, int T2>
__device__ __forceinline__ double lap_entry(double d0, double d1, double d2, double o01, double o02, double o12) {
  constexpr double m0 = 0.25 + T0 / 12.0, m1 = 0.25 + T1 / 12.0, m2 = 0.25 + T2 / 12.0;
  double v = d0 * (T0 * m1 * m2);
  v = fma(d1, m0 * T1 * m2, v);
  v = fma(d2, m0 * m1 * T2, v);
  if (T0 + T1 != 0) v = fma(o01, m2 * (T0 + T1), v);
  if (T0 + T2 != 0) v = fma(o02, m1 * (T0 + T2), v);
  if (T1 + T2 != 0) v = fma(o12, m0 * (T1 + T2), v);
  return v;
}
template <int T0, int T1, int T2>
__device__ __forceinline__ double mass_entry(double ad) {
  constexpr double m0 = 0.25 + T0 / 12.0, m1 = 0.25 + T1 / 12.0, m2 = 0.25 + T2 / 12.0;
  return ad * (m0 * m1 * m2);
}

// the 8 entries K_e[li][lj], indexed by the flip mask m = li ^ lj (bit d set <=> a_d != b_d)
template <int FORM>
__device__ __forceinline__ void column_entries(const double *__restrict__ G, int64_t ncells, int64_t cell, int lj, double coef, double *vals) {
  if (FORM == GB200_FORM_LAPLACIAN) {
    const double t0 = (lj & 1) ? 1.0 : -1.0, t1 = (lj & 2) ? 1.0 : -1.0, t2 = (lj & 4) ? 1.0 : -1.0;
    const double d0 = coef * __ldg(G + cell), d1 = coef * __ldg(G + ncells + cell), d2 = coef * __ldg(G + 2 * ncells + cell);
    const double o01 = 0.25 * coef * t0 * t1 * __ldg(G + 3 * ncells + cell), o02 = 0.25 * coef * t0 * t2 * __ldg(G + 4 * ncells + cell),
                 o12 = 0.25 * coef * t1 * t2 * __ldg(G + 5 * ncells + cell);
    vals[0] = lap_entry<+1, +1, +1>(d0, d1, d2, o01, o02, o12);
    vals[1] = lap_entry<-1, +1, +1>(d0, d1, d2, o01, o02, o12);
    vals[2] = lap_entry<+1, -1, +1>(d0, d1, d2, o01, o02, o12);
    vals[3] = lap_entry<-1, -1, +1>(d0, d1, d2, o01, o02, o12);
    vals[4] = lap_entry<+1, +1, -1>(d0, d1, d2, o01, o02, o12);
    vals[5] = lap_entry<-1, +1, -1>(d0, d1, d2, o01, o02, o12);
    vals[6] = lap_entry<+1, -1, -1>(d0, d1, d2, o01, o02, o12);
    vals[7] = lap_entry<-1, -1, -1>(d0, d1, d2, o01, o02, o12);
  } else {
    const double ad = coef * G[6 * ncells + cell];
    vals[0] = mass_entry<+1, +1, +1>(ad);
    vals[1] = mass_entry<-1, +1, +1>(ad);
    vals[2] = mass_entry<+1, -1, +1>(ad);
    vals[3] = mass_entry<-1, -1, +1>(ad);
    vals[4] = mass_entry<+1, +1, -1>(ad);
    vals[5] = mass_entry<-1, +1, -1>(ad);
    vals[6] = mass_entry<+1, -1, -1>(ad);
    vals[7] = mass_entry<-1, -1, -1>(ad);
  }
}

// canonical block: rank of the row with flip mask M in the column, for the Q-th incident cell (lj = 7 - Q)
__host__ __device__ constexpr int canon_rank(int Q, int M) {
  int r = 0, pw = 1;
  for (int d = 0; d < 3; d++) {
    int c = (Q >> d) & 1;
    int a = (1 - c) ^ ((M >> d) & 1);  // a_d = b_d ^ m_d with b_d = 1 - c_d
    r += pw * (c + a);
    pw *= 3;
  }
  return r;
}

template <int FORM, int Q>
__device__ __forceinline__ void canon_cell(int32_t e, const double *__restrict__ G, int64_t ncells, double coef, double *acc) {
  double vals[8];
  column_entries<FORM>(G, ncells, (int64_t)(e >> 3), 7 - Q, coef, vals);
#pragma unroll
  for (int m = 0; m < 8; m++) acc[canon_rank(Q, m)] += vals[m];
}

template <int FORM, int MINB>
__global__ void __launch_bounds__(GATHER_THREADS, MINB) q1hex_gather_kernel(const int64_t *__restrict__ colptr, const int64_t *__restrict__ blk_ptr,
                                                                      const uint8_t *__restrict__ blk_flag, const uint32_t *__restrict__ col_mask,
                                                                      const int32_t *__restrict__ blk_base,
                                                                      const int32_t *__restrict__ adjT_cell,
                                                                      const uint64_t *__restrict__ adjT_rank, const double *__restrict__ G,
                                                                      int64_t ncells, int64_t ncols, double coef, double *__restrict__ nzval,
                                                                      int add, int use_canon, int wspan_max, int prefetch) {
  // persistent warps: warp w handles the 32-column blocks w, w + W, w + 2W, ... with its own staging buffer
  extern __shared__ double stage[];
  const int lane = threadIdx.x & 31;
  double *wstage = stage + (size_t)(threadIdx.x >> 5) * wspan_max;
  const int64_t nblocks = (ncols + 31) >> 5;
  const int64_t wstride = (int64_t)gridDim.x * (GATHER_THREADS / 32);
  for (int64_t blk = (int64_t)blockIdx.x * (GATHER_THREADS / 32) + (threadIdx.x >> 5); blk < nblocks; blk += wstride) {
  const int64_t jw0 = blk * 32;
  const int64_t jw1 = min(jw0 + 32, ncols);
  const int64_t wbase = colptr[jw0];
  const int wspan = (int)(colptr[jw1] - wbase);
  const int64_t j = jw0 + lane;
  const int64_t row0 = blk_ptr[blk];
  const int flag = use_canon ? blk_flag[blk] : 0;
  if (prefetch) {
    // software prefetch of the geometry factors of this warp's NEXT block (run-compressed blocks only: the 8 rows are
    // 256-byte runs, 96 cache lines in total, 3 per lane) so that their DRAM latency overlaps this block's work
    const int64_t nb = blk + wstride;
    if (nb < nblocks && (blk_flag[nb] & 4)) {
      const int32_t *bb = blk_base + nb * 8;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const int idx = lane + 32 * k, q = idx / 12, rem = idx - 12 * q, a = rem >> 1, h = rem & 1;
        const double *ptr = G + (int64_t)a * ncells + (__ldg(bb + q) >> 3) + 16 * h;
        if (prefetch == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
        else asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr));
      }
    }
  }
  if (flag) {
    int32_t e[8];
    if (flag & 4) {  // run-length compressed rows: consecutive cells across the lanes
      const int4 *bb = reinterpret_cast<const int4 *>(blk_base + blk * 8);
      const int4 b0 = __ldg(bb), b1 = __ldg(bb + 1);
      e[0] = b0.x + 8 * lane; e[1] = b0.y + 8 * lane; e[2] = b0.z + 8 * lane; e[3] = b0.w + 8 * lane;
      e[4] = b1.x + 8 * lane; e[5] = b1.y + 8 * lane; e[6] = b1.z + 8 * lane; e[7] = b1.w + 8 * lane;
    } else {
      const int32_t *rows = adjT_cell + row0 * 32;
#pragma unroll
      for (int q = 0; q < 8; q++) e[q] = __ldg(rows + q * 32 + lane);  // 8 independent coalesced loads
    }
    double acc[27];
#pragma unroll
    for (int r = 0; r < 27; r++) acc[r] = 0.0;
    canon_cell<FORM, 0>(e[0], G, ncells, coef, acc);
    canon_cell<FORM, 1>(e[1], G, ncells, coef, acc);
    canon_cell<FORM, 2>(e[2], G, ncells, coef, acc);
    canon_cell<FORM, 3>(e[3], G, ncells, coef, acc);
    canon_cell<FORM, 4>(e[4], G, ncells, coef, acc);
    canon_cell<FORM, 5>(e[5], G, ncells, coef, acc);
    canon_cell<FORM, 6>(e[6], G, ncells, coef, acc);
    canon_cell<FORM, 7>(e[7], G, ncells, coef, acc);
    if ((flag & 3) == 1) {
      double *my = wstage + 27 * lane;
#pragma unroll
      for (int r = 0; r < 27; r++) my[r] = acc[r];
    } else {
      // stencil subset (e.g. next to a Dirichlet boundary): static register index, compacted in-column rank
      const uint32_t mask = col_mask[j];
      double *my = wstage + (colptr[j] - wbase);
#pragma unroll
      for (int r = 0; r < 27; r++)
        if ((mask >> r) & 1u) my[__popc(mask & ((1u << r) - 1u))] = acc[r];
    }
  } else {
    for (int k = lane; k < wspan; k += 32) wstage[k] = 0.0;
    __syncwarp();
    if (j < ncols) {
      double *my = wstage + (colptr[j] - wbase);
      const int nq = (int)(blk_ptr[blk + 1] - row0);
      for (int q = 0; q < nq; q++) {
        const int32_t e = adjT_cell[(row0 + q) * 32 + lane];
        const uint64_t ranks = adjT_rank[(row0 + q) * 32 + lane];
        if (e < 0) continue;
        const int lj = e & 7;
        double vals[8];
        column_entries<FORM>(G, ncells, (int64_t)(e >> 3), lj, coef, vals);
        // vals[m] belongs to the row li = m ^ lj; one cell adds to a slot at most once, so the order inside this
        // loop does not affect the per-slot summation order (ascending cells)
#pragma unroll
        for (int m = 0; m < 8; m++) {
          const unsigned r = (unsigned)(ranks >> (8 * (m ^ lj))) & 0xFFu;
          if (r != 0xFFu) my[r] += vals[m];
        }
      }
    }
  }
  __syncwarp();
  double *out = nzval + wbase;
  if (add)
    for (int k = lane; k < wspan; k += 32) out[k] += wstage[k];
  else
    for (int k = lane; k < wspan; k += 32) out[k] = wstage[k];
  __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Pipelined variant: the geometry factors of a warp's NEXT block are gathered into shared memory with cp.async (LDGSTS,
// 8 bytes per lane and factor) while the current block is computed from the previous buffer.  This takes the DRAM latency
// of the indirect G loads off the critical path without spending registers on it (the register variant above is bound
// by long-scoreboard stalls at 16 warps per SM).  Per warp: two G buffers [8 cells][NA factors][32 lanes]; the buffer of
// the current block doubles as the staging area of the warp's nzval range once its factors are in registers.
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int FORM>
__device__ __forceinline__ void column_entries_smem(const double *__restrict__ gq, int lane, int lj, double coef, double *vals) {
  // gq = buffer of one incident cell: [NA][32]
  if (FORM == GB200_FORM_LAPLACIAN) {
    const double t0 = (lj & 1) ? 1.0 : -1.0, t1 = (lj & 2) ? 1.0 : -1.0, t2 = (lj & 4) ? 1.0 : -1.0;
    const double d0 = coef * gq[lane], d1 = coef * gq[32 + lane], d2 = coef * gq[64 + lane];
    const double o01 = 0.25 * coef * t0 * t1 * gq[96 + lane], o02 = 0.25 * coef * t0 * t2 * gq[128 + lane], o12 = 0.25 * coef * t1 * t2 * gq[160 + lane];
    vals[0] = lap_entry<+1, +1, +1>(d0, d1, d2, o01, o02, o12);
    vals[1] = lap_entry<-1, +1, +1>(d0, d1, d2, o01, o02, o12);
    vals[2] = lap_entry<+1, -1, +1>(d0, d1, d2, o01, o02, o12);
    vals[3] = lap_entry<-1, -1, +1>(d0, d1, d2, o01, o02, o12);
    vals[4] = lap_entry<+1, +1, -1>(d0, d1, d2, o01, o02, o12);
    vals[5] = lap_entry<-1, +1, -1>(d0, d1, d2, o01, o02, o12);
    vals[6] = lap_entry<+1, -1, -1>(d0, d1, d2, o01, o02, o12);
    vals[7] = lap_entry<-1, -1, -1>(d0, d1, d2, o01, o02, o12);
  } else {
    const double ad = coef * gq[lane];
    vals[0] = mass_entry<+1, +1, +1>(ad);
    vals[1] = mass_entry<-1, +1, +1>(ad);
    vals[2] = mass_entry<+1, -1, +1>(ad);
    vals[3] = mass_entry<-1, -1, +1>(ad);
    vals[4] = mass_entry<+1, +1, -1>(ad);
    vals[5] = mass_entry<-1, +1, -1>(ad);
    vals[6] = mass_entry<+1, -1, -1>(ad);
    vals[7] = mass_entry<-1, -1, -1>(ad);
  }
}

template <int FORM, int Q>
__device__ __forceinline__ void canon_cell_smem(const double *__restrict__ buf, int lane, double coef, double *acc) {
  constexpr int NA = FORM == GB200_FORM_LAPLACIAN ? 6 : 1;
  double vals[8];
  column_entries_smem<FORM>(buf + Q * NA * 32, lane, 7 - Q, coef, vals);
#pragma unroll
  for (int m = 0; m < 8; m++) acc[canon_rank(Q, m)] += vals[m];
}

constexpr int ASYNC_WARPS = 8;

template <int FORM>
__global__ void __launch_bounds__(ASYNC_WARPS * 32, 1)
    q1hex_gather_async_kernel(const int64_t *__restrict__ colptr, const int64_t *__restrict__ blk_ptr, const uint8_t *__restrict__ blk_flag,
                              const uint32_t *__restrict__ col_mask, const int32_t *__restrict__ blk_base, const int32_t *__restrict__ adjT_cell,
                              const uint64_t *__restrict__ adjT_rank, const double *__restrict__ G, int64_t ncells, int64_t ncols, double coef,
                              double *__restrict__ nzval, int add, int wbuf_doubles) {
  constexpr int NA = FORM == GB200_FORM_LAPLACIAN ? 6 : 1;
  constexpr int A0 = FORM == GB200_FORM_LAPLACIAN ? 0 : 6;
  extern __shared__ double smem_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *buf0 = smem_all + (size_t)warp * 2 * wbuf_doubles;
  double *buf1 = buf0 + wbuf_doubles;
  const int64_t nblocks = (ncols + 31) >> 5;
  const int64_t wstride = (int64_t)gridDim.x * ASYNC_WARPS;
  int64_t blk0 = (int64_t)blockIdx.x * ASYNC_WARPS + warp;
  if (blk0 >= nblocks) return;

  // cell entries (cell*8 + lj) of the 8 incident cells of this lane's column in block b (stencil blocks only)
  auto load_entries = [&](int64_t b, int flag, int32_t *e) {
    if (flag & 4) {
      const int4 *bb = reinterpret_cast<const int4 *>(blk_base + b * 8);
      const int4 b0 = __ldg(bb), b1 = __ldg(bb + 1);
      e[0] = b0.x + 8 * lane; e[1] = b0.y + 8 * lane; e[2] = b0.z + 8 * lane; e[3] = b0.w + 8 * lane;
      e[4] = b1.x + 8 * lane; e[5] = b1.y + 8 * lane; e[6] = b1.z + 8 * lane; e[7] = b1.w + 8 * lane;
    } else {
      const int32_t *rows = adjT_cell + __ldg(blk_ptr + b) * 32;
#pragma unroll
      for (int q = 0; q < 8; q++) e[q] = __ldg(rows + q * 32 + lane);
    }
  };
  auto issue_prefetch = [&](const int32_t *e, double *buf) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int64_t cell = e[q] >> 3;
#pragma unroll
      for (int a = 0; a < NA; a++) cp_async8(buf + (q * NA + a) * 32 + lane, G + (int64_t)(A0 + a) * ncells + cell);
    }
  };

  int flag0 = blk_flag[blk0];
  int32_t e1[8];
  if (flag0) { load_entries(blk0, flag0, e1); issue_prefetch(e1, buf0); }
  cp_async_commit();
  int64_t blk1 = blk0 + wstride;
  int flag1 = 0;
  if (blk1 < nblocks) { flag1 = blk_flag[blk1]; if (flag1) load_entries(blk1, flag1, e1); }

  for (int it = 0;; it++) {
    double *cur = (it & 1) ? buf1 : buf0, *nxt = (it & 1) ? buf0 : buf1;
    // 1. start the gather of the next block's factors; fetch the entries of the block after it
    if (blk1 < nblocks && flag1) issue_prefetch(e1, nxt);
    cp_async_commit();
    const int64_t blk2 = blk1 + wstride;
    int flag2 = 0;
    if (blk2 < nblocks) { flag2 = blk_flag[blk2]; if (flag2) load_entries(blk2, flag2, e1); }
    // 2. this block
    const int64_t jw0 = blk0 * 32, jw1 = min(jw0 + 32, ncols), j = jw0 + lane;
    const int64_t wbase = colptr[jw0];
    const int wspan = (int)(colptr[jw1] - wbase);
    cp_async_wait<1>();
    __syncwarp();
    if (flag0) {
      double acc[27];
#pragma unroll
      for (int r = 0; r < 27; r++) acc[r] = 0.0;
      canon_cell_smem<FORM, 0>(cur, lane, coef, acc);
      canon_cell_smem<FORM, 1>(cur, lane, coef, acc);
      canon_cell_smem<FORM, 2>(cur, lane, coef, acc);
      canon_cell_smem<FORM, 3>(cur, lane, coef, acc);
      canon_cell_smem<FORM, 4>(cur, lane, coef, acc);
      canon_cell_smem<FORM, 5>(cur, lane, coef, acc);
      canon_cell_smem<FORM, 6>(cur, lane, coef, acc);
      canon_cell_smem<FORM, 7>(cur, lane, coef, acc);
      __syncwarp();  // every lane has its factors in registers: the buffer becomes the staging area
      if ((flag0 & 3) == 1) {
        double *my = cur + 27 * lane;
#pragma unroll
        for (int r = 0; r < 27; r++) my[r] = acc[r];
      } else {
        const uint32_t mask = col_mask[j];
        double *my = cur + (colptr[j] - wbase);
#pragma unroll
        for (int r = 0; r < 27; r++)
          if ((mask >> r) & 1u) my[__popc(mask & ((1u << r) - 1u))] = acc[r];
      }
    } else {
      for (int k = lane; k < wspan; k += 32) cur[k] = 0.0;
      __syncwarp();
      if (j < ncols) {
        double *my = cur + (colptr[j] - wbase);
        const int64_t row0 = blk_ptr[blk0];
        const int nq = (int)(blk_ptr[blk0 + 1] - row0);
        for (int q = 0; q < nq; q++) {
          const int32_t e = adjT_cell[(row0 + q) * 32 + lane];
          const uint64_t ranks = adjT_rank[(row0 + q) * 32 + lane];
          if (e < 0) continue;
          const int lj = e & 7;
          double vals[8];
          column_entries<FORM>(G, ncells, (int64_t)(e >> 3), lj, coef, vals);
#pragma unroll
          for (int m = 0; m < 8; m++) {
            const unsigned r = (unsigned)(ranks >> (8 * (m ^ lj))) & 0xFFu;
            if (r != 0xFFu) my[r] += vals[m];
          }
        }
      }
    }
    __syncwarp();
    double *out = nzval + wbase;
    if (add)
      for (int k = lane; k < wspan; k += 32) out[k] += cur[k];
    else
      for (int k = lane; k < wspan; k += 32) out[k] = cur[k];
    __syncwarp();
    // 3. rotate
    blk0 = blk1; flag0 = flag1;
    blk1 = blk2; flag1 = flag2;
    if (blk0 >= nblocks) break;
  }
  cp_async_wait<0>();
}

__global__ void affine_check_kernel(const double *__restrict__ X, const int32_t *__restrict__ cell_nodes, int64_t ncells, int D, int nn,
                                    int *flag) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  // n-cube with first-axis-fastest vertices: affine  <=>  x_v = x_0 + sum_d bit_d(v) (x_{2^d} - x_0) for every vertex v
  const int32_t *cn = cell_nodes + c * nn;
  double x0[3], e[3][3], scale = 0.0;
  for (int d = 0; d < D; d++) x0[d] = X[(int64_t)cn[0] * D + d];
  for (int k = 0; k < D; k++)
    for (int d = 0; d < D; d++) {
      e[k][d] = X[(int64_t)cn[1 << k] * D + d] - x0[d];
      scale = fmax(scale, fabs(e[k][d]));
    }
  bool ok = true;
  for (int v = 0; v < nn; v++)
    for (int d = 0; d < D; d++) {
      double pred = x0[d];
      for (int k = 0; k < D; k++)
        if ((v >> k) & 1) pred += e[k][d];
      if (fabs(pred - X[(int64_t)cn[v] * D + d]) > 1e-13 * scale) ok = false;
    }
  if (!ok) atomicExch(flag, 0);
}

}  // namespace

int mesh_check_affine(gb200_mesh mesh) {
  if (mesh->affine >= 0) return mesh->affine;
  gb200_ctx ctx = mesh->ctx;
  if (mesh->celltype == GB200_TET4 || mesh->celltype == GB200_TRI3) return mesh->affine = 1;
  DevBuf<int> flag;
  int one = 1;
  flag.upload(&one, 1, ctx->stream);
  int grid = (int)((mesh->ncells + 255) / 256);
  affine_check_kernel<<<grid, 256, 0, ctx->stream>>>(mesh->X.p, mesh->cell_nodes.p, mesh->ncells, mesh->D, mesh->nn, flag.p);
  check_launch(ctx, "affine_check_kernel");
  int h = 0;
  flag.download(&h, ctx->stream);
  GB_CUDA(cudaStreamSynchronize(ctx->stream));
  return mesh->affine = h;
}

// The closed forms above assume M^{kl}_{ab} = sum_q w_q d_kN_a(q) d_lN_b(q) (and the mass analogue) take their exact
// Q1 values; verify that against the tabulation the host actually passed (any rule that integrates them exactly passes).
static bool tabulation_is_exact_q1(const gb200_refel_s *r) {
  if (r->D != 3 || r->nd != 8 || r->ncomp != 1) return false;
  for (int a = 0; a < 8; a++)
    for (int b = 0; b < 8; b++) {
      double mass = 0, M[3][3] = {{0}};
      for (int p = 0; p < r->np; p++) {
        mass += r->w[p] * r->N[p * 8 + a] * r->N[p * 8 + b];
        for (int k = 0; k < 3; k++)
          for (int l = 0; l < 3; l++) M[k][l] += r->w[p] * r->dN[(p * 8 + a) * 3 + k] * r->dN[(p * 8 + b) * 3 + l];
      }
      double m[3], s[3], cx[3], cy[3];
      for (int d = 0; d < 3; d++) {
        int ad = (a >> d) & 1, bd = (b >> d) & 1;
        m[d] = ad == bd ? 1.0 / 3 : 1.0 / 6;
        s[d] = ad == bd ? 1.0 : -1.0;
        cx[d] = ad ? 0.5 : -0.5;  // int N'_a N_b
        cy[d] = bd ? 0.5 : -0.5;  // int N_a N'_b
      }
      if (fabs(mass - m[0] * m[1] * m[2]) > 1e-13) return false;
      for (int k = 0; k < 3; k++)
        for (int l = 0; l < 3; l++) {
          double ex = 1.0;
          for (int d = 0; d < 3; d++) {
            if (k == l) ex *= (d == k) ? s[d] : m[d];
            else ex *= (d == k) ? cx[d] : (d == l) ? cy[d] : m[d];
          }
          if (fabs(M[k][l] - ex) > 1e-13) return false;
        }
    }
  return true;
}

bool gather_supported(gb200_plan plan, int form) {
  if (form != GB200_FORM_LAPLACIAN && form != GB200_FORM_MASS) return false;
  if (plan->nfields != 1 || plan->mesh->celltype != GB200_HEX8 || plan->NL != 8) return false;
  if (!plan->has_gather) return false;
  if (plan->gather_ok < 0) plan->gather_ok = (tabulation_is_exact_q1(plan->test[0]->refel) && mesh_check_affine(plan->mesh) != 0) ? 1 : 0;
  return plan->gather_ok == 1;
}

void launch_gather(gb200_plan plan, int form, const double *params, double *nzval, bool add) {
  gb200_ctx ctx = plan->ctx;
  const int64_t nc = plan->mesh->ncells;
  static const int variant = getenv("GB200_GATHER_VARIANT") ? atoi(getenv("GB200_GATHER_VARIANT")) : 1;
  if (plan->cellG.n != (size_t)(7 * nc)) plan->cellG.alloc((size_t)(7 * nc));
  {
    ScopedTimer t(ctx, "k:cell_geom");
    cell_geom_kernel<<<(int)((nc + 255) / 256), 256, 0, ctx->stream>>>(plan->mesh->X.p, plan->mesh->cell_nodes.p, nc, plan->cellG.p,
                                                                               form == GB200_FORM_MASS ? 1 : 0);
    check_launch(ctx, "cell_geom_kernel");
  }
  ScopedTimer t2(ctx, "k:q1hex_gather");
  static const int use_async = getenv("GB200_GATHER_ASYNC") ? atoi(getenv("GB200_GATHER_ASYNC")) : 1;
  if (use_async && variant != 0) {
    // per-warp buffer: room for 8 cells x NA factors x 32 lanes and for the warp's nzval range
    const int na = form == GB200_FORM_LAPLACIAN ? 6 : 1;
    const int wbuf = (int)std::max<int64_t>(8 * na * 32, (plan->gather_span_max + 1) & ~1ll);
    const size_t smem_async = (size_t)ASYNC_WARPS * 2 * wbuf * sizeof(double);
    if (smem_async <= 227 * 1024) {
      auto ak = form == GB200_FORM_LAPLACIAN ? q1hex_gather_async_kernel<GB200_FORM_LAPLACIAN> : q1hex_gather_async_kernel<GB200_FORM_MASS>;
      int &configured = plan->gather_ctas_per_sm[form == GB200_FORM_MASS ? 1 : 0];
      if (configured == 0) {
        GB_CUDA(cudaFuncSetAttribute(ak, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_async));
        configured = 1;
      }
      const int64_t nblk = (plan->ncols + 31) / 32;
      int grid = (int)std::min<int64_t>((nblk + ASYNC_WARPS - 1) / ASYNC_WARPS, (int64_t)ctx->num_sms);
      ak<<<grid, ASYNC_WARPS * 32, smem_async, ctx->stream>>>(plan->colptr.p, plan->blk_ptr.p, plan->blk_flag.p, plan->col_mask.p, plan->blk_base.p,
                                                             plan->adjT_cell.p, plan->adjT_rank.p, plan->cellG.p, nc, plan->ncols, params[0], nzval,
                                                             add ? 1 : 0, wbuf);
      check_launch(ctx, "q1hex_gather_async_kernel");
      return;
    }
  }
  const int wspan = (int)plan->gather_span_max;  // max nnz of one 32-column block
  size_t smem = (size_t)(GATHER_THREADS / 32) * wspan * sizeof(double);
  static const int minb = getenv("GB200_GATHER_MINB") ? atoi(getenv("GB200_GATHER_MINB")) : 4;
  static const int prefetch = getenv("GB200_GATHER_PREFETCH") ? atoi(getenv("GB200_GATHER_PREFETCH")) : 0;
  auto kern = form == GB200_FORM_MASS ? q1hex_gather_kernel<GB200_FORM_MASS, 4>
              : minb >= 6 ? q1hex_gather_kernel<GB200_FORM_LAPLACIAN, 6>
              : minb >= 4 ? q1hex_gather_kernel<GB200_FORM_LAPLACIAN, 4>
              : minb >= 3 ? q1hex_gather_kernel<GB200_FORM_LAPLACIAN, 3>
                          : q1hex_gather_kernel<GB200_FORM_LAPLACIAN, 2>;
  int &ctas_per_sm = plan->gather_ctas_per_sm[form == GB200_FORM_MASS ? 1 : 0];
  if (ctas_per_sm == 0) {  // once per plan and form: keeps the per-call host overhead to the two launches
    GB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, GATHER_THREADS, smem));
    ctas_per_sm = std::max(ctas_per_sm, 1);
  }
  const int64_t nblocks = (plan->ncols + 31) / 32;
  static const int oversub = getenv("GB200_GATHER_OVERSUB") ? atoi(getenv("GB200_GATHER_OVERSUB")) : 1;
  int grid = (int)std::min<int64_t>((nblocks + 3) / 4, (int64_t)ctx->num_sms * std::max(ctas_per_sm, 1) * oversub);
  kern<<<grid, GATHER_THREADS, smem, ctx->stream>>>(plan->colptr.p, plan->blk_ptr.p, plan->blk_flag.p, plan->col_mask.p, plan->blk_base.p, plan->adjT_cell.p, plan->adjT_rank.p,
                                                   plan->cellG.p, nc, plan->ncols, params[0], nzval, add ? 1 : 0, variant != 0, wspan, prefetch);
  check_launch(ctx, "q1hex_gather_kernel");
}

}  // namespace gb
