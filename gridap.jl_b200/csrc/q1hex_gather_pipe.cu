// q1hex_gather_pipe.cu -- software-pipelined owner-computes gather (scalar Q1 hexahedra, affine cells).
//
// Same algorithm and same per-slot summation order as q1hex_gather.cu (results are bitwise identical); what changes is
// how the per-cell geometry factors reach the SM.  The register kernel is bound by long-scoreboard stalls on the indirect
// G loads (ncu: 5.9 stalled warps per issue, DRAM at 66 % of peak).  Here each persistent warp gathers the factors of its
// NEXT 32-column block into shared memory with cp.async (LDGSTS, no registers held) while it computes the current block
// from the previous buffer, so the DRAM latency is off the critical path at an unchanged 16 warps per SM.
//
// It applies to "paired-run" stencil blocks (plan flag bit 8): the 8 incident-cell rows of the block are 4 pairs of runs
// [c_k, c_k+32) / [c_k+1, c_k+33) of consecutive cells, i.e. 4 x 33 distinct cells instead of 8 x 32 -- detected from the
// adjacency in the plan, not assumed from the mesh type.  Other blocks take the register / generic path of the same warp.
// Per warp: two buffers of 864 doubles; the current buffer doubles as the staging area for the warp's nzval range.
#include "common.cuh"
#include "q1hex_common.cuh"

namespace gb {

using namespace q1;

namespace {

constexpr int PIPE_WARPS = 4;
constexpr int WBUF = 864;  // >= 27*32 (staging) and >= 4*6*33 = 792 (factors)

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// entries K_e[:, lj] of the cell whose factors sit at gq[a*33 + idx]
template <int FORM>
__device__ __forceinline__ void entries_from_smem(const double *__restrict__ gq, int idx, int lj, double coef, double *vals) {
  if (FORM == GB200_FORM_LAPLACIAN) {
    const double t0 = (lj & 1) ? 1.0 : -1.0, t1 = (lj & 2) ? 1.0 : -1.0, t2 = (lj & 4) ? 1.0 : -1.0;
    const double d0 = coef * gq[idx], d1 = coef * gq[33 + idx], d2 = coef * gq[66 + idx];
    const double o01 = 0.25 * coef * t0 * t1 * gq[99 + idx], o02 = 0.25 * coef * t0 * t2 * gq[132 + idx], o12 = 0.25 * coef * t1 * t2 * gq[165 + idx];
    vals[0] = lap_entry<+1, +1, +1>(d0, d1, d2, o01, o02, o12);
    vals[1] = lap_entry<-1, +1, +1>(d0, d1, d2, o01, o02, o12);
    vals[2] = lap_entry<+1, -1, +1>(d0, d1, d2, o01, o02, o12);
    vals[3] = lap_entry<-1, -1, +1>(d0, d1, d2, o01, o02, o12);
    vals[4] = lap_entry<+1, +1, -1>(d0, d1, d2, o01, o02, o12);
    vals[5] = lap_entry<-1, +1, -1>(d0, d1, d2, o01, o02, o12);
    vals[6] = lap_entry<+1, -1, -1>(d0, d1, d2, o01, o02, o12);
    vals[7] = lap_entry<-1, -1, -1>(d0, d1, d2, o01, o02, o12);
  } else {
    const double ad = coef * gq[idx];
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
__device__ __forceinline__ void pipe_cell(const double *__restrict__ buf, int lane, double coef, double *acc) {
  constexpr int NA = FORM == GB200_FORM_LAPLACIAN ? 6 : 1;
  double vals[8];
  entries_from_smem<FORM>(buf + (Q >> 1) * NA * 33, lane + (Q & 1), 7 - Q, coef, vals);
#pragma unroll
  for (int m = 0; m < 8; m++) acc[canon_rank(Q, m)] += vals[m];
}

template <int FORM, int Q>
__device__ __forceinline__ void reg_cell(int32_t e, const double *__restrict__ G, int64_t ncells, double coef, double *acc) {
  double vals[8];
  column_entries<FORM>(G, ncells, (int64_t)(e >> 3), 7 - Q, coef, vals);
#pragma unroll
  for (int m = 0; m < 8; m++) acc[canon_rank(Q, m)] += vals[m];
}

template <int FORM>
__global__ void __launch_bounds__(PIPE_WARPS * 32, 4)
    q1hex_gather_pipe_kernel(const int64_t *__restrict__ colptr, const int64_t *__restrict__ blk_ptr, const uint8_t *__restrict__ blk_flag,
                             const uint32_t *__restrict__ col_mask, const int32_t *__restrict__ blk_base, const int32_t *__restrict__ blk_pair,
                             const int32_t *__restrict__ adjT_cell, const uint64_t *__restrict__ adjT_rank, const double *__restrict__ G,
                             int64_t ncells, int64_t ncols, double coef, double *__restrict__ nzval, int add, int wbuf) {
  constexpr int NA = FORM == GB200_FORM_LAPLACIAN ? 6 : 1;
  constexpr int A0 = FORM == GB200_FORM_LAPLACIAN ? 0 : 6;
  extern __shared__ double smem_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *buf0 = smem_all + (size_t)warp * 2 * wbuf;
  double *buf1 = buf0 + wbuf;
  const int64_t nblocks = (ncols + 31) >> 5;
  const int64_t wstride = (int64_t)gridDim.x * PIPE_WARPS;
  int64_t blk0 = (int64_t)blockIdx.x * PIPE_WARPS + warp;
  if (blk0 >= nblocks) return;

  // gather the factors of the 4 x 33 distinct cells of a paired-run block: [pair k][factor a][33]
  auto issue_prefetch = [&](int4 c, double *buf) {
    const int cells[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
#pragma unroll
      for (int a = 0; a < NA; a++) {
        const double *src = G + (int64_t)(A0 + a) * ncells + cells[k];
        cp_async8(buf + (k * NA + a) * 33 + lane, src + lane);
        if (lane == 0) cp_async8(buf + (k * NA + a) * 33 + 32, src + 32);
      }
    }
  };

  int flag0 = blk_flag[blk0];
  int4 pair1 = make_int4(0, 0, 0, 0);
  if (flag0 & 8) issue_prefetch(__ldg(reinterpret_cast<const int4 *>(blk_pair) + blk0), buf0);
  cp_async_commit();
  int64_t blk1 = blk0 + wstride;
  int flag1 = 0;
  if (blk1 < nblocks) {
    flag1 = blk_flag[blk1];
    if (flag1 & 8) pair1 = __ldg(reinterpret_cast<const int4 *>(blk_pair) + blk1);
  }

  for (int it = 0;; it++) {
    double *cur = (it & 1) ? buf1 : buf0, *nxt = (it & 1) ? buf0 : buf1;
    // 1. start gathering the next block's factors; fetch the cell bases of the block after it
    if (flag1 & 8) issue_prefetch(pair1, nxt);
    cp_async_commit();
    const int64_t blk2 = blk1 + wstride;
    int flag2 = 0;
    if (blk2 < nblocks) {
      flag2 = blk_flag[blk2];
      if (flag2 & 8) pair1 = __ldg(reinterpret_cast<const int4 *>(blk_pair) + blk2);
    }
    // 2. this block
    const int64_t jw0 = blk0 * 32, jw1 = min(jw0 + 32, ncols), j = jw0 + lane;
    const int64_t wbase = colptr[jw0];
    const int wspan = (int)(colptr[jw1] - wbase);
    cp_async_wait<1>();
    __syncwarp();
    if (flag0 & 3) {
      double acc[27];
#pragma unroll
      for (int r = 0; r < 27; r++) acc[r] = 0.0;
      if (flag0 & 8) {
        pipe_cell<FORM, 0>(cur, lane, coef, acc);
        pipe_cell<FORM, 1>(cur, lane, coef, acc);
        pipe_cell<FORM, 2>(cur, lane, coef, acc);
        pipe_cell<FORM, 3>(cur, lane, coef, acc);
        pipe_cell<FORM, 4>(cur, lane, coef, acc);
        pipe_cell<FORM, 5>(cur, lane, coef, acc);
        pipe_cell<FORM, 6>(cur, lane, coef, acc);
        pipe_cell<FORM, 7>(cur, lane, coef, acc);
      } else {
        const int32_t *rows = adjT_cell + blk_ptr[blk0] * 32;
        int32_t e[8];
#pragma unroll
        for (int q = 0; q < 8; q++) e[q] = __ldg(rows + q * 32 + lane);
        reg_cell<FORM, 0>(e[0], G, ncells, coef, acc);
        reg_cell<FORM, 1>(e[1], G, ncells, coef, acc);
        reg_cell<FORM, 2>(e[2], G, ncells, coef, acc);
        reg_cell<FORM, 3>(e[3], G, ncells, coef, acc);
        reg_cell<FORM, 4>(e[4], G, ncells, coef, acc);
        reg_cell<FORM, 5>(e[5], G, ncells, coef, acc);
        reg_cell<FORM, 6>(e[6], G, ncells, coef, acc);
        reg_cell<FORM, 7>(e[7], G, ncells, coef, acc);
      }
      __syncwarp();  // every lane has consumed its factors: the buffer becomes the staging area
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

}  // namespace

bool launch_gather_pipelined(gb200_plan plan, int form, double coef, double *nzval, bool add) {
  gb200_ctx ctx = plan->ctx;
  const int64_t nblk = (plan->ncols + 31) / 32;
  if (plan->n_paired_blocks * 2 < nblk) return false;  // mostly irregular adjacency: the register kernel is the better fit
  const int wbuf = (int)std::max<int64_t>(WBUF, (plan->gather_span_max + 1) & ~1ll);
  const size_t smem = (size_t)PIPE_WARPS * 2 * wbuf * sizeof(double);
  if (smem > 100 * 1024) return false;
  auto kern = form == GB200_FORM_LAPLACIAN ? q1hex_gather_pipe_kernel<GB200_FORM_LAPLACIAN> : q1hex_gather_pipe_kernel<GB200_FORM_MASS>;
  int &ctas_per_sm = plan->pipe_ctas_per_sm[form == GB200_FORM_MASS ? 1 : 0];
  if (ctas_per_sm == 0) {
    GB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, PIPE_WARPS * 32, smem));
    ctas_per_sm = std::max(ctas_per_sm, 1);
  }
  int grid = (int)std::min<int64_t>((nblk + PIPE_WARPS - 1) / PIPE_WARPS, (int64_t)ctx->num_sms * ctas_per_sm);
  kern<<<grid, PIPE_WARPS * 32, smem, ctx->stream>>>(plan->colptr.p, plan->blk_ptr.p, plan->blk_flag.p, plan->col_mask.p, plan->blk_base.p,
                                                    plan->blk_pair.p, plan->adjT_cell.p, plan->adjT_rank.p, plan->cellG.p, plan->mesh->ncells,
                                                    plan->ncols, coef, nzval, add ? 1 : 0, wbuf);
  check_launch(ctx, "q1hex_gather_pipe_kernel");
  return true;
}

}  // namespace gb
