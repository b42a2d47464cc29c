// q1hex_fused.cu -- ONE persistent kernel for the headline path: geometry-producer warps and gather-consumer warps.
//
// The two-kernel version (cell_geom_kernel -> q1hex_gather_kernel) writes the per-cell factors G to HBM (0.77 GB) and reads
// them back (0.77 GB), and the DRAM-bound geometry pass (0.32 ms) cannot overlap the latency-bound gather (1.0 ms).
// Here every CTA has 4 gather warps (same code and summation order as q1hex_gather.cu: results are bitwise identical) and
// 1 geometry warp.  Geometry warps claim chunks of 1024 cells from a global counter in ascending order, compute G and
// publish a per-chunk flag (release); a gather warp acquires the flags of the chunks its block's cells live in before it
// loads their factors.  Consumers trail producers by a few MB of G, so the factor reads are L2 hits, and the geometry
// traffic overlaps the gather's latency stalls.  The grid never exceeds the resident capacity and chunks are claimed
// dynamically, so a spinning consumer can never starve the producer it waits for.
#include "common.cuh"
#include "q1hex_common.cuh"

namespace gb {

using namespace q1;

namespace {

constexpr int GWARPS = 4;                 // gather warps per CTA
constexpr int CHUNK_SHIFT = 10;           // 1024 cells per geometry chunk (8 KB per factor array: whole cache lines)

__device__ __forceinline__ int ld_acquire(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int *p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// The factors are written by other SMs during this launch: they are read at the coherence point (ld.global.cg, L2) after the
// chunk flag has been observed (volatile poll + control dependency), so no L1 invalidation / consumer-side fence is needed.
template <int FORM>
__device__ __forceinline__ void fused_entries(const double *G, int64_t ncells, int64_t cell, int lj, double coef, double *vals) {
  if (FORM == GB200_FORM_LAPLACIAN) {
    const double t0 = (lj & 1) ? 1.0 : -1.0, t1 = (lj & 2) ? 1.0 : -1.0, t2 = (lj & 4) ? 1.0 : -1.0;
    const double d0 = coef * __ldcg(G + cell), d1 = coef * __ldcg(G + ncells + cell), d2 = coef * __ldcg(G + 2 * ncells + cell);
    const double o01 = 0.25 * coef * t0 * t1 * __ldcg(G + 3 * ncells + cell), o02 = 0.25 * coef * t0 * t2 * __ldcg(G + 4 * ncells + cell),
                 o12 = 0.25 * coef * t1 * t2 * __ldcg(G + 5 * ncells + cell);
    vals[0] = lap_entry<+1, +1, +1>(d0, d1, d2, o01, o02, o12);
    vals[1] = lap_entry<-1, +1, +1>(d0, d1, d2, o01, o02, o12);
    vals[2] = lap_entry<+1, -1, +1>(d0, d1, d2, o01, o02, o12);
    vals[3] = lap_entry<-1, -1, +1>(d0, d1, d2, o01, o02, o12);
    vals[4] = lap_entry<+1, +1, -1>(d0, d1, d2, o01, o02, o12);
    vals[5] = lap_entry<-1, +1, -1>(d0, d1, d2, o01, o02, o12);
    vals[6] = lap_entry<+1, -1, -1>(d0, d1, d2, o01, o02, o12);
    vals[7] = lap_entry<-1, -1, -1>(d0, d1, d2, o01, o02, o12);
  } else {
    const double ad = coef * __ldcg(G + 6 * ncells + cell);
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
__device__ __forceinline__ void fused_cell(int32_t e, const double *G, int64_t ncells, double coef, double *acc) {
  double vals[8];
  fused_entries<FORM>(G, ncells, (int64_t)(e >> 3), 7 - Q, coef, vals);
#pragma unroll
  for (int m = 0; m < 8; m++) acc[canon_rank(Q, m)] += vals[m];
}

struct FusedArgs {
  const double *X;
  const int32_t *cell_nodes;
  const int64_t *colptr, *blk_ptr;
  const uint8_t *blk_flag;
  const uint32_t *col_mask;
  const int32_t *blk_base, *adjT_cell;
  const uint64_t *adjT_rank;
  double *G;
  int64_t ncells, ncols;
  double coef;
  double *nzval;
  int add, wspan_max, epoch;
  int *chunk_counter, *chunk_flags;
};

template <int FORM, int GEOW, int MINB>
__global__ void __launch_bounds__((GWARPS + GEOW) * 32, MINB) q1hex_fused_kernel(FusedArgs k) {
  extern __shared__ double stage[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t ncells = k.ncells;
  double *G = k.G;

  if (warp >= GWARPS) {
    // ------------------------------------------------------------------ geometry producer
    const int nchunks = (int)((ncells + (1 << CHUNK_SHIFT) - 1) >> CHUNK_SHIFT);
    for (;;) {
      int chunk = 0;
      if (lane == 0) chunk = atomicAdd(k.chunk_counter, 1);
      chunk = __shfl_sync(0xffffffffu, chunk, 0);
      if (chunk >= nchunks) break;
      const int64_t c0 = (int64_t)chunk << CHUNK_SHIFT, c1 = min(c0 + (1 << CHUNK_SHIFT), ncells);
      // 4 cells per lane and iteration (128 cells per warp iteration) for memory-level parallelism; an affine cell needs only
      // the corner node and its three edge neighbours: Jt[i][:] = x_{2^i} - x_0
      for (int64_t cb = c0; cb < c1; cb += 128) {
        double x[4][4][3];
        bool valid[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int64_t c = cb + lane + 32 * u;
          valid[u] = c < c1;
          const int64_t cc = valid[u] ? c : c0;
          const int4 n0 = __ldg(reinterpret_cast<const int4 *>(k.cell_nodes + cc * 8));
          const int n4 = __ldg(k.cell_nodes + cc * 8 + 4);
          const int ids[4] = {n0.x, n0.y, n0.z, n4};
#pragma unroll
          for (int a = 0; a < 4; a++) {
            const double *p = k.X + (int64_t)ids[a] * 3;
            x[u][a][0] = __ldg(p); x[u][a][1] = __ldg(p + 1); x[u][a][2] = __ldg(p + 2);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          if (!valid[u]) continue;
          const int64_t c = cb + lane + 32 * u;
          double J[9];
#pragma unroll
          for (int d = 0; d < 3; d++) {
            J[0 + d] = x[u][1][d] - x[u][0][d];
            J[3 + d] = x[u][2][d] - x[u][0][d];
            J[6 + d] = x[u][3][d] - x[u][0][d];
          }
          const double det = J[0] * J[4] * J[8] + J[1] * J[5] * J[6] + J[2] * J[3] * J[7] - (J[0] * J[5] * J[7] + J[1] * J[3] * J[8] + J[2] * J[4] * J[6]);
          const double ci = 1.0 / det;
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
          const double ad = fabs(det);
          if (FORM == GB200_FORM_LAPLACIAN) {
            G[c] = ad * (I[0] * I[0] + I[3] * I[3] + I[6] * I[6]);
            G[ncells + c] = ad * (I[1] * I[1] + I[4] * I[4] + I[7] * I[7]);
            G[2 * ncells + c] = ad * (I[2] * I[2] + I[5] * I[5] + I[8] * I[8]);
            G[3 * ncells + c] = ad * (I[0] * I[1] + I[3] * I[4] + I[6] * I[7]);
            G[4 * ncells + c] = ad * (I[0] * I[2] + I[3] * I[5] + I[6] * I[8]);
            G[5 * ncells + c] = ad * (I[1] * I[2] + I[4] * I[5] + I[7] * I[8]);
          } else {
            G[6 * ncells + c] = ad;
          }
        }
      }
      __threadfence();   // every lane: its factor stores are visible device-wide before the flag
      __syncwarp();
      if (lane == 0) st_release(k.chunk_flags + chunk, k.epoch);
    }
    return;
  }

  // -------------------------------------------------------------------- gather consumers (persistent warps)
  double *wstage = stage + (size_t)warp * k.wspan_max;
  const int64_t ncols = k.ncols;
  const int64_t nblocks = (ncols + 31) >> 5;
  const int64_t wstride = (int64_t)gridDim.x * GWARPS;
  int64_t blk = (int64_t)blockIdx.x * GWARPS + warp;
  if (blk >= nblocks) return;
  int64_t n_wbase = k.colptr[blk * 32], n_wend = k.colptr[min(blk * 32 + 32, ncols)];
  int n_flag = k.blk_flag[blk];
  int4 n_b0 = make_int4(0, 0, 0, 0), n_b1 = n_b0;
  {
    const int4 *bb = reinterpret_cast<const int4 *>(k.blk_base + blk * 8);
    n_b0 = __ldg(bb);
    n_b1 = __ldg(bb + 1);
  }
  for (; blk < nblocks; blk += wstride) {
    const int64_t jw0 = blk * 32;
    const int64_t wbase = n_wbase;
    const int wspan = (int)(n_wend - n_wbase);
    const int64_t j = jw0 + lane;
    const int flag = n_flag;
    const int4 b0 = n_b0, b1 = n_b1;
    {
      const int64_t nb = blk + wstride;
      if (nb < nblocks) {
        n_wbase = k.colptr[nb * 32];
        n_wend = k.colptr[min(nb * 32 + 32, ncols)];
        n_flag = k.blk_flag[nb];
        const int4 *bb = reinterpret_cast<const int4 *>(k.blk_base + nb * 8);
        n_b0 = __ldg(bb);
        n_b1 = __ldg(bb + 1);
      }
    }
    if (flag) {
      int32_t e[8];
      if (flag & 4) {
        e[0] = b0.x + 8 * lane; e[1] = b0.y + 8 * lane; e[2] = b0.z + 8 * lane; e[3] = b0.w + 8 * lane;
        e[4] = b1.x + 8 * lane; e[5] = b1.y + 8 * lane; e[6] = b1.z + 8 * lane; e[7] = b1.w + 8 * lane;
      } else {
        const int32_t *rows = k.adjT_cell + k.blk_ptr[blk] * 32;
#pragma unroll
        for (int q = 0; q < 8; q++) e[q] = __ldg(rows + q * 32 + lane);
      }
      // wait for the geometry of this block's cells.  Run-compressed rows span at most two chunks (first / last lane):
      // lanes 0..15 poll one flag each; otherwise every lane polls the chunks of its own 8 cells.  One fence afterwards.
      if (flag & 4) {
        if (lane < 16) {
          const int q = lane & 7;
          const int base = q == 0 ? b0.x : q == 1 ? b0.y : q == 2 ? b0.z : q == 3 ? b0.w : q == 4 ? b1.x : q == 5 ? b1.y : q == 6 ? b1.z : b1.w;
          const volatile int *fl = k.chunk_flags + (((base >> 3) + ((lane >> 3) ? 31 : 0)) >> CHUNK_SHIFT);
          while (*fl != k.epoch) __nanosleep(64);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const volatile int *fl = k.chunk_flags + ((e[q] >> 3) >> CHUNK_SHIFT);
          while (*fl != k.epoch) __nanosleep(64);
        }
      }
      __syncwarp();
      double acc[27];
#pragma unroll
      for (int r = 0; r < 27; r++) acc[r] = 0.0;
      fused_cell<FORM, 0>(e[0], G, ncells, k.coef, acc);
      fused_cell<FORM, 1>(e[1], G, ncells, k.coef, acc);
      fused_cell<FORM, 2>(e[2], G, ncells, k.coef, acc);
      fused_cell<FORM, 3>(e[3], G, ncells, k.coef, acc);
      fused_cell<FORM, 4>(e[4], G, ncells, k.coef, acc);
      fused_cell<FORM, 5>(e[5], G, ncells, k.coef, acc);
      fused_cell<FORM, 6>(e[6], G, ncells, k.coef, acc);
      fused_cell<FORM, 7>(e[7], G, ncells, k.coef, acc);
      if ((flag & 3) == 1) {
        double *my = wstage + 27 * lane;
#pragma unroll
        for (int r = 0; r < 27; r++) my[r] = acc[r];
      } else {
        const uint32_t mask = k.col_mask[j];
        double *my = wstage + (k.colptr[j] - wbase);
#pragma unroll
        for (int r = 0; r < 27; r++)
          if ((mask >> r) & 1u) my[__popc(mask & ((1u << r) - 1u))] = acc[r];
      }
    } else {
      for (int i = lane; i < wspan; i += 32) wstage[i] = 0.0;
      __syncwarp();
      if (j < ncols) {
        double *my = wstage + (k.colptr[j] - wbase);
        const int64_t row0 = k.blk_ptr[blk];
        const int nq = (int)(k.blk_ptr[blk + 1] - row0);
        for (int q = 0; q < nq; q++) {
          const int32_t e = k.adjT_cell[(row0 + q) * 32 + lane];
          const uint64_t ranks = k.adjT_rank[(row0 + q) * 32 + lane];
          if (e < 0) continue;
          const volatile int *fl = k.chunk_flags + ((e >> 3) >> CHUNK_SHIFT);
          while (*fl != k.epoch) __nanosleep(64);
          const int lj = e & 7;
          double vals[8];
          fused_entries<FORM>(G, ncells, (int64_t)(e >> 3), lj, k.coef, vals);
#pragma unroll
          for (int m = 0; m < 8; m++) {
            const unsigned r = (unsigned)(ranks >> (8 * (m ^ lj))) & 0xFFu;
            if (r != 0xFFu) my[r] += vals[m];
          }
        }
      }
    }
    __syncwarp();
    double *out = k.nzval + wbase;
    if (k.add)
      for (int i = lane; i < wspan; i += 32) out[i] += wstage[i];
    else
      for (int i = lane; i < wspan; i += 32) out[i] = wstage[i];
    __syncwarp();
  }
}

}  // namespace

bool launch_gather_fused(gb200_plan plan, int form, double coef, double *nzval, bool add) {
  gb200_ctx ctx = plan->ctx;
  const int64_t nc = plan->mesh->ncells;
  const int wspan = (int)plan->gather_span_max;
  const size_t smem = (size_t)GWARPS * wspan * sizeof(double);
  if (smem > 64 * 1024) return false;
  static const int geow = getenv("GB200_FUSED_GEOW") ? atoi(getenv("GB200_FUSED_GEOW")) : 1;
  static const int minb = getenv("GB200_FUSED_MINB") ? atoi(getenv("GB200_FUSED_MINB")) : 3;
  const int threads = (GWARPS + (geow >= 2 ? 2 : 1)) * 32;
  auto kern = form == GB200_FORM_MASS ? q1hex_fused_kernel<GB200_FORM_MASS, 1, 3>
              : geow >= 2 ? (minb >= 3 ? q1hex_fused_kernel<GB200_FORM_LAPLACIAN, 2, 3> : q1hex_fused_kernel<GB200_FORM_LAPLACIAN, 2, 2>)
                          : (minb >= 3 ? q1hex_fused_kernel<GB200_FORM_LAPLACIAN, 1, 3> : q1hex_fused_kernel<GB200_FORM_LAPLACIAN, 1, 2>);
  const int threads_used = form == GB200_FORM_MASS ? (GWARPS + 1) * 32 : threads;
  int &cps = plan->fused_ctas_per_sm[form == GB200_FORM_MASS ? 1 : 0];
  if (cps == 0) {
    GB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cps, kern, threads_used, smem));
    if (cps < 1) return false;
  }
  const int nchunks = (int)((nc + (1 << CHUNK_SHIFT) - 1) >> CHUNK_SHIFT);
  if (plan->chunk_sync.n != (size_t)(nchunks + 1)) {
    plan->chunk_sync.alloc((size_t)nchunks + 1);
    plan->chunk_sync.zero(ctx->stream);
    plan->fused_epoch = 0;
  }
  if (plan->cellG.n != (size_t)(7 * nc)) plan->cellG.alloc((size_t)(7 * nc));
  plan->fused_epoch += 1;
  GB_CUDA(cudaMemsetAsync(plan->chunk_sync.p, 0, sizeof(int), ctx->stream));  // chunk counter; the flags carry the epoch
  FusedArgs k;
  k.X = plan->mesh->X.p; k.cell_nodes = plan->mesh->cell_nodes.p; k.colptr = plan->colptr.p; k.blk_ptr = plan->blk_ptr.p;
  k.blk_flag = plan->blk_flag.p; k.col_mask = plan->col_mask.p; k.blk_base = plan->blk_base.p; k.adjT_cell = plan->adjT_cell.p;
  k.adjT_rank = plan->adjT_rank.p; k.G = plan->cellG.p; k.ncells = nc; k.ncols = plan->ncols; k.coef = coef; k.nzval = nzval;
  k.add = add ? 1 : 0; k.wspan_max = wspan; k.epoch = plan->fused_epoch;
  k.chunk_counter = plan->chunk_sync.p; k.chunk_flags = plan->chunk_sync.p + 1;
  const int64_t nblocks = (plan->ncols + 31) / 32;
  // the grid must be co-resident (consumers spin on producers): never more CTAs than the device can hold at once
  int grid = (int)std::min<int64_t>((nblocks + GWARPS - 1) / GWARPS, (int64_t)ctx->num_sms * cps);
  ScopedTimer t(ctx, "k:q1hex_fused");
  kern<<<grid, threads_used, smem, ctx->stream>>>(k);
  check_launch(ctx, "q1hex_fused_kernel");
  return true;
}

}  // namespace gb
