// symbolic.cu -- the symbolic phase on the device: CSC pattern + cell->slot maps.
//
// Replaces nz_counter -> symbolic_loop_matrix! -> nz_allocation -> (insertion) -> create_from_nz
// (reference: src/FESpaces/SparseMatrixAssemblers.jl:174-210, src/Algebra/SparseMatrixCSC.jl:72-283).
// The reference counts an upper bound per column (one per admissible (i,j) pair, duplicates included),
// allocates that, inserts with a per-entry binary search and finally compacts.  The data-parallel
// statement of the same thing: count the same bound, fill the candidate rows of every column,
// sort + unique each column segment, compact.  The result is the canonical CSC (rows ascending and
// unique per column), i.e. bit-identical to the reference's colptr/rowval.
#include <cub/device/device_scan.cuh>
#include <cub/device/device_reduce.cuh>

#include <thread>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

#include "common.cuh"

namespace gb {

namespace {

struct SymDesc {
  int nfields, NL;
  int nld[MAX_FIELDS], lofs[MAX_FIELDS];
  const int32_t *row_ids[MAX_FIELDS], *col_ids[MAX_FIELDS];
  int64_t row_off[MAX_FIELDS], col_off[MAX_FIELDS];
  unsigned char touched[MAX_FIELDS][MAX_FIELDS];
  int64_t ncells;
};

// which field does concatenated local index l belong to
__device__ __forceinline__ int field_of(const SymDesc &d, int l) { return (d.nfields > 1 && l >= d.lofs[1]) ? 1 : 0; }

// number of admissible rows of `cell` for a column of field bj
__device__ int admissible_rows(const SymDesc &d, int64_t cell, int bj) {
  int cnt = 0;
  for (int bi = 0; bi < d.nfields; bi++) {
    if (!d.touched[bi][bj]) continue;
    const int32_t *r = d.row_ids[bi] + cell * d.nld[bi];
    for (int k = 0; k < d.nld[bi]; k++) cnt += (r[k] > 0);
  }
  return cnt;
}

__global__ void count_bound_kernel(SymDesc d, unsigned long long *bound) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t total = d.ncells * d.NL;
  for (; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t cell = t / d.NL;
    int lj = (int)(t % d.NL);
    int bj = field_of(d, lj);
    int32_t j = d.col_ids[bj][cell * d.nld[bj] + (lj - d.lofs[bj])];
    if (j <= 0) continue;
    int cnt = admissible_rows(d, cell, bj);
    if (cnt) atomicAdd(&bound[j - 1 + d.col_off[bj]], (unsigned long long)cnt);
  }
}

__global__ void fill_candidates_kernel(SymDesc d, const int64_t *tmp_ptr, unsigned long long *cursor, int32_t *tmp_rows) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t total = d.ncells * d.NL;
  for (; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t cell = t / d.NL;
    int lj = (int)(t % d.NL);
    int bj = field_of(d, lj);
    int32_t j = d.col_ids[bj][cell * d.nld[bj] + (lj - d.lofs[bj])];
    if (j <= 0) continue;
    int cnt = admissible_rows(d, cell, bj);
    if (!cnt) continue;
    int64_t col = j - 1 + d.col_off[bj];
    int64_t pos = tmp_ptr[col] + (int64_t)atomicAdd(&cursor[col], (unsigned long long)cnt);
    for (int bi = 0; bi < d.nfields; bi++) {
      if (!d.touched[bi][bj]) continue;
      const int32_t *r = d.row_ids[bi] + cell * d.nld[bi];
      for (int k = 0; k < d.nld[bi]; k++)
        if (r[k] > 0) tmp_rows[pos++] = (int32_t)(r[k] - 1 + d.row_off[bi]);
    }
  }
}

// One warp per column: bitonic sort of the candidate rows in shared memory, unique, write back in place.
template <int WARPS>
__global__ void sort_unique_kernel(const int64_t *tmp_ptr, int32_t *tmp_rows, int64_t *uniq, int64_t ncols, int cap) {
  extern __shared__ int32_t sm[];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int32_t *buf = sm + (size_t)warp * cap;
  for (int64_t j = blockIdx.x * (int64_t)WARPS + warp; j < ncols; j += (int64_t)gridDim.x * WARPS) {
    int64_t beg = tmp_ptr[j];
    int L = (int)(tmp_ptr[j + 1] - beg);
    if (L == 0) {
      if (lane == 0) uniq[j] = 0;
      continue;
    }
    int P = 32;
    while (P < L) P <<= 1;
    for (int k = lane; k < P; k += 32) buf[k] = k < L ? tmp_rows[beg + k] : 0x7fffffff;
    __syncwarp();
    for (int size = 2; size <= P; size <<= 1)
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int k = lane; k < (P >> 1); k += 32) {
          int lo = 2 * k - (k & (stride - 1));  // index with bit `stride` cleared
          int hi = lo + stride;
          bool up = ((lo & size) == 0);
          int32_t a = buf[lo], b = buf[hi];
          if ((a > b) == up) { buf[lo] = b; buf[hi] = a; }
        }
        __syncwarp();
      }
    int base = 0;
    for (int k0 = 0; k0 < P; k0 += 32) {
      int k = k0 + lane;
      int32_t v = buf[k];
      bool keep = (v != 0x7fffffff) && (k == 0 || buf[k - 1] != v);
      unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) tmp_rows[beg + base + __popc(m & ((1u << lane) - 1))] = v;
      base += __popc(m);
    }
    if (lane == 0) uniq[j] = base;
    __syncwarp();
  }
}

__global__ void compact_kernel(const int64_t *tmp_ptr, const int32_t *tmp_rows, const int64_t *colptr, int32_t *rowval,
                               int64_t ncols) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int64_t j = warp; j < ncols; j += nwarps) {
    int64_t src = tmp_ptr[j], dst = colptr[j];
    int L = (int)(colptr[j + 1] - dst);
    for (int k = lane; k < L; k += 32) rowval[dst + k] = tmp_rows[src + k];
  }
}

// rank[cell][lj][li] = position of row(li) inside column(lj), 0xFFFF when the entry is not stored
__global__ void rank_kernel(SymDesc d, const int64_t *colptr, const int32_t *rowval, uint16_t *rank) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t total = d.ncells * d.NL;
  for (; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t cell = t / d.NL;
    int lj = (int)(t % d.NL);
    int bj = field_of(d, lj);
    int32_t j = d.col_ids[bj][cell * d.nld[bj] + (lj - d.lofs[bj])];
    uint16_t *out = rank + t * d.NL;
    int64_t beg = 0;
    int len = 0;
    if (j > 0) {
      int64_t col = j - 1 + d.col_off[bj];
      beg = colptr[col];
      len = (int)(colptr[col + 1] - beg);
    }
    for (int bi = 0; bi < d.nfields; bi++) {
      const int32_t *r = d.row_ids[bi] + cell * d.nld[bi];
      for (int k = 0; k < d.nld[bi]; k++) {
        uint16_t res = 0xFFFF;
        if (j > 0 && d.touched[bi][bj] && r[k] > 0) {
          int32_t row = (int32_t)(r[k] - 1 + d.row_off[bi]);
          int lo = 0, hi = len;
          while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (rowval[beg + mid] < row) lo = mid + 1; else hi = mid;
          }
          res = (uint16_t)lo;  // present by construction
        }
        out[d.lofs[bi] + k] = res;
      }
    }
  }
}

__global__ void to_one_based_kernel(const int64_t *colptr, const int32_t *rowval, int64_t *colptr1, int64_t *rowval1,
                                    int64_t ncols, int64_t nnz) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (int64_t k = t; k < nnz; k += (int64_t)gridDim.x * blockDim.x) rowval1[k] = (int64_t)rowval[k] + 1;
  for (int64_t k = t; k <= ncols; k += (int64_t)gridDim.x * blockDim.x) colptr1[k] = colptr[k] + 1;
}

// ---- adjacency (column -> incident (cell, lj)) for the owner-computes gather path (nld <= 8, one field)
__global__ void adj_count_kernel(const int32_t *col_ids, int64_t n, unsigned long long *cnt) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    int32_t j = col_ids[t];
    if (j > 0) atomicAdd(&cnt[j - 1], 1ull);
  }
}
__global__ void adj_fill_kernel(const int32_t *col_ids, int64_t n, const int64_t *adj_ptr, unsigned long long *cursor,
                                int32_t *adj) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    int32_t j = col_ids[t];
    if (j > 0) adj[adj_ptr[j - 1] + (int64_t)atomicAdd(&cursor[j - 1], 1ull)] = (int32_t)t;  // t = cell*nld + lj
  }
}
__global__ void adj_sort_pack_kernel(const int64_t *adj_ptr, int32_t *adj, const uint16_t *rank, int nld, uint64_t *adj_rank,
                                     int64_t ncols) {
  int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (; j < ncols; j += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = adj_ptr[j], e = adj_ptr[j + 1];
    for (int64_t k = b + 1; k < e; k++) {  // insertion sort: ascending cell order = the reference's summation order
      int32_t v = adj[k];
      int64_t m = k - 1;
      while (m >= b && adj[m] > v) { adj[m + 1] = adj[m]; m--; }
      adj[m + 1] = v;
    }
    for (int64_t k = b; k < e; k++) {
      const uint16_t *r = rank + (int64_t)adj[k] * nld;
      uint64_t packed = 0;
      for (int li = 0; li < 8; li++) {
        uint64_t v = (li < nld && r[li] != 0xFFFF) ? (uint64_t)(r[li] & 0xFF) : 0xFFull;
        packed |= v << (8 * li);
      }
      adj_rank[k] = packed;
    }
  }
}

__global__ void span_max_kernel(const int64_t *colptr, int64_t ncols, int cols_per_cta, unsigned long long *out) {
  int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t j0 = b * cols_per_cta;
  if (j0 >= ncols) return;
  int64_t j1 = min(j0 + (int64_t)cols_per_cta, ncols);
  atomicMax(out, (unsigned long long)(colptr[j1] - colptr[j0]));
}

// ---- blocked-transposed adjacency: 32 columns per block (one warp), entries [block row q][lane] so that every
// load of the gather kernel is one coalesced 128/256-byte request.  A block is "canonical" when all its columns
// look like an interior node of a structured hexahedral patch: 8 incident cells, lj = 7 - q, 27 stored rows and
// the in-column ranks of a 3x3x3 stencil.  The gather kernel then accumulates in registers with static indices.
__global__ void blk_count_kernel(const int64_t *adj_ptr, const int32_t *adj, const uint64_t *adj_rank, const int64_t *colptr,
                                 int64_t ncols, int64_t nblocks, int64_t *blk_nq, uint8_t *blk_flag, uint32_t *col_mask) {
  // flag 1: every column is a full 3x3x3 stencil (27 rows);  flag 2: every column has the 2x2x2 arrangement of incident
  // cells (valence 8, lj = 7 - q) and its stored rows are a subset of the stencil (col_mask = present positions, the
  // in-column rank of a position is the number of present positions below it);  flag 0: anything else.
  int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  int nq = 0;
  bool stencil = true, full = true;
  for (int l = 0; l < 32; l++) {
    int64_t j = b * 32 + l;
    if (j >= ncols) { stencil = false; break; }
    int64_t kb = adj_ptr[j], ke = adj_ptr[j + 1];
    nq = max(nq, (int)(ke - kb));
    uint32_t mask = 0;
    bool ok = (ke - kb == 8);
    for (int q = 0; q < 8 && ok; q++) {
      if ((adj[kb + q] & 7) != 7 - q) { ok = false; break; }
      uint64_t rk = adj_rank[kb + q];
      for (int li = 0; li < 8; li++)
        if (((rk >> (8 * li)) & 0xFF) != 0xFF) {
          int r = 0, pw = 1;
          for (int d = 0; d < 3; d++) { r += pw * (((q >> d) & 1) + ((li >> d) & 1)); pw *= 3; }
          mask |= 1u << r;
        }
    }
    for (int q = 0; q < 8 && ok; q++) {
      uint64_t rk = adj_rank[kb + q];
      for (int li = 0; li < 8; li++) {
        unsigned actual = (unsigned)((rk >> (8 * li)) & 0xFF);
        if (actual == 0xFF) continue;
        int r = 0, pw = 1;
        for (int d = 0; d < 3; d++) { r += pw * (((q >> d) & 1) + ((li >> d) & 1)); pw *= 3; }
        if (actual != (unsigned)__popc(mask & ((1u << r) - 1))) ok = false;
      }
    }
    if (ok && colptr[j + 1] - colptr[j] != __popc(mask)) ok = false;
    col_mask[j] = ok ? mask : 0;
    stencil = stencil && ok;
    full = full && ok && mask == 0x7FFFFFFu;
  }
  blk_nq[b] = nq;
  blk_flag[b] = (stencil && full) ? 1 : stencil ? 2 : 0;
}

__global__ void blk_fill_kernel(const int64_t *adj_ptr, const int32_t *adj, const uint64_t *adj_rank, const int64_t *blk_ptr,
                                int64_t ncols, int64_t nblocks, int32_t *adjT_cell, uint64_t *adjT_rank, uint8_t *blk_flag,
                                int32_t *blk_base) {
  // one warp per block (blockDim is a multiple of 32)
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t b = t >> 5;
  int lane = (int)(t & 31);
  if (b >= nblocks) return;
  int64_t j = b * 32 + lane;
  int64_t row0 = blk_ptr[b];
  int nq = (int)(blk_ptr[b + 1] - row0);
  int64_t kb = 0, ke = 0;
  if (j < ncols) { kb = adj_ptr[j]; ke = adj_ptr[j + 1]; }
  bool runs = (nq == 8);
  int32_t base[8];
  for (int q = 0; q < nq; q++) {
    bool has = kb + q < ke;
    int32_t e = has ? adj[kb + q] : -1;
    adjT_cell[(row0 + q) * 32 + lane] = e;
    adjT_rank[(row0 + q) * 32 + lane] = has ? adj_rank[kb + q] : ~0ull;
    // run-length compression: lanes own consecutive cells at the same local position  <=>  e(lane) = e(0) + 8 lane
    int32_t e0 = __shfl_sync(0xffffffffu, e, 0);
    runs = runs && __all_sync(0xffffffffu, has && e == e0 + 8 * lane);
    if (q < 8) base[q] = e0;
  }
  if (lane == 0) {
    uint8_t f = blk_flag[b];
    if (runs && f != 0) {
      f |= 4;
      for (int q = 0; q < 8; q++) blk_base[b * 8 + q] = base[q];
      blk_flag[b] = f;
    }
  }
}

// ---- fast symbolic phase for one field with 8 local DoFs (scalar Q1 hexahedra) --------------------------------------------
// Same result as the generic path below (canonical CSC + cell-centric rank map), built column by column from the
// column -> incident (cell, lj) adjacency the gather plan needs anyway: a warp owns a column, loads the <= 64 candidate rows of
// its <= 8 incident cells straight from the cells' row ids, sorts them in registers (bitonic network over 2 keys per lane, key =
// row << 6 | origin), marks the unique ones and hands every candidate its in-column rank.  No 4.3 GB candidate array is written
// and re-read, no per-entry binary search: 256^3 takes ~10 ms instead of ~40 ms.  Falls back to the generic path when a column
// has more than 8 incident cells or the row ids do not fit 25 bits.
__device__ __forceinline__ uint32_t cmpx(uint32_t v, uint32_t p, bool keep_min) { return keep_min ? min(v, p) : max(v, p); }

__global__ void __launch_bounds__(128) q1_column_kernel(const int64_t *__restrict__ adj_ptr, int32_t *__restrict__ adj, const int32_t *__restrict__ row_ids,
                                                        int64_t ncols, int32_t *__restrict__ tmp_rows, int64_t *__restrict__ tmp_ptr,
                                                        int64_t *__restrict__ uniq, uint16_t *__restrict__ rank, uint64_t *__restrict__ adj_rank) {
  __shared__ unsigned char s_pos[4][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char *sp = s_pos[warp];
  const unsigned FULL = 0xffffffffu;
  for (int64_t j = blockIdx.x * 4ll + warp; j < ncols; j += (int64_t)gridDim.x * 4) {
    const int64_t b = adj_ptr[j];
    const int L = (int)(adj_ptr[j + 1] - b);
    if (lane == 0) tmp_ptr[j] = 8 * b;
    if (L == 0) {
      if (lane == 0) uniq[j] = 0;
      continue;
    }
    // incident (cell, lj) entries in ascending order (= ascending cells: the reference's summation order)
    const int32_t e_raw = lane < L ? adj[b + lane] : 0x7fffffff;
    int myrank = 0;
#pragma unroll
    for (int m = 0; m < 8; m++) myrank += (__shfl_sync(FULL, e_raw, m) < e_raw) ? 1 : 0;
    int32_t e_sorted = 0x7fffffff;
#pragma unroll
    for (int m = 0; m < 8; m++) {
      const int32_t v = __shfl_sync(FULL, e_raw, m);
      const int r = __shfl_sync(FULL, myrank, m);
      if (r == lane && m < L) e_sorted = v;
    }
    if (lane < L) adj[b + lane] = e_sorted;
    // candidates: slot 0 = (incident cell lane >> 3, local row lane & 7), slot 1 = (4 + (lane >> 3), lane & 7)
    const int li = lane & 7, q0 = lane >> 3, q1 = 4 + (lane >> 3);
    const int32_t e0 = __shfl_sync(FULL, e_sorted, q0), e1 = __shfl_sync(FULL, e_sorted, q1);
    int32_t row0 = 0, row1 = 0;
    if (q0 < L) row0 = row_ids[(int64_t)(e0 >> 3) * 8 + li];
    if (q1 < L) row1 = row_ids[(int64_t)(e1 >> 3) * 8 + li];
    const uint32_t INV = 0x7fffffffu;
    uint32_t v0 = row0 > 0 ? (((uint32_t)(row0 - 1) << 6) | (uint32_t)lane) : INV;
    uint32_t v1 = row1 > 0 ? (((uint32_t)(row1 - 1) << 6) | (uint32_t)(lane + 32)) : INV;
    // bitonic sort of the 64 keys, element i = lane + 32 * slot
#pragma unroll
    for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
      for (int jj = k >> 1; jj > 0; jj >>= 1) {
        if (jj == 32) {
          const uint32_t lo = min(v0, v1), hi = max(v0, v1);
          v0 = lo; v1 = hi;
        } else {
          const bool lower = (lane & jj) == 0;
          const bool up0 = (lane & k) == 0 || k == 64;                  // element lane
          const bool up1 = k == 64 ? true : (k == 32 ? false : (lane & k) == 0);  // element lane + 32
          const uint32_t p0 = __shfl_xor_sync(FULL, v0, jj), p1 = __shfl_xor_sync(FULL, v1, jj);
          v0 = cmpx(v0, p0, lower == up0);
          v1 = cmpx(v1, p1, lower == up1);
        }
      }
    }
    // unique rows and their in-column positions
    const uint32_t prev0 = __shfl_up_sync(FULL, v0, 1);
    const uint32_t last0 = __shfl_sync(FULL, v0, 31);
    uint32_t prev1 = __shfl_up_sync(FULL, v1, 1);
    if (lane == 0) prev1 = last0;
    const bool f0 = v0 != INV && (lane == 0 || (prev0 >> 6) != (v0 >> 6));
    const bool f1 = v1 != INV && (prev1 >> 6) != (v1 >> 6);
    const unsigned m0 = __ballot_sync(FULL, f0), m1 = __ballot_sync(FULL, f1);
    const unsigned le = 0xffffffffu >> (31 - lane);
    const int n0 = __popc(m0);
    const int pos0 = __popc(m0 & le) - 1, pos1 = n0 + __popc(m1 & le) - 1;
    const int64_t tb = 8 * b;
    if (f0) tmp_rows[tb + pos0] = (int32_t)(v0 >> 6);
    if (f1) tmp_rows[tb + pos1] = (int32_t)(v1 >> 6);
    if (lane == 0) uniq[j] = n0 + __popc(m1);
    if (v0 != INV) sp[v0 & 63u] = (unsigned char)pos0;
    if (v1 != INV) sp[v1 & 63u] = (unsigned char)pos1;
    __syncwarp();
    const unsigned r0 = row0 > 0 ? sp[lane] : 0xFFu, r1 = row1 > 0 ? sp[lane + 32] : 0xFFu;
    __syncwarp();
    if (q0 < L) rank[((int64_t)(e0 >> 3) * 8 + (e0 & 7)) * 8 + li] = r0 == 0xFFu ? (uint16_t)0xFFFF : (uint16_t)r0;
    if (q1 < L) rank[((int64_t)(e1 >> 3) * 8 + (e1 & 7)) * 8 + li] = r1 == 0xFFu ? (uint16_t)0xFFFF : (uint16_t)r1;
    // packed ranks of the gather plan: 8 x u8 per incident entry
    unsigned long long pk0 = (unsigned long long)r0 << (8 * li), pk1 = (unsigned long long)r1 << (8 * li);
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
      pk0 |= __shfl_xor_sync(FULL, pk0, d);
      pk1 |= __shfl_xor_sync(FULL, pk1, d);
    }
    if (li == 0 && q0 < L) adj_rank[b + q0] = pk0;
    if (li == 0 && q1 < L) adj_rank[b + q1] = pk1;
  }
}

int64_t exclusive_scan_i64(gb200_ctx ctx, const int64_t *in, int64_t *out, int64_t n) {
  // out[0..n] = exclusive prefix sums of in[0..n), out[n] = total; returns the total
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, n, ctx->stream);
  DevBuf<char> tmp;
  tmp.alloc(tmp_bytes);
  cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, in, out, n, ctx->stream);
  count_launch(ctx, 2);
  int64_t last_in = 0, last_out = 0;
  if (n > 0) {
    GB_CUDA(cudaMemcpyAsync(&last_in, in + n - 1, 8, cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(cudaMemcpyAsync(&last_out, out + n - 1, 8, cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  int64_t total = last_in + last_out;
  GB_CUDA(cudaMemcpyAsync(out + n, &total, 8, cudaMemcpyHostToDevice, ctx->stream));
  GB_CUDA(cudaStreamSynchronize(ctx->stream));
  return total;
}

int64_t max_i64(gb200_ctx ctx, const int64_t *in, int64_t n) {
  if (n == 0) return 0;
  DevBuf<int64_t> out;
  out.alloc(1);
  size_t tmp_bytes = 0;
  cub::DeviceReduce::Max(nullptr, tmp_bytes, in, out.p, n, ctx->stream);
  DevBuf<char> tmp;
  tmp.alloc(tmp_bytes);
  cub::DeviceReduce::Max(tmp.p, tmp_bytes, in, out.p, n, ctx->stream);
  count_launch(ctx, 2);
  int64_t r = 0;
  GB_CUDA(cudaMemcpyAsync(&r, out.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  GB_CUDA(cudaStreamSynchronize(ctx->stream));
  return r;
}

SymDesc make_symdesc(gb200_plan plan) {
  SymDesc d;
  memset(&d, 0, sizeof(d));
  d.nfields = plan->nfields;
  d.NL = plan->NL;
  d.ncells = plan->mesh->ncells;
  int ofs = 0;
  for (int f = 0; f < plan->nfields; f++) {
    d.nld[f] = plan->test[f]->nld;
    d.lofs[f] = ofs;
    ofs += d.nld[f];
    d.row_ids[f] = plan->test[f]->cell_dofs.p;
    d.col_ids[f] = plan->trial[f]->cell_dofs.p;
    d.row_off[f] = plan->row_off[f];
    d.col_off[f] = plan->col_off[f];
    for (int g = 0; g < plan->nfields; g++) d.touched[f][g] = plan->touched[f + plan->nfields * g];
  }
  return d;
}

inline int grid_for(int64_t n, int block, int num_sms) {
  int64_t g = (n + block - 1) / block;
  int64_t cap = (int64_t)num_sms * 32;
  return (int)std::max<int64_t>(1, std::min<int64_t>(g, cap));
}

}  // namespace

namespace {
__global__ void zero_based_kernel(int32_t *ids, int64_t n, int64_t nmax, unsigned long long *bad) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  unsigned long long local = 0;
  for (; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    int32_t v = ids[t];
    if (v < 1 || v > nmax) local++;
    ids[t] = v - 1;
  }
  if (local) atomicAdd(bad, local);
}
__global__ void range_check_kernel(const int32_t *ids, int64_t n, int64_t nfree, int64_t ndir, unsigned long long *bad) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  unsigned long long local = 0;
  for (; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    int32_t v = ids[t];
    if (v > nfree || -(int64_t)v > ndir) local++;
  }
  if (local) atomicAdd(bad, local);
}
}  // namespace

int64_t ids_to_zero_based(gb200_ctx ctx, int32_t *ids, int64_t n, int64_t nmax) {
  DevBuf<int64_t> bad;
  bad.alloc(1);
  bad.zero(ctx->stream);
  zero_based_kernel<<<grid_for(n, 256, ctx->num_sms), 256, 0, ctx->stream>>>(ids, n, nmax, (unsigned long long *)bad.p);
  check_launch(ctx, "zero_based_kernel");
  int64_t h = 0;
  bad.download(&h, ctx->stream);
  GB_CUDA(cudaStreamSynchronize(ctx->stream));
  return h;
}

int64_t count_ids_out_of_range(gb200_ctx ctx, const int32_t *ids, int64_t n, int64_t nfree, int64_t ndir) {
  DevBuf<int64_t> bad;
  bad.alloc(1);
  bad.zero(ctx->stream);
  range_check_kernel<<<grid_for(n, 256, ctx->num_sms), 256, 0, ctx->stream>>>(ids, n, nfree, ndir, (unsigned long long *)bad.p);
  check_launch(ctx, "range_check_kernel");
  int64_t h = 0;
  bad.download(&h, ctx->stream);
  GB_CUDA(cudaStreamSynchronize(ctx->stream));
  return h;
}

// Fast symbolic phase (see q1_column_kernel).  Leaves the sorted adjacency and its packed ranks in the plan for the gather plan.
static bool build_pattern_q1(gb200_plan plan) {
  gb200_ctx ctx = plan->ctx;
  cudaStream_t s = ctx->stream;
  static const bool disabled = getenv("GB200_NO_FAST_SYMBOLIC") != nullptr;
  if (disabled || plan->nfields != 1 || plan->NL != 8 || plan->nrows >= (1 << 25) || plan->row_off[0] != 0 || plan->col_off[0] != 0) return false;
  const int64_t ncols = plan->ncols, nc = plan->mesh->ncells, n = nc * 8;
  const int32_t *col_ids = plan->trial[0]->cell_dofs.p, *row_ids = plan->test[0]->cell_dofs.p;
  DevBuf<int64_t> cnt, cursor, tmp_ptr, uniq;
  cnt.alloc(ncols + 1);
  cnt.zero(s);
  adj_count_kernel<<<grid_for(n, 256, ctx->num_sms), 256, 0, s>>>(col_ids, n, (unsigned long long *)cnt.p);
  check_launch(ctx, "adj_count_kernel");
  if (max_i64(ctx, cnt.p, ncols) > 8) return false;  // a column with more than 8 incident cells: generic path
  plan->adj_ptr.alloc(ncols + 1);
  const int64_t total = exclusive_scan_i64(ctx, cnt.p, plan->adj_ptr.p, ncols);
  plan->adj_cell.alloc((size_t)std::max<int64_t>(total, 1));
  plan->adj_rank.alloc((size_t)std::max<int64_t>(total, 1));
  cursor.alloc(ncols + 1);
  cursor.zero(s);
  adj_fill_kernel<<<grid_for(n, 256, ctx->num_sms), 256, 0, s>>>(col_ids, n, plan->adj_ptr.p, (unsigned long long *)cursor.p, plan->adj_cell.p);
  check_launch(ctx, "adj_fill_kernel");
  DevBuf<int32_t> tmp_rows;
  tmp_rows.alloc((size_t)std::max<int64_t>(8 * total, 1));
  tmp_ptr.alloc(ncols + 1);
  uniq.alloc(ncols + 1);
  plan->rank.alloc((size_t)n * 8);
  GB_CUDA(cudaMemsetAsync(plan->rank.p, 0xFF, (size_t)n * 8 * sizeof(uint16_t), s));  // entries of masked columns stay 0xFFFF
  const int G = (int)std::max<int64_t>(1, std::min<int64_t>((ncols + 3) / 4, (int64_t)ctx->num_sms * 16));
  q1_column_kernel<<<G, 128, 0, s>>>(plan->adj_ptr.p, plan->adj_cell.p, row_ids, ncols, tmp_rows.p, tmp_ptr.p, uniq.p, plan->rank.p, plan->adj_rank.p);
  check_launch(ctx, "q1_column_kernel");
  plan->colptr.alloc(ncols + 1);
  plan->nnz = exclusive_scan_i64(ctx, uniq.p, plan->colptr.p, ncols);
  plan->rowval.alloc((size_t)std::max<int64_t>(plan->nnz, 1));
  compact_kernel<<<grid_for(ncols * 32, 256, ctx->num_sms), 256, 0, s>>>(tmp_ptr.p, tmp_rows.p, plan->colptr.p, plan->rowval.p, ncols);
  check_launch(ctx, "compact_kernel");
  GB_CUDA(cudaStreamSynchronize(s));
  plan->adj_ready = true;
  return true;
}

void build_pattern(gb200_plan plan) {
  gb200_ctx ctx = plan->ctx;
  cudaStream_t s = ctx->stream;
  ScopedTimer timer(ctx, "symbolic");
  if (build_pattern_q1(plan)) return;
  SymDesc d = make_symdesc(plan);
  const int64_t ncols = plan->ncols;
  const int64_t work = d.ncells * d.NL;
  const int B = 256;
  const int G = grid_for(work, B, ctx->num_sms);

  DevBuf<int64_t> bound, tmp_ptr, cursor, uniq;
  bound.alloc(ncols + 1);
  bound.zero(s);
  count_bound_kernel<<<G, B, 0, s>>>(d, (unsigned long long *)bound.p);
  check_launch(ctx, "count_bound_kernel");
  int64_t maxlen = max_i64(ctx, bound.p, ncols);
  GB_REQUIRE(maxlen <= 4096, GB200_ERR_UNSUPPORTED, "a column receives %lld candidate entries (limit 4096)", (long long)maxlen);
  tmp_ptr.alloc(ncols + 1);
  int64_t ncand = exclusive_scan_i64(ctx, bound.p, tmp_ptr.p, ncols);
  DevBuf<int32_t> tmp_rows;
  tmp_rows.alloc((size_t)std::max<int64_t>(ncand, 1));
  cursor.alloc(ncols + 1);
  cursor.zero(s);
  fill_candidates_kernel<<<G, B, 0, s>>>(d, tmp_ptr.p, (unsigned long long *)cursor.p, tmp_rows.p);
  check_launch(ctx, "fill_candidates_kernel");
  cursor.release();

  uniq.alloc(ncols + 1);
  int cap = 32;
  while (cap < maxlen) cap <<= 1;
  constexpr int WARPS = 4;
  size_t smem = (size_t)WARPS * cap * sizeof(int32_t);
  if (smem > 48 * 1024)
    GB_CUDA(cudaFuncSetAttribute(sort_unique_kernel<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int Gs = (int)std::max<int64_t>(1, std::min<int64_t>((ncols + WARPS - 1) / WARPS, (int64_t)ctx->num_sms * 16));
  sort_unique_kernel<WARPS><<<Gs, WARPS * 32, smem, s>>>(tmp_ptr.p, tmp_rows.p, uniq.p, ncols, cap);
  check_launch(ctx, "sort_unique_kernel");

  plan->colptr.alloc(ncols + 1);
  plan->nnz = exclusive_scan_i64(ctx, uniq.p, plan->colptr.p, ncols);
  GB_REQUIRE(max_i64(ctx, uniq.p, ncols) < 0xFFFF, GB200_ERR_UNSUPPORTED, "a column has more than 65534 stored entries");
  plan->rowval.alloc((size_t)std::max<int64_t>(plan->nnz, 1));
  compact_kernel<<<grid_for(ncols * 32, B, ctx->num_sms), B, 0, s>>>(tmp_ptr.p, tmp_rows.p, plan->colptr.p, plan->rowval.p, ncols);
  check_launch(ctx, "compact_kernel");
  GB_CUDA(cudaStreamSynchronize(s));
  tmp_rows.release();

  plan->rank.alloc((size_t)work * d.NL);
  rank_kernel<<<G, B, 0, s>>>(d, plan->colptr.p, plan->rowval.p, plan->rank.p);
  check_launch(ctx, "rank_kernel");
  GB_CUDA(cudaStreamSynchronize(s));
}

namespace {
struct WidenTask {
  const int32_t *src;
  int64_t *dst;
  int64_t n;
  int threads;
};
// executed by the CUDA runtime on the copy stream once the Int32 rows have arrived (no CUDA calls in here)
void CUDART_CB widen_rows_on_host(void *arg) {
  WidenTask *t = static_cast<WidenTask *>(arg);
  const int T = std::max(1, t->threads);
  std::vector<std::thread> pool;
  for (int k = 0; k < T; k++)
    pool.emplace_back([t, k, T] {
      const int64_t b = t->n * k / T, e = t->n * (k + 1) / T;
      const int32_t *src = t->src;
      int64_t *dst = t->dst;
      // 1-based Int64, as SparseMatrixCSC{Float64,Int} stores them; streaming stores: the destination is written once and not read
      // here (no read-for-ownership of 3.6 GB while the download of the values writes into the same memory system)
#if defined(__x86_64__)
      for (int64_t i = b; i < e; i++) _mm_stream_si64(reinterpret_cast<long long *>(dst + i), (long long)src[i] + 1);
      _mm_sfence();
#else
      for (int64_t i = b; i < e; i++) dst[i] = (int64_t)src[i] + 1;
#endif
    });
  for (auto &th : pool) th.join();
  delete t;
}
}  // namespace

void pattern_to_host(gb200_plan plan, int64_t *colptr, int64_t *rowval, bool async) {
  // Julia's SparseMatrixCSC{Float64,Int64}: 1-based Int64 colptr / rowval, copied on the context's copy stream; async = the caller's
  // next synchronising call on this context (or gb200_synchronize) completes the copy, so the transfer overlaps the numeric phase
  // that follows allocate_matrix inside assemble_matrix.  The row indices cross the link as the Int32 the device holds (half the
  // bytes of the Int64 the caller gets: 1.8 instead of 3.6 GB at 256^3) into a page-locked staging buffer and are widened by host
  // threads while the download of the values (which waits for the rows: the two copies do not share the link) is in flight.
  // GB200_HOST_WIDEN=0: widen on the device and copy Int64 (the round-1 path); small patterns always take that path.
  gb200_ctx ctx = plan->ctx;
  sync_copies(ctx);
  static const bool host_widen_off = getenv("GB200_HOST_WIDEN") && getenv("GB200_HOST_WIDEN")[0] == '0';
  // threads of the host-side widening: GB200_HOST_THREADS, else the host's cores shared among the ranks of the node (LOCAL_WORLD_SIZE,
  // as torchrun / most MPI launchers export it), at most 16
  static const int host_threads = [] {
    if (getenv("GB200_HOST_THREADS")) return std::max(1, atoi(getenv("GB200_HOST_THREADS")));
    const int local_world = getenv("LOCAL_WORLD_SIZE") ? std::max(1, atoi(getenv("LOCAL_WORLD_SIZE"))) : 1;
    return (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency() / (unsigned)local_world));
  }();
  // (with many ranks per node the host's memory system is the bottleneck of the downloads already: 8 ranks measured 93 ms with the
  // host-side widening against 89 ms without, so it is used with up to two ranks per node)
  static const int local_world = getenv("LOCAL_WORLD_SIZE") ? std::max(1, atoi(getenv("LOCAL_WORLD_SIZE"))) : 1;
  static const bool host_widen_forced = getenv("GB200_HOST_WIDEN") && getenv("GB200_HOST_WIDEN")[0] == '1';
  const bool host_widen = !host_widen_off && plan->nnz >= (1 << 22) && (local_world <= 2 || host_widen_forced);
  int64_t *c1 = static_cast<int64_t *>(dev_alloc((size_t)(plan->ncols + 1) * 8));
  ctx->copy_keep.push_back(c1);
  int64_t *r1 = nullptr;
  if (!host_widen) {
    r1 = static_cast<int64_t *>(dev_alloc((size_t)std::max<int64_t>(plan->nnz, 1) * 8));
    ctx->copy_keep.push_back(r1);
  } else if (ctx->host_stage_bytes < (size_t)plan->nnz * 4) {
    if (ctx->host_stage) cudaFreeHost(ctx->host_stage);
    ctx->host_stage = nullptr;
    ctx->host_stage_bytes = 0;
    GB_CUDA(cudaHostAlloc(&ctx->host_stage, (size_t)plan->nnz * 4, cudaHostAllocDefault));
    ctx->host_stage_bytes = (size_t)plan->nnz * 4;
  }
  to_one_based_kernel<<<grid_for(std::max(host_widen ? (int64_t)0 : plan->nnz, plan->ncols + 1), 256, ctx->num_sms), 256, 0, ctx->stream>>>(
      plan->colptr.p, plan->rowval.p, c1, r1, plan->ncols, host_widen ? 0 : plan->nnz);
  check_launch(ctx, "to_one_based_kernel");
  cudaEvent_t ready;
  GB_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  GB_CUDA(cudaEventRecord(ready, ctx->stream));
  GB_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ready, 0));
  GB_CUDA(cudaEventDestroy(ready));
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0);
  cudaEventCreate(&t1);
  cudaEventRecord(t0, ctx->copy_stream);
  GB_CUDA(cudaMemcpyAsync(colptr, c1, (size_t)(plan->ncols + 1) * 8, cudaMemcpyDeviceToHost, ctx->copy_stream));
  if (plan->nnz && !host_widen) GB_CUDA(cudaMemcpyAsync(rowval, r1, (size_t)plan->nnz * 8, cudaMemcpyDeviceToHost, ctx->copy_stream));
  if (plan->nnz && host_widen) {
    GB_CUDA(cudaMemcpyAsync(ctx->host_stage, plan->rowval.p, (size_t)plan->nnz * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
    if (!ctx->pattern_copied) GB_CUDA(cudaEventCreateWithFlags(&ctx->pattern_copied, cudaEventDisableTiming));
    GB_CUDA(cudaEventRecord(ctx->pattern_copied, ctx->copy_stream));
    ctx->pattern_copied_pending = true;
    WidenTask *task = new WidenTask{static_cast<const int32_t *>(ctx->host_stage), rowval, plan->nnz, host_threads};
    GB_CUDA(cudaLaunchHostFunc(ctx->copy_stream, widen_rows_on_host, task));
  }
  cudaEventRecord(t1, ctx->copy_stream);
  ctx->pending.push_back({"pattern_d2h", t0, t1});
  ctx->copy_pending = true;
  if (!async) sync_copies(ctx);
}


// ---- BlockMultiFieldStyle output (src/MultiField/BlockSparseMatrixAssemblers.jl:19-33,197-230): the matrix of field block
// (bi, bj) as its own CSC with block-local ids.  Rows are sorted inside a column of the consecutive matrix and field k's rows
// are the contiguous id range [row_off[k], row_off[k+1]): block (bi, bj) of a column is one contiguous piece of it.
namespace {
__global__ void block_ranges_kernel(const int64_t *colptr, const int32_t *rowval, int64_t col0, int64_t ncols_b, int32_t row_lo,
                                    int32_t row_hi, int64_t *beg, int64_t *len) {
  int64_t jl = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (jl >= ncols_b) return;
  const int64_t b = colptr[col0 + jl], e = colptr[col0 + jl + 1];
  auto lower = [&](int32_t v) {
    int64_t lo = b, hi = e;
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if (rowval[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
  };
  const int64_t p0 = lower(row_lo), p1 = lower(row_hi);
  beg[jl] = p0;
  len[jl] = p1 - p0;
}
__global__ void block_gather_kernel(const int64_t *beg, const int64_t *bptr, int64_t ncols_b, const int32_t *rowval, int32_t row_lo,
                                    const double *nzval, int64_t *colptr1, int64_t *rowval1, double *nzval_b) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int64_t j = warp; j < ncols_b; j += nwarps) {
    const int64_t src = beg[j], dst = bptr[j], L = bptr[j + 1] - dst;
    if (colptr1 && lane == 0) {
      colptr1[j] = dst + 1;
      if (j == ncols_b - 1) colptr1[ncols_b] = bptr[ncols_b] + 1;
    }
    for (int64_t q = lane; q < L; q += 32) {
      if (rowval1) rowval1[dst + q] = (int64_t)(rowval[src + q] - row_lo) + 1;
      if (nzval_b) nzval_b[dst + q] = nzval[src + q];
    }
  }
}
}  // namespace

// Fills plan->blk_beg / blk_bptr for block (bi, bj) (cached) and returns its nnz.
int64_t block_layout(gb200_plan plan, int bi, int bj) {
  gb200_ctx ctx = plan->ctx;
  const int nf = plan->nfields, id = bi + nf * bj;
  if (plan->block_beg.size() != (size_t)nf * nf) {
    plan->block_beg.resize((size_t)nf * nf);
    plan->block_ptr.resize((size_t)nf * nf);
    plan->block_nnz.assign((size_t)nf * nf, -1);
  }
  if (plan->block_nnz[id] >= 0) return plan->block_nnz[id];
  const int64_t col0 = plan->col_off[bj], col1 = bj + 1 < nf ? plan->col_off[bj + 1] : plan->ncols;
  const int64_t row0 = plan->row_off[bi], row1 = bi + 1 < nf ? plan->row_off[bi + 1] : plan->nrows;
  const int64_t ncb = col1 - col0;
  DevBuf<int64_t> len;
  plan->block_beg[id].alloc((size_t)ncb + 1);
  plan->block_ptr[id].alloc((size_t)ncb + 1);
  len.alloc((size_t)ncb + 1);
  block_ranges_kernel<<<(int)((ncb + 255) / 256), 256, 0, ctx->stream>>>(plan->colptr.p, plan->rowval.p, col0, ncb, (int32_t)row0, (int32_t)row1,
                                                                        plan->block_beg[id].p, len.p);
  check_launch(ctx, "block_ranges_kernel");
  plan->block_nnz[id] = exclusive_scan_i64(ctx, len.p, plan->block_ptr[id].p, ncb);
  return plan->block_nnz[id];
}

// colptr1 / rowval1 (1-based, block-local, Int64) and / or the values of block (bi, bj) to host arrays (any may be null).
void block_to_host(gb200_plan plan, int bi, int bj, int64_t *colptr, int64_t *rowval, double *nzval) {
  gb200_ctx ctx = plan->ctx;
  const int nf = plan->nfields, id = bi + nf * bj;
  const int64_t nnzb = block_layout(plan, bi, bj);
  const int64_t col0 = plan->col_off[bj], col1 = bj + 1 < nf ? plan->col_off[bj + 1] : plan->ncols;
  const int64_t ncb = col1 - col0;
  DevBuf<int64_t> c1, r1;
  DevBuf<double> v1;
  if (colptr) c1.alloc((size_t)ncb + 1);
  if (rowval) r1.alloc((size_t)std::max<int64_t>(nnzb, 1));
  if (nzval) v1.alloc((size_t)std::max<int64_t>(nnzb, 1));
  block_gather_kernel<<<grid_for(ncb * 32, 256, ctx->num_sms), 256, 0, ctx->stream>>>(plan->block_beg[id].p, plan->block_ptr[id].p, ncb, plan->rowval.p,
                                                                                     (int32_t)plan->row_off[bi], plan->nzval.p, c1.p, r1.p, v1.p);
  check_launch(ctx, "block_gather_kernel");
  if (colptr) GB_CUDA(cudaMemcpyAsync(colptr, c1.p, (size_t)(ncb + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (rowval && nnzb) GB_CUDA(cudaMemcpyAsync(rowval, r1.p, (size_t)nnzb * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (nzval && nnzb) GB_CUDA(cudaMemcpyAsync(nzval, v1.p, (size_t)nnzb * 8, cudaMemcpyDeviceToHost, ctx->stream));
  GB_CUDA(cudaStreamSynchronize(ctx->stream));
}

// ---- forms over several triangulations: add every stored value of `src` at the slot of the same (row, column) in `dst`
namespace {
__global__ void add_from_kernel(const int64_t *scolptr, const int32_t *srowval, const double *snz, int64_t ncols, const int64_t *dcolptr,
                                const int32_t *drowval, double *dnz, unsigned long long *missing) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int64_t j = warp; j < ncols; j += nwarps) {
    const int64_t db = dcolptr[j], de = dcolptr[j + 1];
    for (int64_t s = scolptr[j] + lane; s < scolptr[j + 1]; s += 32) {
      const int32_t row = srowval[s];
      int64_t lo = db, hi = de;
      while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (drowval[mid] < row) lo = mid + 1; else hi = mid;
      }
      if (lo < de && drowval[lo] == row) dnz[lo] += snz[s];  // one writer per slot: src entries are unique
      else atomicAdd(missing, 1ull);
    }
  }
}
}  // namespace

void add_matrix_from(gb200_plan dst, gb200_plan src) {
  gb200_ctx ctx = dst->ctx;
  if (!src->nnz) return;
  DevBuf<int64_t> missing;
  missing.alloc(1);
  missing.zero(ctx->stream);
  add_from_kernel<<<grid_for(src->ncols * 32, 256, ctx->num_sms), 256, 0, ctx->stream>>>(src->colptr.p, src->rowval.p, src->nzval.p, src->ncols,
                                                                                        dst->colptr.p, dst->rowval.p, dst->nzval.p,
                                                                                        (unsigned long long *)missing.p);
  check_launch(ctx, "add_from_kernel");
  int64_t h = 0;
  missing.download(&h, ctx->stream);
  GB_CUDA(cudaStreamSynchronize(ctx->stream));
  GB_REQUIRE(h == 0, GB200_ERR_INVALID, "%lld entries of the added triangulation are not in the pattern of the target matrix", (long long)h);
}

// ---- linear constraints (FESpaceWithLinearConstraints): dst = T^T src T on the free master DoFs, the Dirichlet master columns moved
// to the vector.  The reference applies the cell-wise constraint matrices to every cell matrix / vector before the scatter
// (ConstrainRowsMap / ConstrainColsMap, src/FESpaces/FESpaceInterface.jl:361-387 and src/FESpaces/ConstantFESpaces... attach_constraints_*);
// summed over the cells that is the triple product with the global constraint table, applied here to the assembled arrays of the
// unconstrained space (free and Dirichlet DoFs as one positive numbering): the fast cell kernels run unchanged.
namespace {
__global__ void fold_matrix_kernel(const int64_t *scolptr, const int32_t *srowval, const double *snz, int64_t ncols, const int64_t *ptrs,
                                   const int32_t *mdofs, const double *coeffs, const double *dir_vals, const int64_t *dcolptr,
                                   const int32_t *drowval, double *dnz, double *dvec, unsigned long long *missing) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int64_t J = warp; J < ncols; J += nwarps) {
    const int64_t qb = ptrs[J], qe = ptrs[J + 1];
    for (int64_t s = scolptr[J] + lane; s < scolptr[J + 1]; s += 32) {
      const int64_t I = srowval[s];
      const double a = snz[s];
      for (int64_t q = qb; q < qe; q++) {          // masters of the column DoF
        const int32_t n = mdofs[q];
        const double cn = coeffs[q];
        for (int64_t r = ptrs[I]; r < ptrs[I + 1]; r++) {   // masters of the row DoF
          const int32_t m = mdofs[r];
          if (m <= 0) continue;                     // rows of Dirichlet masters are not assembled
          const double v = coeffs[r] * cn * a;
          if (n > 0) {
            if (!dnz) continue;
            const int64_t db = dcolptr[n - 1], de = dcolptr[n];
            int64_t lo = db, hi = de;
            while (lo < hi) {
              int64_t mid = (lo + hi) >> 1;
              if (drowval[mid] < m - 1) lo = mid + 1; else hi = mid;
            }
            if (lo < de && drowval[lo] == m - 1) atomicAdd(dnz + lo, v);
            else atomicAdd(missing, 1ull);
          } else if (n < 0 && dvec && dir_vals) {   // Dirichlet master column: lifting
            atomicAdd(dvec + (m - 1), -v * dir_vals[-n - 1]);
          }
        }
      }
    }
  }
}
__global__ void fold_vector_kernel(const double *svec, int64_t n, const int64_t *ptrs, const int32_t *mdofs, const double *coeffs, double *dvec) {
  for (int64_t I = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; I < n; I += (int64_t)gridDim.x * blockDim.x) {
    const double b = svec[I];
    for (int64_t r = ptrs[I]; r < ptrs[I + 1]; r++)
      if (mdofs[r] > 0) atomicAdd(dvec + (mdofs[r] - 1), coeffs[r] * b);
  }
}
}  // namespace

void fold_constraints(gb200_plan dst, gb200_plan src, const int64_t *h_ptrs, const int32_t *h_mdofs, const double *h_coeffs, const double *h_dir,
                      int64_t ndir, bool with_matrix, bool with_vector) {
  gb200_ctx ctx = dst->ctx;
  cudaStream_t s = ctx->stream;
  const int64_t n = src->ncols;
  std::vector<int64_t> ptrs0((size_t)n + 1);
  for (int64_t i = 0; i <= n; i++) ptrs0[(size_t)i] = h_ptrs[i] - 1;
  const int64_t nd = ptrs0[(size_t)n];
  DevBuf<int64_t> ptrs, missing;
  DevBuf<int32_t> mdofs;
  DevBuf<double> coeffs, dir;
  ptrs.upload(ptrs0.data(), ptrs0.size(), s);
  mdofs.upload(h_mdofs, (size_t)nd, s);
  coeffs.upload(h_coeffs, (size_t)nd, s);
  if (h_dir && ndir > 0) dir.upload(h_dir, (size_t)ndir, s);
  missing.alloc(1);
  missing.zero(s);
  if (with_matrix && dst->nnz) dst->nzval.zero(s);
  if (with_vector) dst->bvec.zero(s);
  count_launch(ctx, (with_matrix ? 1 : 0) + (with_vector ? 1 : 0));
  if ((with_matrix || (with_vector && dir.p)) && src->nnz) {
    fold_matrix_kernel<<<grid_for(n * 32, 256, ctx->num_sms), 256, 0, s>>>(src->colptr.p, src->rowval.p, src->nzval.p, n, ptrs.p, mdofs.p, coeffs.p, dir.p,
                                                                           dst->colptr.p, dst->rowval.p, with_matrix ? dst->nzval.p : nullptr,
                                                                           with_vector ? dst->bvec.p : nullptr, (unsigned long long *)missing.p);
    check_launch(ctx, "fold_matrix_kernel");
  }
  if (with_vector) {
    fold_vector_kernel<<<grid_for(n, 256, ctx->num_sms), 256, 0, s>>>(src->bvec.p, n, ptrs.p, mdofs.p, coeffs.p, dst->bvec.p);
    check_launch(ctx, "fold_vector_kernel");
  }
  int64_t h = 0;
  missing.download(&h, s);
  GB_CUDA(cudaStreamSynchronize(s));
  GB_REQUIRE(h == 0, GB200_ERR_INVALID, "%lld folded entries are not in the pattern of the constrained matrix", (long long)h);
}

// ---- SparseMatrixCSR output (src/Algebra/SparseMatrixCSR.jl:31-75: the reference assembles the CSC of the transpose and
// transposes it): rowptr / colval with columns ascending inside a row, and for every CSR position the CSC slot it comes from.
namespace {
__global__ void csr_count_kernel(const int32_t *rowval, int64_t nnz, unsigned long long *cnt) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (; t < nnz; t += (int64_t)gridDim.x * blockDim.x) atomicAdd(&cnt[rowval[t]], 1ull);
}
__global__ void csr_fill_kernel(const int64_t *colptr, const int32_t *rowval, int64_t ncols, const int64_t *rowptr, unsigned long long *cursor,
                                int32_t *col_tmp, int64_t *src_tmp) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int64_t j = warp; j < ncols; j += nwarps)
    for (int64_t s = colptr[j] + lane; s < colptr[j + 1]; s += 32) {
      const int32_t i = rowval[s];
      const int64_t pos = rowptr[i] + (int64_t)atomicAdd(&cursor[i], 1ull);
      col_tmp[pos] = (int32_t)j;
      src_tmp[pos] = s;
    }
}
// one warp per row: bitonic sort of (column << 32 | position in the unsorted segment) in shared memory
template <int WARPS>
__global__ void csr_sort_kernel(const int64_t *rowptr, const int32_t *col_tmp, const int64_t *src_tmp, int64_t nrows, int cap, int32_t *colval,
                                int64_t *src) {
  extern __shared__ unsigned long long smk[];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long *buf = smk + (size_t)warp * cap;
  for (int64_t i = blockIdx.x * (int64_t)WARPS + warp; i < nrows; i += (int64_t)gridDim.x * WARPS) {
    const int64_t beg = rowptr[i];
    const int L = (int)(rowptr[i + 1] - beg);
    if (L == 0) continue;
    int P = 32;
    while (P < L) P <<= 1;
    for (int q = lane; q < P; q += 32) buf[q] = q < L ? (((unsigned long long)(uint32_t)col_tmp[beg + q] << 32) | (unsigned)q) : ~0ull;
    __syncwarp();
    for (int size = 2; size <= P; size <<= 1)
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int q = lane; q < (P >> 1); q += 32) {
          int lo = 2 * q - (q & (stride - 1));
          int hi = lo + stride;
          bool up = ((lo & size) == 0);
          unsigned long long a = buf[lo], b = buf[hi];
          if ((a > b) == up) { buf[lo] = b; buf[hi] = a; }
        }
        __syncwarp();
      }
    for (int q = lane; q < L; q += 32) {
      const unsigned long long v = buf[q];
      colval[beg + q] = (int32_t)(v >> 32);
      src[beg + q] = src_tmp[beg + (int64_t)(v & 0xffffffffull)];
    }
    __syncwarp();
  }
}
__global__ void csr_gather_kernel(const int64_t *rowptr, const int32_t *colval, const int64_t *src, const double *nzval, int64_t nrows, int64_t nnz,
                                  int64_t base, int64_t *rowptr1, int64_t *colval1, double *nz_csr) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (int64_t q = t; q < nnz; q += (int64_t)gridDim.x * blockDim.x) {
    if (colval1) colval1[q] = (int64_t)colval[q] + base;
    if (nz_csr) nz_csr[q] = nzval[src[q]];
  }
  if (rowptr1)
    for (int64_t q = t; q <= nrows; q += (int64_t)gridDim.x * blockDim.x) rowptr1[q] = rowptr[q] + base;
}
}  // namespace

void ensure_csr(gb200_plan plan) {
  if (plan->csr_rowptr.n) return;
  gb200_ctx ctx = plan->ctx;
  cudaStream_t s = ctx->stream;
  ScopedTimer timer(ctx, "csr_plan");
  const int64_t nrows = plan->nrows, nnz = plan->nnz;
  DevBuf<int64_t> cnt, cursor, src_tmp;
  DevBuf<int32_t> col_tmp;
  cnt.alloc((size_t)nrows + 1);
  cnt.zero(s);
  if (nnz) {
    csr_count_kernel<<<grid_for(nnz, 256, ctx->num_sms), 256, 0, s>>>(plan->rowval.p, nnz, (unsigned long long *)cnt.p);
    check_launch(ctx, "csr_count_kernel");
  }
  const int64_t maxlen = max_i64(ctx, cnt.p, nrows);
  GB_REQUIRE(maxlen <= 4096, GB200_ERR_UNSUPPORTED, "a row has %lld stored entries (CSR limit 4096)", (long long)maxlen);
  plan->csr_rowptr.alloc((size_t)nrows + 1);
  exclusive_scan_i64(ctx, cnt.p, plan->csr_rowptr.p, nrows);
  plan->csr_colval.alloc((size_t)std::max<int64_t>(nnz, 1));
  plan->csr_src.alloc((size_t)std::max<int64_t>(nnz, 1));
  if (!nnz) return;
  col_tmp.alloc((size_t)nnz);
  src_tmp.alloc((size_t)nnz);
  cursor.alloc((size_t)nrows + 1);
  cursor.zero(s);
  csr_fill_kernel<<<grid_for(plan->ncols * 32, 256, ctx->num_sms), 256, 0, s>>>(plan->colptr.p, plan->rowval.p, plan->ncols, plan->csr_rowptr.p,
                                                                               (unsigned long long *)cursor.p, col_tmp.p, src_tmp.p);
  check_launch(ctx, "csr_fill_kernel");
  int cap = 32;
  while (cap < maxlen) cap <<= 1;
  constexpr int WARPS = 4;
  const size_t smem = (size_t)WARPS * cap * sizeof(unsigned long long);
  if (smem > 48 * 1024) GB_CUDA(cudaFuncSetAttribute(csr_sort_kernel<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int G = (int)std::max<int64_t>(1, std::min<int64_t>((nrows + WARPS - 1) / WARPS, (int64_t)ctx->num_sms * 16));
  csr_sort_kernel<WARPS><<<G, WARPS * 32, smem, s>>>(plan->csr_rowptr.p, col_tmp.p, src_tmp.p, nrows, cap, plan->csr_colval.p, plan->csr_src.p);
  check_launch(ctx, "csr_sort_kernel");
  GB_CUDA(cudaStreamSynchronize(s));
}

// rowptr / colval (Int64, index base `base` = the Bi of SparseMatrixCSR{Bi}) and / or the values in CSR order, to host arrays
void csr_to_host(gb200_plan plan, int64_t base, int64_t *rowptr, int64_t *colval, double *nzval) {
  gb200_ctx ctx = plan->ctx;
  ensure_csr(plan);
  const int64_t nrows = plan->nrows, nnz = plan->nnz;
  DevBuf<int64_t> r1, c1;
  DevBuf<double> v1;
  if (rowptr) r1.alloc((size_t)nrows + 1);
  if (colval) c1.alloc((size_t)std::max<int64_t>(nnz, 1));
  if (nzval) v1.alloc((size_t)std::max<int64_t>(nnz, 1));
  csr_gather_kernel<<<grid_for(std::max(nnz, nrows + 1), 256, ctx->num_sms), 256, 0, ctx->stream>>>(
      plan->csr_rowptr.p, plan->csr_colval.p, plan->csr_src.p, plan->nzval.p, nrows, nnz, base, r1.p, c1.p, v1.p);
  check_launch(ctx, "csr_gather_kernel");
  if (rowptr) GB_CUDA(cudaMemcpyAsync(rowptr, r1.p, (size_t)(nrows + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (colval && nnz) GB_CUDA(cudaMemcpyAsync(colval, c1.p, (size_t)nnz * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (nzval && nnz) GB_CUDA(cudaMemcpyAsync(nzval, v1.p, (size_t)nnz * 8, cudaMemcpyDeviceToHost, ctx->stream));
  GB_CUDA(cudaStreamSynchronize(ctx->stream));
}

void ensure_gather_plan(gb200_plan plan) {
  if (!plan->gather_plan_pending) return;
  plan->gather_plan_pending = false;
  build_gather_plan(plan);
}

void build_gather_plan(gb200_plan plan) {
  gb200_ctx ctx = plan->ctx;
  cudaStream_t s = ctx->stream;
  ScopedTimer timer(ctx, "gather_plan");
  const int nld = plan->NL;
  const int64_t n = plan->mesh->ncells * nld;
  const int64_t ncols = plan->ncols;
  const int32_t *col_ids = plan->trial[0]->cell_dofs.p;
  if (!plan->adj_ready) {  // (the fast symbolic phase leaves the sorted adjacency and its packed ranks behind)
    DevBuf<int64_t> cnt, cursor;
    cnt.alloc(ncols + 1);
    cnt.zero(s);
    adj_count_kernel<<<grid_for(n, 256, ctx->num_sms), 256, 0, s>>>(col_ids, n, (unsigned long long *)cnt.p);
    check_launch(ctx, "adj_count_kernel");
    plan->adj_ptr.alloc(ncols + 1);
    int64_t total = exclusive_scan_i64(ctx, cnt.p, plan->adj_ptr.p, ncols);
    plan->adj_cell.alloc((size_t)std::max<int64_t>(total, 1));
    plan->adj_rank.alloc((size_t)std::max<int64_t>(total, 1));
    cursor.alloc(ncols + 1);
    cursor.zero(s);
    adj_fill_kernel<<<grid_for(n, 256, ctx->num_sms), 256, 0, s>>>(col_ids, n, plan->adj_ptr.p, (unsigned long long *)cursor.p,
                                                                  plan->adj_cell.p);
    check_launch(ctx, "adj_fill_kernel");
    adj_sort_pack_kernel<<<grid_for(ncols, 128, ctx->num_sms), 128, 0, s>>>(plan->adj_ptr.p, plan->adj_cell.p, plan->rank.p, nld,
                                                                           plan->adj_rank.p, ncols);
    check_launch(ctx, "adj_sort_pack_kernel");
  }
  plan->adj_ready = false;
  // blocked-transposed layout
  const int64_t nblocks = (ncols + 31) / 32;
  DevBuf<int64_t> blk_nq;
  blk_nq.alloc(nblocks + 1);
  plan->blk_flag.alloc(nblocks);
  plan->col_mask.alloc((size_t)ncols);
  plan->blk_base.alloc((size_t)nblocks * 8);
  blk_count_kernel<<<(int)((nblocks + 127) / 128), 128, 0, s>>>(plan->adj_ptr.p, plan->adj_cell.p, plan->adj_rank.p, plan->colptr.p,
                                                               ncols, nblocks, blk_nq.p, plan->blk_flag.p, plan->col_mask.p);
  check_launch(ctx, "blk_count_kernel");
  plan->blk_ptr.alloc(nblocks + 1);
  int64_t nrowsT = exclusive_scan_i64(ctx, blk_nq.p, plan->blk_ptr.p, nblocks);
  plan->adjT_cell.alloc((size_t)std::max<int64_t>(nrowsT * 32, 1));
  plan->adjT_rank.alloc((size_t)std::max<int64_t>(nrowsT * 32, 1));
  blk_fill_kernel<<<(int)((nblocks * 32 + 255) / 256), 256, 0, s>>>(plan->adj_ptr.p, plan->adj_cell.p, plan->adj_rank.p, plan->blk_ptr.p,
                                                                   ncols, nblocks, plan->adjT_cell.p, plan->adjT_rank.p, plan->blk_flag.p,
                                                                   plan->blk_base.p);
  check_launch(ctx, "blk_fill_kernel");
  GB_CUDA(cudaStreamSynchronize(s));
  plan->adj_cell.release();
  plan->adj_rank.release();
  plan->adj_ptr.release();
  DevBuf<int64_t> spanmax;
  spanmax.alloc(1);
  spanmax.zero(s);
  const int cols_per_cta = 32;  // one warp of the gather kernel
  int64_t nctas = (ncols + cols_per_cta - 1) / cols_per_cta;
  span_max_kernel<<<(int)((nctas + 255) / 256), 256, 0, s>>>(plan->colptr.p, ncols, cols_per_cta, (unsigned long long *)spanmax.p);
  check_launch(ctx, "span_max_kernel");
  spanmax.download(&plan->gather_span_max, s);
  GB_CUDA(cudaStreamSynchronize(s));
  plan->has_gather = plan->gather_span_max * 8 * 4 <= 200 * 1024;
}

}  // namespace gb
