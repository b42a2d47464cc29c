// q1hex_common.cuh -- closed-form local-matrix entries of scalar Q1 hexahedra on affine cells, shared by the gather kernels.
#pragma once
#include "common.cuh"

namespace gb {
namespace q1 {

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
// FORM == Q1_STAGED: the local matrices were computed by a cell-parallel kernel for general (non-affine) geometry and staged
// as their 36 unique entries, SoA [36][ncells] (symmetric forms); the gather then only reads them.
constexpr int Q1_STAGED = 100;
__host__ __device__ constexpr int sym_index(int i, int j) {  // upper-triangular packed index of (min, max), 8x8
  return (i < j ? i : j) * 8 - ((i < j ? i : j) * ((i < j ? i : j) - 1)) / 2 + ((i < j ? j : i) - (i < j ? i : j));
}

// G: per-cell factors, SoA with stride `ncells` between the factor arrays.  DIAG: the metric of every cell of the mesh is diagonal
// (the off-diagonal factors are exactly 0.0 and were not stored): the same arithmetic with those terms folded away.
template <int FORM, bool DIAG = false>
__device__ __forceinline__ void column_entries(const double *__restrict__ G, int64_t ncells, int64_t cell, int lj, double coef, double *vals) {
  if (FORM == Q1_STAGED) {
#pragma unroll
    for (int m = 0; m < 8; m++) vals[m] = coef * __ldg(G + (int64_t)sym_index(m ^ lj, lj) * ncells + cell);
  } else if (FORM == GB200_FORM_LAPLACIAN) {
    const double t0 = (lj & 1) ? 1.0 : -1.0, t1 = (lj & 2) ? 1.0 : -1.0, t2 = (lj & 4) ? 1.0 : -1.0;
    const double d0 = coef * __ldg(G + cell), d1 = coef * __ldg(G + ncells + cell), d2 = coef * __ldg(G + 2 * ncells + cell);
    const double o01 = DIAG ? 0.0 : 0.25 * coef * t0 * t1 * __ldg(G + 3 * ncells + cell), o02 = DIAG ? 0.0 : 0.25 * coef * t0 * t2 * __ldg(G + 4 * ncells + cell),
                 o12 = DIAG ? 0.0 : 0.25 * coef * t1 * t2 * __ldg(G + 5 * ncells + cell);
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


}  // namespace q1
}  // namespace gb
