// vector_kernels.cu -- cell-centric kernels specialised for ONE vector-valued Lagrangian field in 3D (3 components):
// linear elasticity, vector Laplacian / mass and the neo-Hookean Jacobian (BASELINE.json configs 3 and 5).
//
// Reference work being replaced per cell: a4-a8 of SURVEY.md section 8 (Jacobian, inverse, physical gradients, the integrand
// broadcast aq[p,i,j] of src/Fields/FieldArrays.jl:675-696 and the IntegrationMap contraction src/Fields/FieldsInterfaces.jl:737-760).
//
// The generic kernel evaluates one (li,lj) entry per thread with run-time sizes and is instruction-bound (ncu: issue 63 %,
// FP64 pipe 9 % on Q2 elasticity).  Here element sizes are template parameters and a thread owns a NODE PAIR (a,b): it
// accumulates the 3x3 component block K[(a,ci),(b,cj)] over the quadrature points from quantities that depend only on
// (a,p), (b,p) or p -- gradients are loaded once per 9 entries and the constitutive algebra is hoisted out of the entry loop:
//   elasticity:   K += dV [ lambda ga_ci gb_cj + mu ( delta_cicj ga.gb + ga_cj gb_ci ) ]
//   neo-Hookean:  K += dV [ lambda beta_cj alpha_ci + kappa ( c_ab Z_cj,ci + alpha_cj beta_ci ) + delta_cicj s_ab ]
//                 alpha_c = ga.y_c, beta_c = gb.y_c, y_c = C^-1 F[c,:], Z = F C^-1 F^T, c_ab = ga.C^-1.gb, s_ab = ga.S.gb,
//                 kappa = mu - lambda ln J      (dE(grad v):dS(grad du) + grad v:(S.grad du))
// Scatter: slot = colptr[col] + rank (plan), RED.ADD.F64 or plain RMW per colour in deterministic mode.
#include "common.cuh"

namespace gb {

namespace {

struct VArgs {
  const double *X;
  const int32_t *cell_nodes;
  const double *w, *dNg, *N, *dN;  // tabulation (device): w[NP], dNg[NP][NN][3], N[NP][NDS], dN[NP][NDS][3]
  const int32_t *row_ids, *col_ids, *state_ids;
  const double *free_vals, *dir_vals;
  int64_t row_off, col_off;
  const int64_t *colptr;
  const uint16_t *rank;
  double *nzval;
  const int32_t *cell_list;
  int64_t cell_begin, cell_end;
  int atomic;
  double p0, p1;  // lambda, mu  (or coef)
  double *bvec;       // local-vector target (VEC != 0)
  const double *fq;   // source at quadrature points [cell][p][3] or null
  double f0, f1, f2;  // constant source
  int nltot;          // row stride of the plan's rank map (= 3*NDS for one field; larger inside a multi-field plan)
  // Stokes coupling blocks (second field = scalar pressure), np1 = 0 when absent
  int np1;
  const double *N1;                    // [NP][np1]
  const int32_t *row_ids1, *col_ids1;  // [ncells][np1]
  int64_t row_off1, col_off1;
  // staged mode (affine_gather.cu, launch_staged_gather): the 3x3 blocks of the node pairs a <= b are written to
  // ke_out[cell][pair][9] instead of being scattered; the block-owner gather sums them per stored block afterwards
  double *ke_out;
};

__device__ __forceinline__ double inv3(const double *a, double *r) {
  double det = a[0] * a[4] * a[8] + a[1] * a[5] * a[6] + a[2] * a[3] * a[7] - (a[0] * a[5] * a[7] + a[1] * a[3] * a[8] + a[2] * a[4] * a[6]);
  double c = 1.0 / det;
  r[0] = (a[4] * a[8] - a[5] * a[7]) * c;
  r[1] = -(a[1] * a[8] - a[2] * a[7]) * c;
  r[2] = (a[1] * a[5] - a[2] * a[4]) * c;
  r[3] = -(a[3] * a[8] - a[5] * a[6]) * c;
  r[4] = (a[0] * a[8] - a[2] * a[6]) * c;
  r[5] = -(a[0] * a[5] - a[2] * a[3]) * c;
  r[6] = (a[3] * a[7] - a[4] * a[6]) * c;
  r[7] = -(a[0] * a[7] - a[1] * a[6]) * c;
  r[8] = (a[0] * a[4] - a[1] * a[3]) * c;
  return det;
}

constexpr int NH_STRIDE = 20;  // per quadrature point: Y[9] = F^-T, kappa, P[9] = F.S = mu F - kappa F^-T (first Piola stress, residual)

// Shared-memory scratch of one cell (in doubles).
template <int FORM, int VEC, int NDS, int NP>
struct CellScratch {
  static constexpr bool NEED_NH = (FORM == GB200_FORM_NEOHOOKEAN_JAC) || (VEC == GB200_FORM_NEOHOOKEAN_RES);
  static constexpr int NL = 3 * NDS;
  static constexpr int G = 0;                              // [NP][NDS][3] physical gradients
  static constexpr int IJ = G + NP * NDS * 3;              // [NP][9]
  static constexpr int DV = IJ + NP * 9;                   // [NP]
  static constexpr int NH = DV + NP;                       // [NP][NH_STRIDE]
  static constexpr int U = NH + (NEED_NH ? NP * NH_STRIDE : 0);  // [NL] dof values of u_h
  static constexpr int IDS = U + (NEED_NH ? NL : 0);       // int32 rows[NL], cols[NL]
  static constexpr int SIZE0 = IDS + NL + 1;               // (2 NL int32 = NL doubles)
  // cell stride = 8 mod 16 doubles: the gradients of two neighbouring cells of a warp (24-byte stride inside a cell: 16 of the 32
  // banks) fall on complementary banks
  static constexpr int SIZE = SIZE0 + (24 - SIZE0 % 16) % 16;
};

// ---- phases shared by the kernels of this file (flattened over the cells of the CTA's batch) -------------------------
template <class S, int NL, bool NEED_NH, int THREADS, class CellOf>
__device__ __forceinline__ void load_ids(const VArgs &k, double *smem, int nc, int tid, CellOf cell_of) {
  if (k.nzval && !k.ke_out) {  // the batch's slot-rank blocks (read by the scatter at the end of the batch) start their way up from HBM now
    const int rb = k.nltot * k.nltot * 2, lines = (rb + 127) / 128;
    for (int e = tid; e < nc * lines; e += THREADS) {
      const int c = e / lines, l = e - c * lines;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(k.rank + cell_of(c) * (int64_t)k.nltot * k.nltot) + l * 128));
    }
  }
  for (int e = tid; e < nc * NL; e += THREADS) {
    const int c = e / NL, l = e - c * NL;
    const int64_t cell = cell_of(c);
    int32_t *ids = reinterpret_cast<int32_t *>(smem + (size_t)c * S::SIZE + S::IDS);
    const int32_t row = k.row_ids[cell * NL + l], col = k.col_ids[cell * NL + l];
    ids[l] = row;
    ids[NL + l] = col;
    if (NEED_NH) {
      const int32_t sid = k.state_ids[cell * NL + l];
      smem[(size_t)c * S::SIZE + S::U + l] =
          sid > 0 ? (k.free_vals ? k.free_vals[sid - 1] : 0.0) : (sid < 0 && k.dir_vals ? k.dir_vals[-sid - 1] : 0.0);
    }
  }
}

template <class S, int NN, int NP, int THREADS, class CellOf>
__device__ __forceinline__ void geometry_phase(const VArgs &k, double *smem, int nc, int tid, CellOf cell_of) {
  for (int e = tid; e < nc * NP; e += THREADS) {
    const int c = e / NP, p = e - c * NP;
    const int64_t cell = cell_of(c);
    double *sc = smem + (size_t)c * S::SIZE;
    double Jt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int a = 0; a < NN; a++) {
      const double *x = k.X + (int64_t)k.cell_nodes[cell * NN + a] * 3;
      const double *dn = k.dNg + (p * NN + a) * 3;
      const double x0 = x[0], x1 = x[1], x2 = x[2];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        Jt[i * 3 + 0] += dn[i] * x0;
        Jt[i * 3 + 1] += dn[i] * x1;
        Jt[i * 3 + 2] += dn[i] * x2;
      }
    }
    const double det = inv3(Jt, sc + S::IJ + p * 9);
    sc[S::DV + p] = fabs(det) * k.w[p];
  }
}

template <class S, int NDS, int NP, int THREADS>
__device__ __forceinline__ void gradient_phase(const VArgs &k, double *smem, int nc, int tid) {
  for (int e = tid; e < nc * NP * NDS; e += THREADS) {
    const int c = e / (NP * NDS), r = e - c * (NP * NDS);
    const int p = r / NDS;
    double *sc = smem + (size_t)c * S::SIZE;
    const double *dn = k.dN + r * 3;
    const double *iJ = sc + S::IJ + p * 9;
    const double d0 = dn[0], d1 = dn[1], d2 = dn[2];
    double *g = sc + S::G + r * 3;
    g[0] = iJ[0] * d0 + iJ[1] * d1 + iJ[2] * d2;
    g[1] = iJ[3] * d0 + iJ[4] * d1 + iJ[5] * d2;
    g[2] = iJ[6] * d0 + iJ[7] * d1 + iJ[8] * d2;
  }
}

// Scatter of the 3x3 component block K[ci*3+cj] of the node pair (a, b) and, for a != b, of its transpose
// (K_e[(b,cj),(a,ci)] = K_e[(a,ci),(b,cj)]): slot = colptr[col] + rank, RED.ADD.F64 or plain RMW (coloured launches).
// DIAG: the form couples equal components only (vector Laplacian / mass: K = m * I3); the six off-diagonal entries of the
// block are structural zeros of the local matrix -- they are part of the pattern (nothing is dropped for being zero) but
// adding 0.0 to a zero-filled / already assembled value changes nothing, so their REDs are skipped.
template <int NDS, bool DIAG = false>
__device__ __forceinline__ void scatter_pair_block(const VArgs &k, const double *K, int a, int b, const int32_t *sRow, const int32_t *sCol,
                                                   const uint16_t *rk, int NLT) {
  // all slot-rank and column-base loads are issued before the first RED (they are independent; the scatter is latency-bound otherwise)
  int r1[9], r2[9];
  int64_t base1[3], base2[3];
  bool row_a[3], row_b[3];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const int32_t colb = sCol[b + NDS * c], cola = sCol[a + NDS * c];
    base1[c] = colb > 0 ? k.colptr[colb - 1 + k.col_off] : -1;
    base2[c] = (cola > 0 && a != b) ? k.colptr[cola - 1 + k.col_off] : -1;
    row_a[c] = sRow[a + NDS * c] > 0;
    row_b[c] = sRow[b + NDS * c] > 0;
  }
#pragma unroll
  for (int ci = 0; ci < 3; ci++)
#pragma unroll
    for (int cj = 0; cj < 3; cj++) {
      if (DIAG && ci != cj) continue;
      r1[ci * 3 + cj] = rk[(a + NDS * ci) + NLT * (b + NDS * cj)];   // row (a,ci), column (b,cj)
      r2[ci * 3 + cj] = rk[(b + NDS * cj) + NLT * (a + NDS * ci)];   // row (b,cj), column (a,ci)
    }
#pragma unroll
  for (int cj = 0; cj < 3; cj++)
#pragma unroll
    for (int ci = 0; ci < 3; ci++)
      if ((!DIAG || ci == cj) && base1[cj] >= 0 && row_a[ci]) {
        double *dst = k.nzval + base1[cj] + r1[ci * 3 + cj];
        if (k.atomic) atomicAdd(dst, K[ci * 3 + cj]); else *dst += K[ci * 3 + cj];
      }
#pragma unroll
  for (int ci = 0; ci < 3; ci++)
#pragma unroll
    for (int cj = 0; cj < 3; cj++)
      if ((!DIAG || ci == cj) && base2[ci] >= 0 && row_b[cj]) {
        double *dst = k.nzval + base2[ci] + r2[ci * 3 + cj];
        if (k.atomic) atomicAdd(dst, K[ci * 3 + cj]); else *dst += K[ci * 3 + cj];
      }
}

// A CTA of THREADS threads owns CELLS cells at a time; every phase runs over the flattened (cell, item) index space so
// that all lanes stay busy whatever the element size.  The local matrix is symmetric (all supported forms are symmetric
// bilinear forms / a hyperelastic tangent): only node pairs a <= b are integrated, the block and its transpose are
// scattered.  Linear elasticity / Laplacian with constant coefficients: the quadrature loop accumulates only
// A_ab = sum_p dV grad(phi_a) (x) grad(phi_b) (9 FMA per point); the constitutive algebra runs once per pair:
//   K[(a,ci),(b,cj)] = lambda A[ci][cj] + mu A[cj][ci] + delta_cicj mu tr(A).
template <int FORM, int VEC, int NN, int NDS, int NP, int CELLS, int THREADS>
__global__ void __launch_bounds__(THREADS, (FORM == GB200_FORM_NEOHOOKEAN_JAC && NDS == 8) ? 6 : 1) vector_kernel(VArgs k) {
  using S = CellScratch<FORM, VEC, NDS, NP>;
  constexpr bool NEED_NH = S::NEED_NH;
  constexpr int NL = 3 * NDS;
  constexpr int NPAIR = NDS * (NDS + 1) / 2;
  extern __shared__ double smem[];
  __shared__ unsigned char s_pa[NPAIR], s_pb[NPAIR];
  const int tid = threadIdx.x;
  for (int q = tid; q < NPAIR; q += THREADS) {  // q = b (b + 1) / 2 + a,  a <= b
    int b = (int)((sqrtf(8.0f * q + 1.0f) - 1.0f) * 0.5f);
    while ((b + 1) * (b + 2) / 2 <= q) b++;
    while (b * (b + 1) / 2 > q) b--;
    s_pb[q] = (unsigned char)b;
    s_pa[q] = (unsigned char)(q - b * (b + 1) / 2);
  }
  const int NLT = k.nltot;

  for (int64_t it0 = k.cell_begin + (int64_t)blockIdx.x * CELLS; it0 < k.cell_end; it0 += (int64_t)gridDim.x * CELLS) {
    const int nc = (int)min((int64_t)CELLS, k.cell_end - it0);
    auto cell_of = [&](int c) -> int64_t { return k.cell_list ? (int64_t)k.cell_list[it0 + c] : it0 + c; };
    __syncthreads();  // previous batch fully consumed (also orders the pair table)
    load_ids<S, NL, NEED_NH, THREADS>(k, smem, nc, tid, cell_of);
    geometry_phase<S, NN, NP, THREADS>(k, smem, nc, tid, cell_of);
    __syncthreads();
    gradient_phase<S, NDS, NP, THREADS>(k, smem, nc, tid);
    __syncthreads();
    // 3. neo-Hookean state per quadrature point
    if (NEED_NH) {
      for (int e = tid; e < nc * NP; e += THREADS) {
        const int c = e / NP, p = e - c * NP;
        double *sc = smem + (size_t)c * S::SIZE;
        const double *sG = sc + S::G, *sU = sc + S::U;
        double gu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // (grad u)[i][c] = sum_a u_{a,c} d_i N_a
#pragma unroll
        for (int cc = 0; cc < 3; cc++)
          for (int a = 0; a < NDS; a++) {
            const double u = sU[a + NDS * cc];
            const double *g = sG + (p * NDS + a) * 3;
            gu[0 * 3 + cc] += u * g[0];
            gu[1 * 3 + cc] += u * g[1];
            gu[2 * 3 + cc] += u * g[2];
          }
        // F = I + (grad u)^T;  Y = F^-T maps a material gradient ga to the spatial one (alpha_c = ga . Y[c]);  J = det F;
        // S = mu (I - C^-1) + lambda ln J C^-1  =>  F.S = mu F - kappa F^-T,  kappa = mu - lambda ln J
        double F[9], Fi[9];
        for (int i = 0; i < 3; i++)
          for (int j = 0; j < 3; j++) F[i * 3 + j] = (i == j ? 1.0 : 0.0) + gu[j * 3 + i];
        const double detF = inv3(F, Fi);
        const double lnJ = log(fabs(detF));
        const double kap = k.p1 - k.p0 * lnJ;
        double *o = sc + S::NH + p * NH_STRIDE;
        for (int cc = 0; cc < 3; cc++)
          for (int i = 0; i < 3; i++) {
            const double y = Fi[i * 3 + cc];
            o[cc * 3 + i] = y;
            o[10 + cc * 3 + i] = k.p1 * F[cc * 3 + i] - kap * y;
          }
        o[9] = kap;
      }
      __syncthreads();
    }
    // 4. node pairs a <= b: 3x3 component block K[ci*3+cj] of ((a,ci),(b,cj)), scattered together with its transpose
    // neo-Hookean Jacobian: with alpha = F^-T ga (= ga . Y[c], the spatial gradient), beta likewise, Z = F C^-1 F^T = I,
    // c_ab = alpha . beta and s_ab = mu ga.gb - kappa c_ab the integrand of the file header collapses to
    //   K[ci][cj] += dV [ lambda alpha_ci beta_cj + kappa alpha_cj beta_ci + delta_cicj mu ga.gb ]
    // A thread owns three pairs of ONE cell and runs the quadrature loop outermost, so Y and kappa of a point are read from
    // shared memory once per three pairs.
    constexpr bool NH3 = FORM == GB200_FORM_NEOHOOKEAN_JAC && NPAIR % 3 == 0 && CELLS * (NPAIR / 3) <= THREADS;
    double K3[3][9];   // (NH3 only; kept in registers until the end of the batch in staged mode)
    int pa3[3] = {0, 0, 0}, pb3[3] = {0, 0, 0};
    if (NH3) {
      constexpr int TPC = NPAIR / 3;  // threads per cell
      if (tid < nc * TPC) {
        const int c = tid / TPC, j0 = tid - c * TPC;
        const int64_t cell = cell_of(c);
        const double *sc = smem + (size_t)c * S::SIZE;
        const double *sG = sc + S::G, *sdV = sc + S::DV, *sNH = sc + S::NH;
        const int32_t *sRow = reinterpret_cast<const int32_t *>(sc + S::IDS), *sCol = sRow + NL;
#pragma unroll
        for (int r = 0; r < 3; r++) { pa3[r] = s_pa[j0 + r * TPC]; pb3[r] = s_pb[j0 + r * TPC]; }
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int i = 0; i < 9; i++) K3[r][i] = 0.0;
#pragma unroll 1
        for (int p = 0; p < NP; p++) {
          const double *o = sNH + p * NH_STRIDE;
          double Y[9];
#pragma unroll
          for (int i = 0; i < 9; i++) Y[i] = o[i];
          const double dv = sdV[p];
          const double lam_dv = k.p0 * dv, kap_dv = o[9] * dv, mu_dv = k.p1 * dv;
#pragma unroll
          for (int r = 0; r < 3; r++) {
            const double *ga = sG + (p * NDS + pa3[r]) * 3, *gb = sG + (p * NDS + pb3[r]) * 3;
            const double a0 = ga[0], a1 = ga[1], a2 = ga[2], b0 = gb[0], b1 = gb[1], b2 = gb[2];
            double al[3], be[3];
#pragma unroll
            for (int cc = 0; cc < 3; cc++) {
              al[cc] = a0 * Y[cc * 3 + 0] + a1 * Y[cc * 3 + 1] + a2 * Y[cc * 3 + 2];
              be[cc] = b0 * Y[cc * 3 + 0] + b1 * Y[cc * 3 + 1] + b2 * Y[cc * 3 + 2];
            }
            const double gab = mu_dv * (a0 * b0 + a1 * b1 + a2 * b2);
            const double la[3] = {lam_dv * al[0], lam_dv * al[1], lam_dv * al[2]};
            const double ka[3] = {kap_dv * al[0], kap_dv * al[1], kap_dv * al[2]};
#pragma unroll
            for (int ci = 0; ci < 3; ci++)
#pragma unroll
              for (int cj = 0; cj < 3; cj++)
                K3[r][ci * 3 + cj] += la[ci] * be[cj] + ka[cj] * be[ci] + (ci == cj ? gab : 0.0);
          }
        }
        if (!k.ke_out) {
          const uint16_t *rk = k.rank + cell * (int64_t)NLT * NLT;
#pragma unroll
          for (int r = 0; r < 3; r++) scatter_pair_block<NDS>(k, K3[r], pa3[r], pb3[r], sRow, sCol, rk, NLT);
        }
      }
    } else
    if (FORM != GB200_FORM_NONE)
    for (int e = tid; e < nc * NPAIR; e += THREADS) {
      const int c = e / NPAIR, q = e - c * NPAIR;
      const int a = s_pa[q], b = s_pb[q];
      const int64_t cell = cell_of(c);
      const double *sc = smem + (size_t)c * S::SIZE;
      const double *sG = sc + S::G, *sdV = sc + S::DV;
      const int32_t *sRow = reinterpret_cast<const int32_t *>(sc + S::IDS), *sCol = sRow + NL;
      double K[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      if (FORM == GB200_FORM_MASS) {
        double m = 0.0;
        for (int p = 0; p < NP; p++) m += k.N[p * NDS + a] * k.N[p * NDS + b] * sdV[p];
        K[0] = K[4] = K[8] = k.p0 * m;
      } else if (FORM == GB200_FORM_LAPLACIAN) {
        double m = 0.0;
        for (int p = 0; p < NP; p++) {
          const double *ga = sG + (p * NDS + a) * 3, *gb = sG + (p * NDS + b) * 3;
          m += (ga[0] * gb[0] + ga[1] * gb[1] + ga[2] * gb[2]) * sdV[p];
        }
        K[0] = K[4] = K[8] = k.p0 * m;
      } else if (FORM == GB200_FORM_ELASTICITY) {
        double A[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // A[i*3+j] = sum_p dV ga_i gb_j
        for (int p = 0; p < NP; p++) {
          const double dv = sdV[p];
          const double *ga = sG + (p * NDS + a) * 3, *gb = sG + (p * NDS + b) * 3;
          const double t0 = dv * ga[0], t1 = dv * ga[1], t2 = dv * ga[2];
          const double b0 = gb[0], b1 = gb[1], b2 = gb[2];
          A[0] += t0 * b0; A[1] += t0 * b1; A[2] += t0 * b2;
          A[3] += t1 * b0; A[4] += t1 * b1; A[5] += t1 * b2;
          A[6] += t2 * b0; A[7] += t2 * b1; A[8] += t2 * b2;
        }
        const double mtr = k.p1 * (A[0] + A[4] + A[8]);
#pragma unroll
        for (int ci = 0; ci < 3; ci++)
#pragma unroll
          for (int cj = 0; cj < 3; cj++) K[ci * 3 + cj] = k.p0 * A[ci * 3 + cj] + k.p1 * A[cj * 3 + ci] + (ci == cj ? mtr : 0.0);
      } else {  // neo-Hookean Jacobian (collapsed integrand, see above)
        const double *sNH = sc + S::NH;
        for (int p = 0; p < NP; p++) {
          const double dv = sdV[p];
          const double *ga = sG + (p * NDS + a) * 3, *gb = sG + (p * NDS + b) * 3;
          const double a0 = ga[0], a1 = ga[1], a2 = ga[2], b0 = gb[0], b1 = gb[1], b2 = gb[2];
          const double *Y = sNH + p * NH_STRIDE;
          const double lam_dv = k.p0 * dv, kap_dv = Y[9] * dv;
          double al[3], be[3];
#pragma unroll
          for (int cc = 0; cc < 3; cc++) {
            al[cc] = a0 * Y[cc * 3 + 0] + a1 * Y[cc * 3 + 1] + a2 * Y[cc * 3 + 2];
            be[cc] = b0 * Y[cc * 3 + 0] + b1 * Y[cc * 3 + 1] + b2 * Y[cc * 3 + 2];
          }
          const double gab = k.p1 * dv * (a0 * b0 + a1 * b1 + a2 * b2);
#pragma unroll
          for (int ci = 0; ci < 3; ci++)
#pragma unroll
            for (int cj = 0; cj < 3; cj++)
              K[ci * 3 + cj] += lam_dv * al[ci] * be[cj] + kap_dv * al[cj] * be[ci] + (ci == cj ? gab : 0.0);
        }
      }
      if (k.ke_out) {   // staged mode: the block of the pair goes to ke_out[cell][q][9]
        double *o = k.ke_out + (cell * NPAIR + q) * 9;
#pragma unroll
        for (int i = 0; i < 9; i++) o[i] = K[i];
        continue;
      }
      scatter_pair_block<NDS, FORM == GB200_FORM_MASS || FORM == GB200_FORM_LAPLACIAN>(k, K, a, b, sRow, sCol, k.rank + cell * (int64_t)NLT * NLT, NLT);
    }
    // 4b. Stokes coupling blocks: T[c] = sum_p d_c N_a psi_b dV ;  (v,p) entry = -T, (q,u) entry = +T  (StokesTaylorHoodTests.jl:59)
    if (FORM == GB200_FORM_LAPLACIAN && VEC == 0 && k.np1 > 0) {
      const int np1 = k.np1;
      for (int e = tid; e < nc * NDS * np1; e += THREADS) {
        const int c = e / (NDS * np1), pr = e - c * (NDS * np1);
        const int b = pr / NDS, a = pr - b * NDS;
        const int64_t cell = cell_of(c);
        const double *sc = smem + (size_t)c * S::SIZE;
        const double *sG = sc + S::G, *sdV = sc + S::DV;
        const int32_t *sRow = reinterpret_cast<const int32_t *>(sc + S::IDS), *sCol = sRow + NL;
        const uint16_t *rk = k.rank + cell * (int64_t)NLT * NLT;
        double T0 = 0.0, T1 = 0.0, T2 = 0.0;
        for (int p = 0; p < NP; p++) {
          const double wv = k.N1[p * np1 + b] * sdV[p];
          const double *ga = sG + (p * NDS + a) * 3;
          T0 += ga[0] * wv; T1 += ga[1] * wv; T2 += ga[2] * wv;
        }
        const double T[3] = {T0, T1, T2};
        const int32_t prow = k.row_ids1[cell * np1 + b], pcol = k.col_ids1[cell * np1 + b];
        const int lp = NL + b;  // local index of the pressure dof in the concatenated numbering
        // all slot loads first, then the REDs
        const int64_t pbase = pcol > 0 ? k.colptr[pcol - 1 + k.col_off1] : -1;
        int64_t vbase[3];
        int rvp[3], rqu[3];
        bool vrow[3];
#pragma unroll
        for (int cc = 0; cc < 3; cc++) {
          const int lv = a + NDS * cc;
          const int32_t vc = sCol[lv];
          vbase[cc] = (prow > 0 && vc > 0) ? k.colptr[vc - 1 + k.col_off] : -1;
          vrow[cc] = sRow[lv] > 0;
          rvp[cc] = rk[lv + NLT * lp];
          rqu[cc] = rk[lp + NLT * lv];
        }
#pragma unroll
        for (int cc = 0; cc < 3; cc++) {
          if (pbase >= 0 && vrow[cc]) {  // (v,p): row = velocity test dof, column = pressure trial dof
            double *dst = k.nzval + pbase + rvp[cc];
            if (k.atomic) atomicAdd(dst, -T[cc]); else *dst -= T[cc];
          }
          if (vbase[cc] >= 0) {  // (q,u): row = pressure test dof, column = velocity trial dof
            double *dst = k.nzval + vbase[cc] + rqu[cc];
            if (k.atomic) atomicAdd(dst, T[cc]); else *dst += T[cc];
          }
        }
      }
    }
    // 5. local vector: source term or neo-Hookean residual
    if (VEC != 0) {
      for (int e = tid; e < nc * NL; e += THREADS) {
        const int c = e / NL, li = e - c * NL;
        const double *sc = smem + (size_t)c * S::SIZE;
        const int32_t row = reinterpret_cast<const int32_t *>(sc + S::IDS)[li];
        if (row <= 0) continue;
        const double *sG = sc + S::G, *sdV = sc + S::DV;
        const int ci = li / NDS, a = li - ci * NDS;
        double v = 0.0;
        for (int p = 0; p < NP; p++) {
          if (VEC == GB200_FORM_SOURCE) {
            const double f = k.fq ? k.fq[(cell_of(c) * NP + p) * 3 + ci] : (ci == 0 ? k.f0 : ci == 1 ? k.f1 : k.f2);
            v += k.N[p * NDS + a] * f * sdV[p];
          } else {
            const double *g = sG + (p * NDS + a) * 3;
            const double *sf = sc + S::NH + p * NH_STRIDE + 10 + ci * 3;
            v += (g[0] * sf[0] + g[1] * sf[1] + g[2] * sf[2]) * sdV[p];
          }
        }
        double *dst = k.bvec + (row - 1 + k.row_off);
        if (k.atomic) atomicAdd(dst, v); else *dst += v;
      }
    }
    // 6. staged mode, neo-Hookean blocks held in registers: through the (now dead) scratch of the cell, then contiguous stores
    if (NH3 && k.ke_out) {
      constexpr int TPC = NPAIR / 3;
      constexpr bool VIA_SMEM = NPAIR * 9 <= S::U;   // (Q2: the blocks of a cell exceed its scratch -- straight from the registers)
      if (VIA_SMEM) {
        __syncthreads();
        if (tid < nc * TPC) {
          const int c = tid / TPC, j0 = tid - c * TPC;
          double *st = smem + (size_t)c * S::SIZE;
#pragma unroll
          for (int r = 0; r < 3; r++)
#pragma unroll
            for (int i = 0; i < 9; i++) st[(j0 + r * TPC) * 9 + i] = K3[r][i];
        }
        __syncthreads();
        for (int e = tid; e < nc * NPAIR * 9; e += THREADS) {
          const int c = e / (NPAIR * 9), r = e - c * (NPAIR * 9);
          k.ke_out[cell_of(c) * (NPAIR * 9) + r] = smem[(size_t)c * S::SIZE + r];
        }
      } else if (tid < nc * TPC) {
        const int c = tid / TPC, j0 = tid - c * TPC;
        double *o = k.ke_out + cell_of(c) * (NPAIR * 9);
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int i = 0; i < 9; i++) o[(j0 + r * TPC) * 9 + i] = K3[r][i];
      }
    }
  }
}

template <int FORM, int VEC, int NN, int NDS, int NP, int CELLS, int THREADS>
void launch_one(gb200_plan plan, VArgs &k) {
  gb200_ctx ctx = plan->ctx;
  using S = CellScratch<FORM, VEC, NDS, NP>;
  const size_t smem = (size_t)CELLS * S::SIZE * sizeof(double);
  auto kern = vector_kernel<FORM, VEC, NN, NDS, NP, CELLS, THREADS>;
  static std::map<int, int> cps_of_device;  // per instantiation and device (the opt-in attribute belongs to the device's function)
  int &ctas_per_sm = cps_of_device[ctx->device];
  if (ctas_per_sm == 0) {
    if (smem > 48 * 1024) GB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, THREADS, smem));
    ctas_per_sm = std::max(ctas_per_sm, 1);
  }
  auto launch = [&](int64_t begin, int64_t end, const int32_t *list, int atomic) {
    if (end <= begin) return;
    k.cell_begin = begin; k.cell_end = end; k.cell_list = list; k.atomic = atomic;
    int64_t nblocks = (end - begin + CELLS - 1) / CELLS;
    int grid = (int)std::min<int64_t>(nblocks, (int64_t)ctx->num_sms * ctas_per_sm);  // persistent: one wave
    kern<<<grid, THREADS, smem, ctx->stream>>>(k);
    check_launch(ctx, "vector_kernel");
  };
  if (k.ke_out && VEC == 0) {
    launch(0, plan->mesh->ncells, nullptr, 0);   // staged mode: every cell writes its own blocks, nothing to colour
  } else if (ctx->deterministic()) {
    for (int c = 0; c < plan->ncolors; c++) launch(plan->color_ptr[c], plan->color_ptr[c + 1], plan->color_cells.p, 0);
  } else {
    launch(0, plan->mesh->ncells, nullptr, 1);
  }
}

// ---- Q2 hexahedra, linear elasticity: the local contraction on the FP64 tensor cores --------------------------------------
// A[(a,i),(b,j)] = sum_p dV_p d_i phi_a(x_p) d_j phi_b(x_p) is the 81 x 81 x 27 GEMM  A = (dV G)^T G  with G[p][a + 27 i] the physical
// gradients (a8 of SURVEY.md section 8: the IntegrationMap contraction, src/Fields/FieldsInterfaces.jl:737-760).  It runs as
// mma.sync.m8n8k4.f64 (DMMA) over 8x8 tiles of the upper triangle (A is symmetric; the mirror image is written on store), operands
// read from shared memory, 2 tile rows per pass so that a B fragment feeds two MMAs.  The epilogue applies the constitutive law
// per entry,  K[(a,ci),(b,cj)] = lambda A[(a,ci),(b,cj)] + mu A[(a,cj),(b,ci)] + delta_cicj mu sum_k A[(a,k),(b,k)],
// with lanes running over the rows in node-major order (a, ci): the three components of a node are adjacent CSC rows, so a
// warp-level RED touches ~16 sectors instead of ~28 with the node-pair mapping of vector_kernel.
namespace q2mma {
constexpr int NDS = 27, NP = 27, NL = 81, NN = 8;
constexpr int LDG = 88, KP = 28;         // G padded to [28][88], zero rows / columns beyond 27 / 81
constexpr int LDA = 81;                  // sA[n][m] (symmetric), odd stride
constexpr int NT = 11;                   // 8x8 tiles per dimension
constexpr int THREADS = 192;              // 6 warps: one GEMM pass each; scatter: 2 x 81 row owners
constexpr int OFF_G = 0;
constexpr int OFF_A = OFF_G + KP * LDG;
constexpr int OFF_DV = OFF_A + NL * LDA;
constexpr int OFF_CB = OFF_DV + KP;      // int64 colbase[81]
constexpr int OFF_IDS = OFF_CB + NL;     // int32 rows[81], cols[81]
constexpr int OFF_DNG = OFF_IDS + NL + 1; // dNg[27][8][3] (geometry shape-function gradients)
constexpr int OFF_X = OFF_DNG + NP * NN * 3;  // node coordinates of the current / next cell, [2][8][3]
constexpr int SMEM_DOUBLES = OFF_X + 2 * NN * 3;
constexpr int GITEMS = (NP * NDS + THREADS - 1) / THREADS;  // (p, a) gradient items per thread
}  // namespace q2mma

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(q2mma::THREADS) q2_elasticity_mma_kernel(VArgs k) {
  using namespace q2mma;
  extern __shared__ double smem[];
  double *sG = smem + OFF_G, *sA = smem + OFF_A, *sdV = smem + OFF_DV;
  int64_t *sCB = reinterpret_cast<int64_t *>(smem + OFF_CB);
  int32_t *sRow = reinterpret_cast<int32_t *>(smem + OFF_IDS), *sCol = sRow + NL;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NLT = k.nltot;
  double *sdNg = smem + OFF_DNG, *sX = smem + OFF_X;
  // zero padding of G (never overwritten below)
  for (int e = tid; e < KP * LDG; e += THREADS) {
    const int p = e / LDG, m = e - p * LDG;
    if (p >= NP || m >= NL) sG[e] = 0.0;
  }
  if (tid == 0) sdV[NP] = 0.0;
  for (int e = tid; e < NP * NN * 3; e += THREADS) sdNg[e] = k.dNg[e];
  // the reference gradients of this thread's (p, a) items stay in registers for the whole kernel
  double dn_reg[GITEMS][3];
#pragma unroll
  for (int i = 0; i < GITEMS; i++) {
    const int e = tid + i * THREADS;
#pragma unroll
    for (int d = 0; d < 3; d++) dn_reg[i][d] = e < NP * NDS ? k.dN[e * 3 + d] : 0.0;
  }
  // node coordinates are fetched one cell ahead (the ids -> coordinates chain overlaps the previous cell's work)
  auto fetch_x = [&](int64_t itx) -> double {
    const int64_t c = k.cell_list ? (int64_t)k.cell_list[itx] : itx;
    return k.X[(int64_t)k.cell_nodes[c * NN + tid / 3] * 3 + tid % 3];
  };
  int xbuf = 0;
  if (tid < NN * 3 && k.cell_begin + blockIdx.x < k.cell_end) sX[tid] = fetch_x(k.cell_begin + blockIdx.x);

  for (int64_t it = k.cell_begin + blockIdx.x; it < k.cell_end; it += gridDim.x, xbuf ^= 1) {
    const int64_t cell = k.cell_list ? (int64_t)k.cell_list[it] : it;
    __syncthreads();  // previous cell fully scattered
    double x_next = 0.0;
    const bool have_next = it + gridDim.x < k.cell_end;
    if (tid < NN * 3 && have_next) x_next = fetch_x(it + gridDim.x);
    {  // the cell's rank block (13 KB, read by the scatter phase) starts its way up from HBM now
      const char *rkb = reinterpret_cast<const char *>(k.rank + cell * (int64_t)NLT * NLT);
      for (int o = tid * 128; o < NLT * NLT * 2; o += THREADS * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(rkb + o));
    }
    // 0. ids and column bases
    for (int l = tid; l < NL; l += THREADS) {
      const int32_t col = k.col_ids[cell * NL + l];
      sRow[l] = k.row_ids[cell * NL + l];
      sCol[l] = col;
      sCB[l] = col > 0 ? k.colptr[col - 1 + k.col_off] : -1;  // -1: column not stored (Dirichlet / masked)
    }
    // 1. Jacobians Jt[p][i][j] = sum_a d_i N_a(q_p) x_a,j, one entry per thread, then their inverses, one row per thread
    //    (both staged in the not-yet-used A buffer)
    double *siJ = sA, *sJ = sA + 256;
    for (int e = tid; e < NP * 9; e += THREADS) {
      const int p = e / 9, ij = e - p * 9, i = ij / 3, j = ij - 3 * i;
      const double *xc = sX + xbuf * NN * 3 + j;
      const double *dn = sdNg + p * NN * 3 + i;
      double v = 0.0;
#pragma unroll
      for (int a = 0; a < NN; a++) v += dn[a * 3] * xc[a * 3];
      sJ[e] = v;
    }
    __syncthreads();
    if (tid < NP * 3) {
      const int p = tid / 3, i = tid - 3 * p;
      const double *J = sJ + p * 9;
      const double det = J[0] * J[4] * J[8] + J[1] * J[5] * J[6] + J[2] * J[3] * J[7] - (J[0] * J[5] * J[7] + J[1] * J[3] * J[8] + J[2] * J[4] * J[6]);
      const double c = 1.0 / det;
      const int i1 = i == 2 ? 0 : i + 1, i2 = i == 0 ? 2 : i - 1;   // (i+1)%3, (i+2)%3
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        siJ[p * 9 + i * 3 + j] = (J[j1 * 3 + i1] * J[j2 * 3 + i2] - J[j1 * 3 + i2] * J[j2 * 3 + i1]) * c;  // cofactor(j,i) / det
      }
      if (i == 0) sdV[p] = fabs(det) * k.w[p];
    }
    __syncthreads();
    // 2. physical gradients G[p][a + 27 i]
#pragma unroll
    for (int i = 0; i < GITEMS; i++) {
      const int e = tid + i * THREADS;
      if (e < NP * NDS) {
        const int p = e / NDS, a = e - p * NDS;
        const double *iJ = siJ + p * 9;
        const double d0 = dn_reg[i][0], d1 = dn_reg[i][1], d2 = dn_reg[i][2];
        double *g = sG + p * LDG + a;
        g[0] = iJ[0] * d0 + iJ[1] * d1 + iJ[2] * d2;
        g[NDS] = iJ[3] * d0 + iJ[4] * d1 + iJ[5] * d2;
        g[2 * NDS] = iJ[6] * d0 + iJ[7] * d1 + iJ[8] * d2;
      }
    }
    if (tid < NN * 3 && have_next) sX[(xbuf ^ 1) * NN * 3 + tid] = x_next;  // read after the next loop-top barrier
    __syncthreads();
    // 3. A = (dV G)^T G on the tensor cores, upper-triangle tiles only.  Tile row t has 11 - t tiles: warp w owns rows
    //    (w, 11 - w) -- (0), (1,10), (2,9), (3,8), (4,7), (5,6) -- 11 tiles each; both rows share the B fragments.
    {
      const int ti = warp, tj = warp == 0 ? NT : NT - warp;  // tj = NT: no second row
      double acc0[NT][2], acc1[NT][2];
#pragma unroll
      for (int t = 0; t < NT; t++) { acc0[t][0] = acc0[t][1] = acc1[t][0] = acc1[t][1] = 0.0; }
      const int r4 = lane & 3, q8 = lane >> 2;
#pragma unroll 1
      for (int k0 = 0; k0 < KP; k0 += 4) {
        const double *grow = sG + (k0 + r4) * LDG;
        const double dv = sdV[k0 + r4];
        const double a0 = dv * grow[ti * 8 + q8];
        const double a1 = tj < NT ? dv * grow[tj * 8 + q8] : 0.0;
#pragma unroll
        for (int t = 0; t < NT; t++) {
          if (t < ti) continue;  // (warp-uniform) lower-triangle tiles come from the mirror store
          const double b = grow[t * 8 + q8];
          dmma_m8n8k4(acc0[t][0], acc0[t][1], a0, b);
          if (t >= tj) dmma_m8n8k4(acc1[t][0], acc1[t][1], a1, b);
        }
      }
      // store: lane holds C[row q8][cols 2 r4, 2 r4 + 1] of each tile; sA[n][m] = A[m][n] and its mirror
#pragma unroll
      for (int t = 0; t < NT; t++) {
        if (t < ti) continue;
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int trow = h == 0 ? ti : tj;
          if (h == 1 && t < tj) continue;
          const int m = trow * 8 + q8;
          if (m >= NL) continue;
#pragma unroll
          for (int j = 0; j < 2; j++) {
            const int n = t * 8 + 2 * r4 + j;
            if (n >= NL) continue;
            const double v = h == 0 ? acc0[t][j] : acc1[t][j];
            sA[n * LDA + m] = v;
            if (t != trow) sA[m * LDA + n] = v;  // mirror image of an off-diagonal tile
          }
        }
      }
    }
    __syncthreads();
    // 4. constitutive law + scatter.  A thread owns one row r = (a, ci) (node-major order, so that the lanes of a warp hit
    //    adjacent CSC rows) and half of the trial nodes b; per b its three entries cj = 0..2 need A[(a,ci),(b,:)],
    //    A[(a,:),(b,ci)] and the trace of the 3x3 block.  All rank loads of the thread are issued before the first use.
    if (tid < 2 * NL) {
      const int half = tid >= NL ? 1 : 0, r = tid - half * NL;
      const int a = r / 3, ci = r - 3 * a, li = a + NDS * ci;
      constexpr int NB = 14;
      const int b0 = half * NB, nb = half ? NDS - NB : NB;
      if (sRow[li] > 0) {
        const uint16_t *q = k.rank + cell * (int64_t)NLT * NLT + li + (int64_t)NLT * b0;
        int rk0[NB], rk1[NB], rk2[NB];
#pragma unroll
        for (int i = 0; i < NB; i++) {
          if (i < nb) {
            rk0[i] = __ldg(q + i * NLT);
            rk1[i] = __ldg(q + i * NLT + NLT * NDS);
            rk2[i] = __ldg(q + i * NLT + 2 * NLT * NDS);
          }
        }
        const double lam = k.p0, mu = k.p1;
#pragma unroll
        for (int i = 0; i < NB; i++) {
          if (i < nb) {
            const int b = b0 + i;
            const double *Ab = sA + b * LDA;  // rows (b, j) of the symmetric A are Ab + 27 j LDA
            const double x0 = Ab[li], x1 = Ab[NDS * LDA + li], x2 = Ab[2 * NDS * LDA + li];   // A[(a,ci),(b,cj)]
            const double *Ay = Ab + ci * NDS * LDA + a;
            const double y0 = Ay[0], y1 = Ay[NDS], y2 = Ay[2 * NDS];                           // A[(a,cj),(b,ci)]
            const double tr = mu * (Ab[a] + Ab[NDS * LDA + a + NDS] + Ab[2 * NDS * LDA + a + 2 * NDS]);
            const double v0 = lam * x0 + mu * y0 + (ci == 0 ? tr : 0.0);
            const double v1 = lam * x1 + mu * y1 + (ci == 1 ? tr : 0.0);
            const double v2 = lam * x2 + mu * y2 + (ci == 2 ? tr : 0.0);
            const int64_t c0 = sCB[b], c1 = sCB[b + NDS], c2 = sCB[b + 2 * NDS];
            if (k.atomic) {
              if (c0 >= 0) atomicAdd(k.nzval + c0 + rk0[i], v0);
              if (c1 >= 0) atomicAdd(k.nzval + c1 + rk1[i], v1);
              if (c2 >= 0) atomicAdd(k.nzval + c2 + rk2[i], v2);
            } else {
              if (c0 >= 0) k.nzval[c0 + rk0[i]] += v0;
              if (c1 >= 0) k.nzval[c1 + rk1[i]] += v1;
              if (c2 >= 0) k.nzval[c2 + rk2[i]] += v2;
            }
          }
        }
      }
    }
  }
}

bool launch_q2_elasticity_mma(gb200_plan plan, VArgs &k) {
  static const bool disabled = getenv("GB200_NO_DMMA") != nullptr;
  if (disabled) return false;
  gb200_ctx ctx = plan->ctx;
  const size_t smem = (size_t)q2mma::SMEM_DOUBLES * sizeof(double);
  auto kern = q2_elasticity_mma_kernel;
  static std::map<int, int> cps_of_device;
  int &ctas_per_sm = cps_of_device[ctx->device];
  if (ctas_per_sm == 0) {
    GB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, q2mma::THREADS, smem));
    ctas_per_sm = std::max(ctas_per_sm, 1);
  }
  auto launch = [&](int64_t begin, int64_t end, const int32_t *list, int atomic) {
    if (end <= begin) return;
    k.cell_begin = begin; k.cell_end = end; k.cell_list = list; k.atomic = atomic;
    int grid = (int)std::min<int64_t>(end - begin, (int64_t)ctx->num_sms * ctas_per_sm);
    kern<<<grid, q2mma::THREADS, smem, ctx->stream>>>(k);
    check_launch(ctx, "q2_elasticity_mma_kernel");
  };
  if (ctx->deterministic()) {
    for (int c = 0; c < plan->ncolors; c++) launch(plan->color_ptr[c], plan->color_ptr[c + 1], plan->color_cells.p, 0);
  } else {
    launch(0, plan->mesh->ncells, nullptr, 1);
  }
  return true;
}


// ---- Q1 hexahedra, neo-Hookean Jacobian (+ residual), staged mode: a kernel shaped by what limits it -------------------------
// The node-pair kernel above spends ~600 shared-memory wavefronts and ~23 k FP64 lane-operations per cell on this form (ncu: LSU
// pipe 70 %, FP64 35 %): every pair re-derives the spatial gradients alpha = F^-T grad(phi) from Y, and the geometry / state phases
// walk shared memory with conflicting strides.  Here
//   phase A  a thread owns one (cell, quadrature point): reference gradients of its point live in registers for the whole kernel;
//            Jt, inv(Jt), grad(phi_a), grad(u), F, F^-T, ln J, the first Piola stress and the spatial gradients alpha_a in
//            registers; it leaves [a][alpha(3) | grad(3)] and lambda dV, kappa dV, mu dV in shared memory and its share of the
//            residual;
//   phase B  a thread owns a 2x2 TILE of node pairs of one cell (nodes in 4 groups of 2: 10 tiles cover the pairs a <= b): per
//            point it loads 4 nodes (24 doubles) for 100 FMAs,  K[ci][cj] += (lambda dV) alpha_ci beta_cj + (kappa dV) alpha_cj beta_ci
//            + delta mu dV grad_a.grad_b  -- 36 independent accumulators per thread;
//   the blocks of the batch leave through shared memory as contiguous stores of ke_out[cell][pair][9].
namespace nhq1 {
constexpr int CELLS = 8, THREADS = 80, NP = 8, ND = 8, NPAIR = 36, TILES = 10;
constexpr int PSTRIDE = ND * 6 + 4;          // per (cell, point): [a][alpha0..2, g0..2], lambda dV, kappa dV, mu dV, pad
constexpr int CSTRIDE = NP * PSTRIDE + 2;    // = 2 mod 16 doubles: the 16-byte chunks a warp loads (4 cells x 4 node groups) tile the banks twice
constexpr int XSTRIDE = 50, RSTRIDE = 25;    // cell stride of X | U, (cell, point) stride of the residual partials: conflict-free
constexpr int O_Q = 0;
constexpr int O_XU = O_Q + CELLS * CSTRIDE;  // [cell][X: a*3 + d | U: 24 + a + 8*c]
constexpr int O_K = O_XU + CELLS * XSTRIDE;  // residual partials [cell][p][RSTRIDE], later the staged blocks [cell][324]
constexpr int O_IDS = O_K + CELLS * NPAIR * 9;   // int32: nodes [cell][8], state ids [cell][24], row ids [cell][24]
constexpr int SMEM_DOUBLES = O_IDS + CELLS * 28;
__constant__ unsigned char c_tile_i[TILES] = {0, 0, 0, 0, 1, 1, 1, 2, 2, 3};
__constant__ unsigned char c_tile_j[TILES] = {0, 1, 2, 3, 1, 2, 3, 2, 3, 3};
}  // namespace nhq1

template <int VEC>
__global__ void __launch_bounds__(nhq1::THREADS, 4) nh_q1_staged_kernel(VArgs k) {
  using namespace nhq1;
  extern __shared__ double smem[];
  double *sQ = smem + O_Q, *sXU = smem + O_XU, *sK = smem + O_K;
  int32_t *sNodes = reinterpret_cast<int32_t *>(smem + O_IDS), *sSid = sNodes + CELLS * 8, *sRow = sSid + CELLS * 24;
  const int tid = threadIdx.x;
  // phase A identity: (cell, point); the reference gradients of the point stay in registers (geometry map = field basis for Q1)
  const int ca = tid >> 3, pa = tid & 7;
  double dn[ND][3];
#pragma unroll
  for (int a = 0; a < ND; a++)
#pragma unroll
    for (int d = 0; d < 3; d++) dn[a][d] = k.dN[(pa * ND + a) * 3 + d];
  const double wp = k.w[pa];
  // phase B identity: (cell, tile)
  const int cb = tid / TILES, tb = tid - cb * TILES;
  const int a0 = 2 * c_tile_i[tb], b0 = 2 * c_tile_j[tb];

  for (int64_t it0 = k.cell_begin + (int64_t)blockIdx.x * CELLS; it0 < k.cell_end; it0 += (int64_t)gridDim.x * CELLS) {
    const int nc = (int)min((int64_t)CELLS, k.cell_end - it0);
    auto cell_of = [&](int c) -> int64_t { return k.cell_list ? (int64_t)k.cell_list[it0 + c] : it0 + c; };
    __syncthreads();   // previous batch fully written out
    // 0. ids of the batch (coalesced when the cells are consecutive), then node coordinates and dof values
    for (int e = tid; e < nc * 8; e += THREADS) sNodes[e] = k.cell_nodes[cell_of(e >> 3) * 8 + (e & 7)];
    for (int e = tid; e < nc * 24; e += THREADS) {
      const int c = e / 24, l = e - c * 24;
      sSid[e] = k.state_ids[cell_of(c) * 24 + l];
      if (VEC) sRow[e] = k.row_ids[cell_of(c) * 24 + l];
    }
    __syncthreads();
    for (int e = tid; e < nc * 24; e += THREADS) {
      const int c = e / 24, l = e - c * 24;
      sXU[c * XSTRIDE + l] = k.X[(int64_t)sNodes[c * 8 + l / 3] * 3 + l % 3];
      const int32_t sid = sSid[e];
      sXU[c * XSTRIDE + 24 + l] = sid > 0 ? (k.free_vals ? k.free_vals[sid - 1] : 0.0) : (sid < 0 && k.dir_vals ? k.dir_vals[-sid - 1] : 0.0);
    }
    __syncthreads();
    // A. one (cell, point) per thread
    if (tid < nc * NP) {
      const double *xu = sXU + ca * XSTRIDE;
      double Jt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int a = 0; a < ND; a++) {
        const double x0 = xu[a * 3], x1 = xu[a * 3 + 1], x2 = xu[a * 3 + 2];
#pragma unroll
        for (int i = 0; i < 3; i++) {
          Jt[i * 3 + 0] += dn[a][i] * x0;
          Jt[i * 3 + 1] += dn[a][i] * x1;
          Jt[i * 3 + 2] += dn[a][i] * x2;
        }
      }
      double iJ[9];
      const double det = inv3(Jt, iJ);
      const double dv = fabs(det) * wp;
      double g[ND][3];
      double gu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // (grad u)[i][c] = sum_a u_{a,c} d_i phi_a
#pragma unroll
      for (int a = 0; a < ND; a++) {
#pragma unroll
        for (int i = 0; i < 3; i++) g[a][i] = iJ[i * 3 + 0] * dn[a][0] + iJ[i * 3 + 1] * dn[a][1] + iJ[i * 3 + 2] * dn[a][2];
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const double u = xu[24 + a + ND * c];
          gu[0 * 3 + c] += u * g[a][0];
          gu[1 * 3 + c] += u * g[a][1];
          gu[2 * 3 + c] += u * g[a][2];
        }
      }
      double F[9], Fi[9];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) F[i * 3 + j] = (i == j ? 1.0 : 0.0) + gu[j * 3 + i];
      const double detF = inv3(F, Fi);
      const double kap = k.p1 - k.p0 * log(fabs(detF));
      double *q = sQ + ca * CSTRIDE + pa * PSTRIDE;
#pragma unroll
      for (int a = 0; a < ND; a++) {
        // alpha_c = sum_i g_i Y[c][i],  Y[c][i] = Fi[i][c]
        double2 *o = reinterpret_cast<double2 *>(q + a * 6);
        const double al0 = g[a][0] * Fi[0] + g[a][1] * Fi[3] + g[a][2] * Fi[6];
        const double al1 = g[a][0] * Fi[1] + g[a][1] * Fi[4] + g[a][2] * Fi[7];
        const double al2 = g[a][0] * Fi[2] + g[a][1] * Fi[5] + g[a][2] * Fi[8];
        o[0] = make_double2(al0, al1);
        o[1] = make_double2(al2, g[a][0]);
        o[2] = make_double2(g[a][1], g[a][2]);
        if (VEC) {   // residual share of the point: dV grad(phi_a) . P[c][:],  P = mu F - kappa F^-T
          double *r = sK + (ca * NP + pa) * RSTRIDE;
#pragma unroll
          for (int c = 0; c < 3; c++) {
            const double P0 = k.p1 * F[c * 3 + 0] - kap * Fi[0 * 3 + c], P1 = k.p1 * F[c * 3 + 1] - kap * Fi[1 * 3 + c], P2 = k.p1 * F[c * 3 + 2] - kap * Fi[2 * 3 + c];
            r[a + ND * c] = dv * (g[a][0] * P0 + g[a][1] * P1 + g[a][2] * P2);
          }
        }
      }
      q[ND * 6 + 0] = k.p0 * dv;
      q[ND * 6 + 1] = kap * dv;
      q[ND * 6 + 2] = k.p1 * dv;
    }
    __syncthreads();
    if (VEC) {   // residual: points summed in a fixed order, then one add per row
      for (int e = tid; e < nc * 24; e += THREADS) {
        const int c = e / 24, l = e - c * 24;
        const int32_t row = sRow[e];
        if (row <= 0) continue;
        double v = 0.0;
#pragma unroll
        for (int p = 0; p < NP; p++) v += sK[(c * NP + p) * RSTRIDE + l];
        double *dst = k.bvec + (row - 1 + k.row_off);
        if (k.atomic) atomicAdd(dst, v); else *dst += v;
      }
    }
    // B. one 2x2 tile of node pairs per thread
    double K[4][9];
    const bool activeB = tid < nc * TILES;
    if (activeB) {
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 9; i++) K[r][i] = 0.0;
      const double *qc = sQ + cb * CSTRIDE;
#pragma unroll 1
      for (int p = 0; p < NP; p++) {
        const double *q = qc + p * PSTRIDE;
        const double lam = q[ND * 6 + 0], kap = q[ND * 6 + 1], mu = q[ND * 6 + 2];
        double A[2][6], B[2][6];
#pragma unroll
        for (int r = 0; r < 2; r++) {
          const double2 *pa2 = reinterpret_cast<const double2 *>(q + (a0 + r) * 6), *pb2 = reinterpret_cast<const double2 *>(q + (b0 + r) * 6);
          const double2 x0 = pa2[0], x1 = pa2[1], x2 = pa2[2], y0 = pb2[0], y1 = pb2[1], y2 = pb2[2];
          A[r][0] = x0.x; A[r][1] = x0.y; A[r][2] = x1.x; A[r][3] = x1.y; A[r][4] = x2.x; A[r][5] = x2.y;
          B[r][0] = y0.x; B[r][1] = y0.y; B[r][2] = y1.x; B[r][3] = y1.y; B[r][4] = y2.x; B[r][5] = y2.y;
        }
#pragma unroll
        for (int ra = 0; ra < 2; ra++) {
          const double la[3] = {lam * A[ra][0], lam * A[ra][1], lam * A[ra][2]};
          const double ka[3] = {kap * A[ra][0], kap * A[ra][1], kap * A[ra][2]};
          const double ma[3] = {mu * A[ra][3], mu * A[ra][4], mu * A[ra][5]};
#pragma unroll
          for (int rb = 0; rb < 2; rb++) {
            const double gab = ma[0] * B[rb][3] + ma[1] * B[rb][4] + ma[2] * B[rb][5];
            double *Kr = K[ra * 2 + rb];
#pragma unroll
            for (int ci = 0; ci < 3; ci++)
#pragma unroll
              for (int cj = 0; cj < 3; cj++) Kr[ci * 3 + cj] += la[ci] * B[rb][cj] + ka[cj] * B[rb][ci] + (ci == cj ? gab : 0.0);
          }
        }
      }
    }
    __syncthreads();   // the residual partials (same buffer as the staged blocks) are consumed
    if (activeB) {
      double *st = sK + cb * (NPAIR * 9);
#pragma unroll
      for (int ra = 0; ra < 2; ra++)
#pragma unroll
        for (int rb = 0; rb < 2; rb++) {
          const int a = a0 + ra, b = b0 + rb;
          if (a > b) continue;   // (diagonal tiles: the pair (a0 + 1, a0) is the transpose of (a0, a0 + 1))
          double *o = st + (b * (b + 1) / 2 + a) * 9;
#pragma unroll
          for (int i = 0; i < 9; i++) o[i] = K[ra * 2 + rb][i];
        }
    }
    __syncthreads();
    if (!k.cell_list) {
      double *out = k.ke_out + it0 * (NPAIR * 9);
      for (int e = tid; e < nc * NPAIR * 9; e += THREADS) out[e] = sK[e];
    } else {
      for (int e = tid; e < nc * NPAIR * 9; e += THREADS) {
        const int c = e / (NPAIR * 9);
        k.ke_out[cell_of(c) * (NPAIR * 9) + (e - c * (NPAIR * 9))] = sK[e];
      }
    }
  }
}

template <int VEC>
bool launch_nh_q1_staged(gb200_plan plan, VArgs &k) {
  gb200_ctx ctx = plan->ctx;
  static const bool disabled = getenv("GB200_NO_NHQ1") != nullptr;
  if (disabled) return false;
  const size_t smem = (size_t)nhq1::SMEM_DOUBLES * sizeof(double);
  auto kern = nh_q1_staged_kernel<VEC>;
  static std::map<int, int> cps_of_device;
  int &cps = cps_of_device[ctx->device];
  if (cps == 0) {
    GB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cps, kern, nhq1::THREADS, smem));
    cps = std::max(cps, 1);
  }
  auto launch = [&](int64_t begin, int64_t end, const int32_t *list, int atomic) {
    if (end <= begin) return;
    k.cell_begin = begin; k.cell_end = end; k.cell_list = list; k.atomic = atomic;
    const int64_t nblocks = (end - begin + nhq1::CELLS - 1) / nhq1::CELLS;
    const int grid = (int)std::min<int64_t>(nblocks, (int64_t)ctx->num_sms * cps);
    kern<<<grid, nhq1::THREADS, smem, ctx->stream>>>(k);
    check_launch(ctx, "nh_q1_staged_kernel");
  };
  if (VEC != 0 && ctx->deterministic()) {
    for (int c = 0; c < plan->ncolors; c++) launch(plan->color_ptr[c], plan->color_ptr[c + 1], plan->color_cells.p, 0);
  } else {
    launch(0, plan->mesh->ncells, nullptr, 1);
  }
  return true;
}

template <int NN, int NDS, int NP, int CELLS, int THREADS>
bool dispatch_form(gb200_plan plan, int form, int vec, VArgs &k) {
  constexpr int NONE = GB200_FORM_NONE, SRC = GB200_FORM_SOURCE, RES = GB200_FORM_NEOHOOKEAN_RES;
  if (vec == 0) {
    switch (form) {
      case GB200_FORM_MASS: launch_one<GB200_FORM_MASS, NONE, NN, NDS, NP, CELLS, THREADS>(plan, k); return true;
      case GB200_FORM_LAPLACIAN: launch_one<GB200_FORM_LAPLACIAN, NONE, NN, NDS, NP, CELLS, THREADS>(plan, k); return true;
      case GB200_FORM_ELASTICITY: launch_one<GB200_FORM_ELASTICITY, NONE, NN, NDS, NP, CELLS, THREADS>(plan, k); return true;
      case GB200_FORM_NEOHOOKEAN_JAC: launch_one<GB200_FORM_NEOHOOKEAN_JAC, NONE, NN, NDS, NP, CELLS, THREADS>(plan, k); return true;
    }
    return false;
  }
  if (form == 0 && vec == SRC) { launch_one<NONE, SRC, NN, NDS, NP, CELLS, THREADS>(plan, k); return true; }
  if (form == 0 && vec == RES) { launch_one<NONE, RES, NN, NDS, NP, CELLS, THREADS>(plan, k); return true; }
  if (form == GB200_FORM_NEOHOOKEAN_JAC && vec == RES) { launch_one<GB200_FORM_NEOHOOKEAN_JAC, RES, NN, NDS, NP, CELLS, THREADS>(plan, k); return true; }
  return false;
}

}  // namespace

bool vector_kernel_pairs(gb200_plan plan, int &npair) {
  const ElemDesc &ed = plan->ed;
  if (ed.D != 3 || ed.Dr != 3 || plan->nfields != 1 || ed.f[0].ncomp != 3 || ed.f[0].lofs != 0 || ed.lface) return false;
  if (getenv("GB200_NO_VECTOR_KERNEL") != nullptr) return false;
  const int nn = ed.nn, nds = ed.f[0].nds, np = ed.np;
  const bool inst = (nn == 8 && nds == 8 && np == 8) || (nn == 8 && nds == 27 && np == 27) || (nn == 4 && nds == 10 && np == 14) || (nn == 4 && nds == 4 && np == 4);
  npair = nds * (nds + 1) / 2;
  return inst;
}

// Returns false when (element, forms) has no specialised instance: the caller then uses the generic kernel.
// field = 0 always (the vector field must be the first field of the plan); inside a multi-field plan (Stokes) only the
// (field 0, field 0) block is handled here.
bool launch_vector_kernel(gb200_plan plan, int form, int form_vec, const double *params, const double *fq, double *nzval, double *bvec, double *ke_out) {
  const ElemDesc &ed = plan->ed;
  if (ed.D != 3 || ed.Dr != 3 || ed.f[0].ncomp != 3 || ed.f[0].lofs != 0 || ed.lface) return false;
  if (plan->nfields != 1 && (form != GB200_FORM_LAPLACIAN || form_vec != 0)) return false;
  static const bool disabled = getenv("GB200_NO_VECTOR_KERNEL") != nullptr;
  if (disabled) return false;
  VArgs k;
  memset(&k, 0, sizeof(k));
  k.X = ed.X; k.cell_nodes = ed.cell_nodes; k.w = ed.w; k.dNg = ed.dNg; k.N = ed.f[0].N; k.dN = ed.f[0].dN;
  k.row_ids = ed.f[0].row_ids; k.col_ids = ed.f[0].col_ids; k.free_vals = ed.f[0].free_vals; k.dir_vals = ed.f[0].dir_vals;
  k.state_ids = ed.f[0].state_ids;
  k.row_off = ed.f[0].row_off; k.col_off = ed.f[0].col_off;
  k.colptr = plan->colptr.p; k.rank = plan->rank.p; k.nzval = nzval; k.bvec = bvec; k.fq = fq;
  k.p0 = params[0]; k.p1 = params[1];
  k.f0 = params[4]; k.f1 = params[5]; k.f2 = params[6];
  k.nltot = plan->NL;
  k.ke_out = ke_out;
  if (ke_out && plan->nfields != 1) return false;
  if (plan->nfields == 2) {  // Stokes: scalar pressure field after the velocity field
    const FieldDesc &f1 = ed.f[1];
    if (f1.ncomp != 1 || f1.lofs != 3 * ed.f[0].nds) return false;
    k.np1 = f1.nds; k.N1 = f1.N; k.row_ids1 = f1.row_ids; k.col_ids1 = f1.col_ids; k.row_off1 = f1.row_off; k.col_off1 = f1.col_off;
  }
  ScopedTimer timer(plan->ctx, "k:vector");
  const int nn = ed.nn, nds = ed.f[0].nds, np = ed.np;
  // <NN, NDS, NP, cells per CTA, threads>: cells x pairs(a<=b) is a multiple of (or just below one of) the thread count
  if (nn == 8 && nds == 8 && np == 8 && ke_out && form == GB200_FORM_NEOHOOKEAN_JAC && (form_vec == 0 || form_vec == GB200_FORM_NEOHOOKEAN_RES) &&
      plan->test[0]->refel->dN == plan->geo->dN) {   // (geometry map and field basis share their tabulation: both Q1)
    if (form_vec ? launch_nh_q1_staged<1>(plan, k) : launch_nh_q1_staged<0>(plan, k)) {
      plan->path_detail[form] = "nhq1";
      return true;
    }
  }
  if (nn == 8 && nds == 8 && np == 8) return dispatch_form<8, 8, 8, 8, 96>(plan, form, form_vec, k);         // Q1 hex, degree 2: 8 x 36 = 3 x 96
  if (nn == 8 && nds == 27 && np == 27 && form == GB200_FORM_ELASTICITY && form_vec == 0 && plan->nfields == 1 && !ke_out && launch_q2_elasticity_mma(plan, k)) {
    plan->path_detail[form] = "dmma";
    return true;
  }
  if (nn == 8 && nds == 27 && np == 27) return dispatch_form<8, 27, 27, 1, 128>(plan, form, form_vec, k);    // Q2 hex, degree 4: 378 ~ 3 x 128
  if (nn == 4 && nds == 10 && np == 14) return dispatch_form<4, 10, 14, 7, 128>(plan, form, form_vec, k);    // P2 tet, degree 4: 7 x 55 = 385 ~ 3 x 128
  if (nn == 4 && nds == 4 && np == 4) return dispatch_form<4, 4, 4, 32, 128>(plan, form, form_vec, k);       // P1 tet, degree 2: 32 x 10 = 2.5 x 128
  return false;
}

}  // namespace gb
