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
  const int32_t *row_ids, *col_ids;
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

constexpr int NH_STRIDE = 48;  // per quadrature point: Y[9] Z[9] Cinv[9] S[9] kappa, SF[9] = S.F^T rows (residual)

template <int FORM, int VEC, int NN, int NDS, int NP, int TEAM>
__global__ void __launch_bounds__(128) vector_kernel(VArgs k) {
  constexpr bool NEED_NH = (FORM == GB200_FORM_NEOHOOKEAN_JAC) || (VEC == GB200_FORM_NEOHOOKEAN_RES);
  constexpr int NL = 3 * NDS;
  constexpr int TEAMS = 128 / TEAM;
  extern __shared__ double smem[];
  constexpr int SCRATCH = NP * NDS * 3 + NP * 10 + (NEED_NH ? NP * NH_STRIDE : 0) + NL + 2;
  const int team = threadIdx.x / TEAM, tid = threadIdx.x % TEAM;
  double *sG = smem + (size_t)team * SCRATCH;      // [NP][NDS][3] physical gradients
  double *siJ = sG + NP * NDS * 3;                 // [NP][9]
  double *sdV = siJ + NP * 9;                      // [NP]
  double *sNH = sdV + NP;                          // [NP][NH_STRIDE]
  int32_t *sRow = reinterpret_cast<int32_t *>(sNH + (NEED_NH ? NP * NH_STRIDE : 0));
  int32_t *sCol = sRow + NL;

  for (int64_t it = k.cell_begin + (int64_t)blockIdx.x * TEAMS + team; it < k.cell_end; it += (int64_t)gridDim.x * TEAMS) {
    const int64_t cell = k.cell_list ? k.cell_list[it] : it;
    for (int l = tid; l < NL; l += TEAM) {
      sRow[l] = k.row_ids[cell * NL + l];
      sCol[l] = k.col_ids[cell * NL + l];
    }
    // 1. geometry at the quadrature points
    for (int p = tid; p < NP; p += TEAM) {
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
      double det = inv3(Jt, siJ + p * 9);
      sdV[p] = fabs(det) * k.w[p];
    }
    if (TEAM == 32) __syncwarp(); else __syncthreads();
    // 2. physical gradients
    for (int e = tid; e < NP * NDS; e += TEAM) {
      const int p = e / NDS;
      const double *dn = k.dN + e * 3;
      const double *iJ = siJ + p * 9;
      const double d0 = dn[0], d1 = dn[1], d2 = dn[2];
      sG[e * 3 + 0] = iJ[0] * d0 + iJ[1] * d1 + iJ[2] * d2;
      sG[e * 3 + 1] = iJ[3] * d0 + iJ[4] * d1 + iJ[5] * d2;
      sG[e * 3 + 2] = iJ[6] * d0 + iJ[7] * d1 + iJ[8] * d2;
    }
    if (TEAM == 32) __syncwarp(); else __syncthreads();
    // 3. neo-Hookean state per quadrature point
    if (NEED_NH) {
      for (int p = tid; p < NP; p += TEAM) {
        double gu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // (grad u)[i][c] = sum_a u_{a,c} d_i N_a
        for (int c = 0; c < 3; c++)
          for (int a = 0; a < NDS; a++) {
            const int32_t id = sCol[a + NDS * c];
            const double u = id > 0 ? (k.free_vals ? k.free_vals[id - 1] : 0.0) : (id < 0 && k.dir_vals ? k.dir_vals[-id - 1] : 0.0);
            const double *g = sG + (p * NDS + a) * 3;
            gu[0 * 3 + c] += u * g[0];
            gu[1 * 3 + c] += u * g[1];
            gu[2 * 3 + c] += u * g[2];
          }
        double F[9], C[9], Ci[9];
        for (int i = 0; i < 3; i++)
          for (int j = 0; j < 3; j++) F[i * 3 + j] = (i == j ? 1.0 : 0.0) + gu[j * 3 + i];
        for (int i = 0; i < 3; i++)
          for (int j = 0; j < 3; j++) C[i * 3 + j] = F[0 * 3 + i] * F[0 * 3 + j] + F[1 * 3 + i] * F[1 * 3 + j] + F[2 * 3 + i] * F[2 * 3 + j];
        const double detC = inv3(C, Ci);
        const double lnJ = log(sqrt(detC));
        double *o = sNH + p * NH_STRIDE;
        for (int c = 0; c < 3; c++)      // Y[c][:] = Cinv . F[c,:]
          for (int i = 0; i < 3; i++) o[c * 3 + i] = Ci[i * 3 + 0] * F[c * 3 + 0] + Ci[i * 3 + 1] * F[c * 3 + 1] + Ci[i * 3 + 2] * F[c * 3 + 2];
        for (int c = 0; c < 3; c++)      // Z[c][d] = F[c,:] . Y[d][:]
          for (int d = 0; d < 3; d++) o[9 + c * 3 + d] = F[c * 3 + 0] * o[d * 3 + 0] + F[c * 3 + 1] * o[d * 3 + 1] + F[c * 3 + 2] * o[d * 3 + 2];
        for (int i = 0; i < 9; i++) o[18 + i] = Ci[i];
        for (int i = 0; i < 3; i++)
          for (int j = 0; j < 3; j++) o[27 + i * 3 + j] = k.p1 * ((i == j ? 1.0 : 0.0) - Ci[i * 3 + j]) + k.p0 * lnJ * Ci[i * 3 + j];
        o[36] = k.p1 - k.p0 * lnJ;
        for (int c = 0; c < 3; c++)      // SF[c][i] = sum_m S[i][m] F[c][m]:  dE(grad v):S = ga . SF[ci]
          for (int i = 0; i < 3; i++) o[37 + c * 3 + i] = o[27 + i * 3 + 0] * F[c * 3 + 0] + o[27 + i * 3 + 1] * F[c * 3 + 1] + o[27 + i * 3 + 2] * F[c * 3 + 2];
      }
      if (TEAM == 32) __syncwarp(); else __syncthreads();
    }
    // 4. node pairs: 3x3 component blocks
    const int NLT = k.nltot;
    const uint16_t *rk = k.rank + cell * (int64_t)NLT * NLT;
    if (FORM != GB200_FORM_NONE)
    for (int pair = tid; pair < NDS * NDS; pair += TEAM) {
      const int b = pair / NDS, a = pair - b * NDS;  // a: test node (row), b: trial node (column)
      double K[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};     // K[ci*3+cj]
      for (int p = 0; p < NP; p++) {
        const double dv = sdV[p];
        if (FORM == GB200_FORM_MASS) {
          const double v = k.N[p * NDS + a] * k.N[p * NDS + b] * dv;
          K[0] += v; K[4] += v; K[8] += v;
          continue;
        }
        const double *ga = sG + (p * NDS + a) * 3, *gb = sG + (p * NDS + b) * 3;
        const double a0 = ga[0], a1 = ga[1], a2 = ga[2], b0 = gb[0], b1 = gb[1], b2 = gb[2];
        if (FORM == GB200_FORM_LAPLACIAN) {
          const double v = (a0 * b0 + a1 * b1 + a2 * b2) * dv;
          K[0] += v; K[4] += v; K[8] += v;
        } else if (FORM == GB200_FORM_ELASTICITY) {
          const double l0 = k.p0 * dv, m0 = k.p1 * dv;
          const double s = m0 * (a0 * b0 + a1 * b1 + a2 * b2);
          const double av[3] = {a0, a1, a2}, bv[3] = {b0, b1, b2};
#pragma unroll
          for (int ci = 0; ci < 3; ci++)
#pragma unroll
            for (int cj = 0; cj < 3; cj++) K[ci * 3 + cj] += l0 * av[ci] * bv[cj] + m0 * av[cj] * bv[ci] + (ci == cj ? s : 0.0);
        } else {  // neo-Hookean Jacobian
          const double *o = sNH + p * NH_STRIDE;
          const double *Y = o, *Z = o + 9, *Ci = o + 18, *S = o + 27;
          const double kap = o[36];
          double al[3], be[3];
#pragma unroll
          for (int c = 0; c < 3; c++) {
            al[c] = a0 * Y[c * 3 + 0] + a1 * Y[c * 3 + 1] + a2 * Y[c * 3 + 2];
            be[c] = b0 * Y[c * 3 + 0] + b1 * Y[c * 3 + 1] + b2 * Y[c * 3 + 2];
          }
          const double cab = a0 * (Ci[0] * b0 + Ci[1] * b1 + Ci[2] * b2) + a1 * (Ci[3] * b0 + Ci[4] * b1 + Ci[5] * b2) + a2 * (Ci[6] * b0 + Ci[7] * b1 + Ci[8] * b2);
          const double sab = a0 * (S[0] * b0 + S[1] * b1 + S[2] * b2) + a1 * (S[3] * b0 + S[4] * b1 + S[5] * b2) + a2 * (S[6] * b0 + S[7] * b1 + S[8] * b2);
#pragma unroll
          for (int ci = 0; ci < 3; ci++)
#pragma unroll
            for (int cj = 0; cj < 3; cj++)
              K[ci * 3 + cj] += dv * (k.p0 * be[cj] * al[ci] + kap * (cab * Z[cj * 3 + ci] + al[cj] * be[ci]) + (ci == cj ? sab : 0.0));
        }
      }
      const double coef = (FORM == GB200_FORM_MASS || FORM == GB200_FORM_LAPLACIAN) ? k.p0 : 1.0;
#pragma unroll
      for (int cj = 0; cj < 3; cj++) {
        const int lj = b + NDS * cj;
        const int32_t col = sCol[lj];
        if (col <= 0) continue;
        const int64_t base = k.colptr[col - 1 + k.col_off];
#pragma unroll
        for (int ci = 0; ci < 3; ci++) {
          const int li = a + NDS * ci;
          if (sRow[li] <= 0) continue;
          double *dst = k.nzval + base + rk[li + NLT * lj];
          const double v = coef * K[ci * 3 + cj];
          if (k.atomic) atomicAdd(dst, v); else *dst += v;
        }
      }
    }
    // 4b. Stokes coupling blocks: T[c] = sum_p d_c N_a psi_b dV ;  (v,p) entry = -T, (q,u) entry = +T  (StokesTaylorHoodTests.jl:59)
    if (FORM == GB200_FORM_LAPLACIAN && VEC == 0 && k.np1 > 0) {
      const int np1 = k.np1;
      for (int pr = tid; pr < NDS * np1; pr += TEAM) {
        const int b = pr / NDS, a = pr - b * NDS;
        double T0 = 0.0, T1 = 0.0, T2 = 0.0;
        for (int p = 0; p < NP; p++) {
          const double wv = k.N1[p * np1 + b] * sdV[p];
          const double *ga = sG + (p * NDS + a) * 3;
          T0 += ga[0] * wv; T1 += ga[1] * wv; T2 += ga[2] * wv;
        }
        const double T[3] = {T0, T1, T2};
        const int32_t prow = k.row_ids1[cell * np1 + b], pcol = k.col_ids1[cell * np1 + b];
        const int lp = NL + b;  // local index of the pressure dof in the concatenated numbering
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const int lv = a + NDS * c;
          if (pcol > 0 && sRow[lv] > 0) {  // (v,p): row = velocity test dof, column = pressure trial dof
            double *dst = k.nzval + k.colptr[pcol - 1 + k.col_off1] + rk[lv + NLT * lp];
            if (k.atomic) atomicAdd(dst, -T[c]); else *dst -= T[c];
          }
          if (prow > 0 && sCol[lv] > 0) {  // (q,u): row = pressure test dof, column = velocity trial dof
            double *dst = k.nzval + k.colptr[sCol[lv] - 1 + k.col_off] + rk[lp + NLT * lv];
            if (k.atomic) atomicAdd(dst, T[c]); else *dst += T[c];
          }
        }
      }
    }
    // 5. local vector: source term or neo-Hookean residual
    if (VEC != 0) {
      for (int li = tid; li < NL; li += TEAM) {
        const int32_t row = sRow[li];
        if (row <= 0) continue;
        const int ci = li / NDS, a = li - ci * NDS;
        double v = 0.0;
        for (int p = 0; p < NP; p++) {
          if (VEC == GB200_FORM_SOURCE) {
            const double f = k.fq ? k.fq[((int64_t)cell * NP + p) * 3 + ci] : (ci == 0 ? k.f0 : ci == 1 ? k.f1 : k.f2);
            v += k.N[p * NDS + a] * f * sdV[p];
          } else {
            const double *g = sG + (p * NDS + a) * 3;
            const double *sf = sNH + p * NH_STRIDE + 37 + ci * 3;
            v += (g[0] * sf[0] + g[1] * sf[1] + g[2] * sf[2]) * sdV[p];
          }
        }
        double *dst = k.bvec + (row - 1 + k.row_off);
        if (k.atomic) atomicAdd(dst, v); else *dst += v;
      }
    }
    if (TEAM == 32) __syncwarp(); else __syncthreads();
  }
}

template <int FORM, int VEC, int NN, int NDS, int NP, int TEAM>
void launch_one(gb200_plan plan, VArgs &k) {
  gb200_ctx ctx = plan->ctx;
  constexpr int NL = 3 * NDS;
  constexpr bool NEED_NH = (FORM == GB200_FORM_NEOHOOKEAN_JAC) || (VEC == GB200_FORM_NEOHOOKEAN_RES);
  constexpr int SCRATCH = NP * NDS * 3 + NP * 10 + (NEED_NH ? NP * NH_STRIDE : 0) + NL + 2;
  constexpr int TEAMS = 128 / TEAM;
  const size_t smem = (size_t)TEAMS * SCRATCH * sizeof(double);
  auto kern = vector_kernel<FORM, VEC, NN, NDS, NP, TEAM>;
  if (smem > 48 * 1024) GB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  auto launch = [&](int64_t begin, int64_t end, const int32_t *list, int atomic) {
    if (end <= begin) return;
    k.cell_begin = begin; k.cell_end = end; k.cell_list = list; k.atomic = atomic;
    int64_t nblocks = (end - begin + TEAMS - 1) / TEAMS;
    int grid = (int)std::min<int64_t>(nblocks, (int64_t)ctx->num_sms * 8);
    kern<<<grid, 128, smem, ctx->stream>>>(k);
    check_launch(ctx, "vector_kernel");
  };
  if (ctx->deterministic()) {
    for (int c = 0; c < plan->ncolors; c++) launch(plan->color_ptr[c], plan->color_ptr[c + 1], plan->color_cells.p, 0);
  } else {
    launch(0, plan->mesh->ncells, nullptr, 1);
  }
}

template <int NN, int NDS, int NP, int TEAM>
bool dispatch_form(gb200_plan plan, int form, int vec, VArgs &k) {
  constexpr int NONE = GB200_FORM_NONE, SRC = GB200_FORM_SOURCE, RES = GB200_FORM_NEOHOOKEAN_RES;
  if (vec == 0) {
    switch (form) {
      case GB200_FORM_MASS: launch_one<GB200_FORM_MASS, NONE, NN, NDS, NP, TEAM>(plan, k); return true;
      case GB200_FORM_LAPLACIAN: launch_one<GB200_FORM_LAPLACIAN, NONE, NN, NDS, NP, TEAM>(plan, k); return true;
      case GB200_FORM_ELASTICITY: launch_one<GB200_FORM_ELASTICITY, NONE, NN, NDS, NP, TEAM>(plan, k); return true;
      case GB200_FORM_NEOHOOKEAN_JAC: launch_one<GB200_FORM_NEOHOOKEAN_JAC, NONE, NN, NDS, NP, TEAM>(plan, k); return true;
    }
    return false;
  }
  if (form == 0 && vec == SRC) { launch_one<NONE, SRC, NN, NDS, NP, TEAM>(plan, k); return true; }
  if (form == 0 && vec == RES) { launch_one<NONE, RES, NN, NDS, NP, TEAM>(plan, k); return true; }
  if (form == GB200_FORM_NEOHOOKEAN_JAC && vec == RES) { launch_one<GB200_FORM_NEOHOOKEAN_JAC, RES, NN, NDS, NP, TEAM>(plan, k); return true; }
  return false;
}

}  // namespace

// Returns false when (element, forms) has no specialised instance: the caller then uses the generic kernel.
// field = 0 always (the vector field must be the first field of the plan); inside a multi-field plan (Stokes) only the
// (field 0, field 0) block is handled here.
bool launch_vector_kernel(gb200_plan plan, int form, int form_vec, const double *params, const double *fq, double *nzval, double *bvec) {
  const ElemDesc &ed = plan->ed;
  if (ed.D != 3 || ed.f[0].ncomp != 3 || ed.f[0].lofs != 0) return false;
  if (plan->nfields != 1 && (form != GB200_FORM_LAPLACIAN || form_vec != 0)) return false;
  static const bool disabled = getenv("GB200_NO_VECTOR_KERNEL") != nullptr;
  if (disabled) return false;
  VArgs k;
  memset(&k, 0, sizeof(k));
  k.X = ed.X; k.cell_nodes = ed.cell_nodes; k.w = ed.w; k.dNg = ed.dNg; k.N = ed.f[0].N; k.dN = ed.f[0].dN;
  k.row_ids = ed.f[0].row_ids; k.col_ids = ed.f[0].col_ids; k.free_vals = ed.f[0].free_vals; k.dir_vals = ed.f[0].dir_vals;
  k.row_off = ed.f[0].row_off; k.col_off = ed.f[0].col_off;
  k.colptr = plan->colptr.p; k.rank = plan->rank.p; k.nzval = nzval; k.bvec = bvec; k.fq = fq;
  k.p0 = params[0]; k.p1 = params[1];
  k.f0 = params[4]; k.f1 = params[5]; k.f2 = params[6];
  k.nltot = plan->NL;
  if (plan->nfields == 2) {  // Stokes: scalar pressure field after the velocity field
    const FieldDesc &f1 = ed.f[1];
    if (f1.ncomp != 1 || f1.lofs != 3 * ed.f[0].nds) return false;
    k.np1 = f1.nds; k.N1 = f1.N; k.row_ids1 = f1.row_ids; k.col_ids1 = f1.col_ids; k.row_off1 = f1.row_off; k.col_off1 = f1.col_off;
  }
  ScopedTimer timer(plan->ctx, "k:vector");
  const int nn = ed.nn, nds = ed.f[0].nds, np = ed.np;
  if (nn == 8 && nds == 8 && np == 8) return dispatch_form<8, 8, 8, 32>(plan, form, form_vec, k);       // Q1 hex, degree 2
  if (nn == 8 && nds == 27 && np == 27) return dispatch_form<8, 27, 27, 128>(plan, form, form_vec, k);  // Q2 hex, degree 4
  if (nn == 4 && nds == 10 && np == 14) return dispatch_form<4, 10, 14, 32>(plan, form, form_vec, k);   // P2 tet, degree 4
  if (nn == 4 && nds == 4 && np == 4) return dispatch_form<4, 4, 4, 32>(plan, form, form_vec, k);       // P1 tet, degree 2
  return false;
}

}  // namespace gb
