// element_kernels.cu -- generic cell-centric quadrature + scatter kernels (any supported form / element).
//
// One team of threads (a warp for small elements, a CTA for large ones) owns a cell at a time:
//   1. per quadrature point: Jt = sum_a dNg_a (x) x_a, inv(Jt), dV = |det Jt| w   (reference: a4/a5,
//      src/Fields/FieldArrays.jl:342-376, src/TensorValues/Operations.jl:875-934,989)
//   2. physical gradients grad(phi_a) = inv(Jt) . dN_a                               (a6, ApplyOptimizations.jl:306-310)
//   3. per entry (li,lj): K_e = sum_p integrand * dV_p                               (a7/a8, FieldsInterfaces.jl:737-760)
//   4. scatter through the precomputed slot map (colptr[col] + rank), skipping ids <= 0 (a13)
//   5. local vector, Dirichlet lifting b_e -= K_e u_e, scatter into b                 (a9/a16)
// Scatter is RED.ADD.F64 (atomic mode) or a plain read-modify-write over one colour of cells at a time
// (deterministic mode; cells of a colour share no DoF).
#include "common.cuh"

namespace gb {

namespace {

struct KArgs {
  ElemDesc ed;
  int form_mat, form_vec;
  double params[8];
  int lift;
  const double *fq;
  const double *Ke_const;
  int skip00;
  const int64_t *colptr;
  const uint16_t *rank;
  double *nzval;
  double *bvec;
  const int32_t *cell_list;  // nullptr => identity
  int64_t cell_begin, cell_end;
  int atomic;
  // scratch layout (in doubles) per team
  int o_iJt, o_dV, o_G, o_nh, o_K, o_ue, o_ids, o_nrm, scratch_doubles;
};

__device__ __forceinline__ double det3(const double *a) {
  return a[0] * a[4] * a[8] + a[1] * a[5] * a[6] + a[2] * a[3] * a[7] - (a[0] * a[5] * a[7] + a[1] * a[3] * a[8] + a[2] * a[4] * a[6]);
}

// a[i*D+j]; writes r = inv(a), returns det(a)
__device__ __forceinline__ double inv_det(int D, const double *a, double *r) {
  if (D == 2) {
    double det = a[0] * a[3] - a[1] * a[2];
    double c = 1.0 / det;
    r[0] = a[3] * c; r[1] = -a[1] * c; r[2] = -a[2] * c; r[3] = a[0] * c;
    return det;
  }
  double det = det3(a);
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

template <int TEAM>
__device__ __forceinline__ void team_sync() {
  if (TEAM == 32) __syncwarp(); else __syncthreads();
}

// neo-Hookean state at a quadrature point, stored as 28 doubles: F[9], Cinv[9], S[9], lnJ
__device__ void nh_point(int D, const double *gu, double lambda, double mu, double *o) {
  double F[9], C[9];
  for (int i = 0; i < D; i++)
    for (int j = 0; j < D; j++) F[i * D + j] = (i == j ? 1.0 : 0.0) + gu[j * D + i];  // F = I + (grad u)^T
  for (int i = 0; i < D; i++)
    for (int j = 0; j < D; j++) {
      double s = 0;
      for (int k = 0; k < D; k++) s += F[k * D + i] * F[k * D + j];  // C = F^T F
      C[i * D + j] = s;
    }
  double Cinv[9];
  double detC = inv_det(D, C, Cinv);
  double lnJ = log(sqrt(detC));
  for (int i = 0; i < D * D; i++) { o[i] = F[i]; o[9 + i] = Cinv[i]; }
  for (int i = 0; i < D; i++)
    for (int j = 0; j < D; j++) o[18 + i * D + j] = mu * ((i == j ? 1.0 : 0.0) - Cinv[i * D + j]) + lambda * lnJ * Cinv[i * D + j];
  o[27] = lnJ;
}

// integrand of the (test a,ci | trial b,cj) pair at one quadrature point
__device__ __forceinline__ double mat_integrand(int form, int D, int bi, int bj, int ci, int cj, double Na, double Nb,
                                                const double *ga, const double *gb, const double *prm, const double *nh) {
  switch (form) {
    case GB200_FORM_MASS: return ci == cj ? prm[0] * Na * Nb : 0.0;
    case GB200_FORM_FACET: {   // coef T(v) U(u), T / U = value or normal derivative (nh = the unit normal at this point)
      if (ci != cj) return 0.0;
      double T = Na, U = Nb;
      if ((int)prm[1] == 1) { T = 0; for (int d = 0; d < D; d++) T += nh[d] * ga[d]; }
      if ((int)prm[2] == 1) { U = 0; for (int d = 0; d < D; d++) U += nh[d] * gb[d]; }
      return prm[0] * T * U;
    }
    case GB200_FORM_SKELETON: {   // coef [w(side of v) T(v)] [z(side of u) U(u)]; bi / bj = 0 plus, 1 minus; nh = the PLUS normal
      if (ci != cj) return 0.0;
      double T = Na, U = Nb;
      if ((int)prm[1] == 1) { T = 0; for (int d = 0; d < D; d++) T += nh[d] * ga[d]; }
      if ((int)prm[4] == 1) { U = 0; for (int d = 0; d < D; d++) U += nh[d] * gb[d]; }
      return prm[0] * prm[2 + bi] * T * prm[5 + bj] * U;
    }
    case GB200_FORM_LAPLACIAN: {
      if (ci != cj) return 0.0;
      double s = 0;
      for (int d = 0; d < D; d++) s += ga[d] * gb[d];
      return prm[0] * s;
    }
    case GB200_FORM_ELASTICITY: {
      double s = 0;
      for (int d = 0; d < D; d++) s += ga[d] * gb[d];
      return prm[0] * ga[ci] * gb[cj] + prm[1] * ((ci == cj ? s : 0.0) + ga[cj] * gb[ci]);
    }
    case GB200_FORM_STOKES: {
      if (bi == 0 && bj == 0) {
        if (ci != cj) return 0.0;
        double s = 0;
        for (int d = 0; d < D; d++) s += ga[d] * gb[d];
        return s;
      }
      if (bi == 0 && bj == 1) return -ga[ci] * Nb;
      if (bi == 1 && bj == 0) return Na * gb[cj];
      return 0.0;
    }
    case GB200_FORM_NEOHOOKEAN_JAC: {
      const double *F = nh, *Ci = nh + 9, *S = nh + 18;
      double lnJ = nh[27], lambda = prm[0], mu = prm[1];
      // dE(grad w) with grad w = g (x) e_c :  1/2 ( g (x) F[c,:] + F[c,:] (x) g )
      double Ev[9], Eu[9];
      for (int i = 0; i < D; i++)
        for (int k = 0; k < D; k++) {
          Ev[i * D + k] = 0.5 * (ga[i] * F[ci * D + k] + F[ci * D + i] * ga[k]);
          Eu[i * D + k] = 0.5 * (gb[i] * F[cj * D + k] + F[cj * D + i] * gb[k]);
        }
      double cd = 0;
      for (int i = 0; i < D * D; i++) cd += Ci[i] * Eu[i];
      // T = Cinv . Eu . Cinv^T
      double T1[9], T2[9];
      for (int i = 0; i < D; i++)
        for (int k = 0; k < D; k++) { double s = 0; for (int m = 0; m < D; m++) s += Ci[i * D + m] * Eu[m * D + k]; T1[i * D + k] = s; }
      for (int i = 0; i < D; i++)
        for (int k = 0; k < D; k++) { double s = 0; for (int m = 0; m < D; m++) s += T1[i * D + m] * Ci[k * D + m]; T2[i * D + k] = s; }
      double t1 = 0;
      for (int i = 0; i < D * D; i++) t1 += Ev[i] * (lambda * cd * Ci[i] + 2.0 * (mu - lambda * lnJ) * T2[i]);
      double t2 = 0;
      if (ci == cj)
        for (int i = 0; i < D; i++) { double s = 0; for (int m = 0; m < D; m++) s += S[i * D + m] * gb[m]; t2 += ga[i] * s; }
      return t1 + t2;
    }
  }
  return 0.0;
}

template <int TEAM>
__global__ void __launch_bounds__(TEAM == 32 ? 128 : TEAM) generic_kernel(KArgs k) {
  extern __shared__ double smem[];
  const ElemDesc &ed = k.ed;
  const int D = ed.D, np = ed.np, NL = ed.NL;
  const int teams_per_block = blockDim.x / TEAM;
  const int team_in_block = threadIdx.x / TEAM;
  const int tid = threadIdx.x % TEAM;
  double *sc = smem + (size_t)team_in_block * k.scratch_doubles;
  double *s_iJt = sc + k.o_iJt, *s_dV = sc + k.o_dV, *s_G = sc + k.o_G, *s_nh = sc + k.o_nh, *s_K = sc + k.o_K, *s_ue = sc + k.o_ue;
  double *s_nrm = sc + k.o_nrm;   // unit normals at the facet points (facet-of-cell plans)
  int32_t *s_rows = reinterpret_cast<int32_t *>(sc + k.o_ids);
  int32_t *s_cols = s_rows + NL;
  const bool need_quad = (k.Ke_const == nullptr);

  for (int64_t it = k.cell_begin + (int64_t)blockIdx.x * teams_per_block + team_in_block; it < k.cell_end;
       it += (int64_t)gridDim.x * teams_per_block) {
    const int64_t cell = k.cell_list ? k.cell_list[it] : it;
    const int lf = ed.lface ? ed.lface[cell] : 0;
    const int p0 = lf * np;   // first point of this local face's block of the tabulations (0 on ordinary cells)
    // ids, Dirichlet values
    bool any_dir = false;
    for (int l = tid; l < NL; l += TEAM) {
      int f = (ed.nfields > 1 && l >= ed.f[1].lofs) ? 1 : 0;
      const FieldDesc &fd = ed.f[f];
      int kk = l - fd.lofs;
      int32_t r = fd.row_ids[cell * fd.nld + kk], c = fd.col_ids[cell * fd.nld + kk];
      s_rows[l] = r;
      s_cols[l] = c;
      double ue = 0.0;
      if (c < 0 && fd.dir_vals) ue = fd.dir_vals[-c - 1];
      s_ue[l] = ue;
    }
    if (need_quad) {
      // 1. geometry per quadrature point
      const int Dr = ed.Dr;
      for (int p = tid; p < np; p += TEAM) {
        double Jt[9];
        for (int i = 0; i < 9; i++) Jt[i] = 0.0;
        for (int a = 0; a < ed.nn; a++) {
          const double *x = ed.X + (int64_t)ed.cell_nodes[cell * ed.nn + a] * D;
          const double *dn = ed.dNg + ((int64_t)(p0 + p) * ed.nn + a) * Dr;
          for (int i = 0; i < Dr; i++)
            for (int j = 0; j < D; j++) Jt[i * D + j] += dn[i] * x[j];
        }
        if (Dr == D && ed.lface) {
          // facet of a cell: n = invJt . nref / |invJt . nref| (push_normal, src/Geometry/BoundaryTriangulations.jl:310-318);
          // surface measure of the facet map = |det Jt| |invJt . nref| (nref carries the ratio of the reference measures)
          double *iJ = s_iJt + p * 9;
          double det = inv_det(D, Jt, iJ);
          double v[3] = {0, 0, 0}, m = 0.0;
          for (int i = 0; i < D; i++) {
            for (int q = 0; q < D; q++) v[i] += iJ[i * D + q] * ed.nref[lf * D + q];
            m += v[i] * v[i];
          }
          m = sqrt(m);
          for (int i = 0; i < D; i++) s_nrm[p * 3 + i] = v[i] / m;
          s_dV[p] = fabs(det) * m * ed.w[p0 + p];
          if (ed.skel) {
            // minus cell of an interior facet: inv(Jt) at the point of ITS local-face block that coincides with plus point p
            // (normal and measure are those of the plus side: n- = -n+, same facet)
            const int pm = ed.lface2[cell] * np + ed.perm[cell * np + p];
            for (int i = 0; i < 9; i++) Jt[i] = 0.0;
            for (int a = 0; a < ed.nn; a++) {
              const double *x = ed.X2 + (int64_t)ed.cell_nodes2[cell * ed.nn + a] * D;
              const double *dn = ed.dNg + ((int64_t)pm * ed.nn + a) * Dr;
              for (int i = 0; i < Dr; i++)
                for (int j = 0; j < D; j++) Jt[i * D + j] += dn[i] * x[j];
            }
            inv_det(D, Jt, s_iJt + (np + p) * 9);
          }
        } else if (Dr == D) {
          double det = inv_det(D, Jt, s_iJt + p * 9);
          s_dV[p] = fabs(det) * ed.w[p];
        } else {
          // boundary facets: meas(Jt) = sqrt(det(Jt . J)) (src/TensorValues/Operations.jl:991-1007); no inverse, no gradients
          double m2;
          if (Dr == 1) {
            m2 = 0.0;
            for (int j = 0; j < D; j++) m2 += Jt[j] * Jt[j];
          } else {  // Dr == 2, D == 3: |t1 x t2|^2
            const double n1 = Jt[1] * Jt[5] - Jt[2] * Jt[4], n2 = Jt[2] * Jt[3] - Jt[0] * Jt[5], n3 = Jt[0] * Jt[4] - Jt[1] * Jt[3];
            m2 = n1 * n1 + n2 * n2 + n3 * n3;
          }
          s_dV[p] = sqrt(m2) * ed.w[p];
        }
      }
      team_sync<TEAM>();
      // 2. physical gradients of every field
      if (Dr == D)
      for (int f = 0; f < ed.nfields; f++) {
        const FieldDesc &fd = ed.f[f];
        double *G = s_G + fd.tab_ofs;
        for (int e = tid; e < np * fd.nds; e += TEAM) {
          int p = e / fd.nds;
          const double *dn = fd.dN + ((int64_t)p0 * fd.nds + e) * D;
          const double *iJ = s_iJt + p * 9;
          if (ed.skel && f == 1) {   // minus side: its own local-face block, permuted point, its own inverse Jacobian
            dn = fd.dN + ((int64_t)(ed.lface2[cell] * np + ed.perm[cell * np + p]) * fd.nds + (e - p * fd.nds)) * D;
            iJ = s_iJt + (np + p) * 9;
          }
          for (int i = 0; i < D; i++) {
            double s = 0;
            for (int m = 0; m < D; m++) s += iJ[i * D + m] * dn[m];
            G[e * D + i] = s;
          }
        }
      }
      team_sync<TEAM>();
      // 3. state at the quadrature points (neo-Hookean)
      if (k.form_mat == GB200_FORM_NEOHOOKEAN_JAC || k.form_vec == GB200_FORM_NEOHOOKEAN_RES) {
        const FieldDesc &fd = ed.f[0];
        const double *G = s_G + fd.tab_ofs;
        for (int p = tid; p < np; p += TEAM) {
          double gu[9];
          for (int i = 0; i < D * D; i++) gu[i] = 0.0;
          for (int c = 0; c < fd.ncomp; c++)
            for (int a = 0; a < fd.nds; a++) {
              const int32_t id = fd.state_ids[cell * (int64_t)fd.nld + a + fd.nds * c];
              double u = id > 0 ? (fd.free_vals ? fd.free_vals[id - 1] : 0.0) : (id < 0 && fd.dir_vals ? fd.dir_vals[-id - 1] : 0.0);
              const double *ga = G + ((int64_t)p * fd.nds + a) * D;
              for (int i = 0; i < D; i++) gu[i * D + c] += u * ga[i];
            }
          nh_point(D, gu, k.params[0], k.params[1], s_nh + p * 28);
        }
      }
    }
    team_sync<TEAM>();
    for (int l = 0; l < NL; l++) any_dir |= (s_cols[l] < 0);
    const bool lift = k.lift && any_dir && k.form_mat && k.bvec;

    // 4. matrix entries
    if (k.form_mat || k.Ke_const) {
      const uint16_t *rk = k.rank + cell * (int64_t)NL * NL;
      for (int e = tid; e < NL * NL; e += TEAM) {
        int lj = e / NL, li = e - lj * NL;
        int bi = (ed.nfields > 1 && li >= ed.f[1].lofs) ? 1 : 0;
        int bj = (ed.nfields > 1 && lj >= ed.f[1].lofs) ? 1 : 0;
        int32_t row = s_rows[li], col = s_cols[lj];
        bool store = ed.touched[bi][bj] && row > 0 && col > 0 && k.nzval && !(k.skip00 && bi == 0 && bj == 0);
        bool for_lift = lift && ed.touched[bi][bj] && row > 0 && col < 0;
        double v = 0.0;
        if (store || for_lift) {
          if (!need_quad) {
            v = k.Ke_const[li + NL * lj];
          } else {
            const FieldDesc &ft = ed.f[bi], &fu = ed.f[bj];
            int ki = li - ft.lofs, kj = lj - fu.lofs;
            int a = ki % ft.nds, ci = ki / ft.nds, b = kj % fu.nds, cj = kj / fu.nds;
            const double *Gt = s_G + ft.tab_ofs, *Gu = s_G + fu.tab_ofs;
            if (ed.skel) {
              const int pm0 = ed.lface2[cell] * np;
              const int32_t *pr = ed.perm + cell * np;
              for (int p = 0; p < np; p++) {
                const int pi = bi ? pm0 + pr[p] : p0 + p, pj = bj ? pm0 + pr[p] : p0 + p;
                v += mat_integrand(k.form_mat, D, bi, bj, ci, cj, ft.N[pi * ft.nds + a], fu.N[pj * fu.nds + b], Gt + (p * ft.nds + a) * D,
                                   Gu + (p * fu.nds + b) * D, k.params, s_nrm + p * 3) * s_dV[p];
              }
            } else
            for (int p = 0; p < np; p++)
              v += mat_integrand(k.form_mat, D, bi, bj, ci, cj, ft.N[(p0 + p) * ft.nds + a], fu.N[(p0 + p) * fu.nds + b],
                                 Gt + (p * ft.nds + a) * D, Gu + (p * fu.nds + b) * D, k.params,
                                 k.form_mat == GB200_FORM_FACET ? s_nrm + p * 3 : s_nh + p * 28) * s_dV[p];
          }
        }
        if (lift) s_K[e] = for_lift ? v : 0.0;
        if (store) {
          int64_t col_g = col - 1 + ed.f[bj].col_off;
          double *dst = k.nzval + k.colptr[col_g] + rk[e];
          if (k.atomic) atomicAdd(dst, v); else *dst += v;
        }
      }
    }
    if (lift) team_sync<TEAM>();
    // 5. local vector (+ lifting) and scatter
    if (k.bvec && (k.form_vec || lift)) {
      for (int li = tid; li < NL; li += TEAM) {
        int32_t row = s_rows[li];
        if (row <= 0) continue;
        int bi = (ed.nfields > 1 && li >= ed.f[1].lofs) ? 1 : 0;
        const FieldDesc &ft = ed.f[bi];
        int ki = li - ft.lofs, a = ki % ft.nds, ci = ki / ft.nds;
        double v = 0.0;
        if (k.form_vec == GB200_FORM_SOURCE) {
          for (int p = 0; p < np; p++) {
            double f = ft.src_fq ? ft.src_fq[((int64_t)cell * np + p) * ft.ncomp + ci] : ft.src[ci];
            v += ft.N[(p0 + p) * ft.nds + a] * f * s_dV[p];
          }
        } else if (k.form_vec == GB200_FORM_FACET_VEC) {
          // coef T(v) d: T = value / normal derivative of the test function, d = g at the points, u_h or n.grad(u_h)
          const double *G = s_G + ft.tab_ofs;
          const int tk = (int)k.params[5], dk = (int)k.params[6];
          for (int p = 0; p < np; p++) {
            const double *nr = s_nrm + p * 3;
            double T = ft.N[(p0 + p) * ft.nds + a];
            if (tk == 1) { T = 0; for (int d = 0; d < D; d++) T += nr[d] * G[(p * ft.nds + a) * D + d]; }
            double dat = 0.0;
            if (dk == 0) {
              dat = ft.src_fq ? ft.src_fq[((int64_t)cell * np + p) * ft.ncomp + ci] : ft.src[ci];
            } else {
              for (int j = 0; j < ft.nds; j++) {
                const int32_t id = ft.state_ids[cell * (int64_t)ft.nld + j + ft.nds * ci];
                const double uj = id > 0 ? (ft.free_vals ? ft.free_vals[id - 1] : 0.0) : (id < 0 && ft.dir_vals ? ft.dir_vals[-id - 1] : 0.0);
                double w = ft.N[(p0 + p) * ft.nds + j];
                if (dk == 2) { w = 0; for (int d = 0; d < D; d++) w += nr[d] * G[(p * ft.nds + j) * D + d]; }
                dat += uj * w;
              }
            }
            v += k.params[4] * T * dat * s_dV[p];
          }
        } else if (k.form_vec == GB200_FORM_NEOHOOKEAN_RES) {
          const double *G = s_G + ft.tab_ofs;
          for (int p = 0; p < np; p++) {
            const double *F = s_nh + p * 28, *S = F + 18;
            const double *ga = G + (p * ft.nds + a) * D;
            double s = 0;  // dE(grad v) : S = ga . (S F[ci,:])
            for (int i = 0; i < D; i++) { double t = 0; for (int m = 0; m < D; m++) t += S[i * D + m] * F[ci * D + m]; s += ga[i] * t; }
            v += s * s_dV[p];
          }
        }
        if (lift)
          for (int lj = 0; lj < NL; lj++) v -= s_K[li + NL * lj] * s_ue[lj];
        double *dst = k.bvec + (row - 1 + ft.row_off);
        if (k.atomic) atomicAdd(dst, v); else *dst += v;
      }
    }
    team_sync<TEAM>();
  }
}

__global__ void quad_points_kernel(ElemDesc ed, double *xq) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t total = ed.ncells * ed.np;
  for (; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t cell = t / ed.np;
    int p = (int)(t % ed.np) + (ed.lface ? ed.lface[cell] * ed.np : 0);
    for (int d = 0; d < ed.D; d++) {
      double s = 0;
      for (int a = 0; a < ed.nn; a++) s += ed.Ng[p * ed.nn + a] * ed.X[(int64_t)ed.cell_nodes[cell * ed.nn + a] * ed.D + d];
      xq[t * ed.D + d] = s;
    }
  }
}

}  // namespace

void launch_generic(gb200_plan plan, const NumericArgs &a, double *nzval, double *bvec) {
  gb200_ctx ctx = plan->ctx;
  KArgs k;
  memset(&k, 0, sizeof(k));
  k.ed = plan->ed;
  k.form_mat = a.form_mat;
  k.form_vec = a.form_vec;
  memcpy(k.params, a.params, sizeof(k.params));
  k.lift = a.lift;
  k.fq = a.fq;
  k.Ke_const = a.Ke_const;
  k.skip00 = a.skip_block00 ? 1 : 0;
  k.colptr = plan->colptr.p;
  k.rank = plan->rank.p;
  k.nzval = nzval;
  k.bvec = bvec;
  const int NL = plan->NL, np = plan->ed.np, D = plan->ed.D;
  // scratch layout
  int o = 0;
  k.o_iJt = o; o += np * 9 * (plan->ed.skel ? 2 : 1);   // (skeleton plans: plus and minus inverse Jacobians)
  k.o_dV = o; o += np;
  k.o_G = o;
  for (int f = 0; f < plan->nfields; f++) o += np * plan->ed.f[f].nds * D;
  k.o_nh = o; o += np * 28;
  k.o_nrm = o; o += np * 3;
  k.o_K = o; o += a.lift ? NL * NL : 0;
  k.o_ue = o; o += NL;
  k.o_ids = o; o += (2 * NL + 1) / 2 + 1;
  k.scratch_doubles = o;

  const bool small = NL * NL <= 256;
  const int team = small ? 32 : 128;
  const int block = 128;
  const int teams_per_block = block / team;
  size_t smem = (size_t)teams_per_block * k.scratch_doubles * sizeof(double);
  GB_REQUIRE(smem <= 200 * 1024, GB200_ERR_UNSUPPORTED, "element too large for the generic kernel (%zu B of shared memory)", smem);
  if (small) GB_CUDA(cudaFuncSetAttribute(generic_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else GB_CUDA(cudaFuncSetAttribute(generic_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

  ScopedTimer timer(ctx, "k:generic");
  auto launch = [&](int64_t begin, int64_t end, const int32_t *list, int atomic) {
    if (end <= begin) return;
    k.cell_begin = begin;
    k.cell_end = end;
    k.cell_list = list;
    k.atomic = atomic;
    int64_t nblocks = (end - begin + teams_per_block - 1) / teams_per_block;
    int grid = (int)std::min<int64_t>(nblocks, (int64_t)ctx->num_sms * 16);
    if (small) generic_kernel<32><<<grid, block, smem, ctx->stream>>>(k);
    else generic_kernel<128><<<grid, block, smem, ctx->stream>>>(k);
    check_launch(ctx, "generic_kernel");
  };
  if (ctx->deterministic()) {
    for (int c = 0; c < plan->ncolors; c++) launch(plan->color_ptr[c], plan->color_ptr[c + 1], plan->color_cells.p, 0);
  } else {
    launch(0, plan->mesh->ncells, nullptr, 1);
  }
}

void launch_quadrature_points(gb200_plan plan, double *xq_dev) {
  gb200_ctx ctx = plan->ctx;
  int64_t total = plan->mesh->ncells * plan->ed.np;
  int grid = (int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)ctx->num_sms * 32));
  quad_points_kernel<<<grid, 256, 0, ctx->stream>>>(plan->ed, xq_dev);
  check_launch(ctx, "quad_points_kernel");
}

}  // namespace gb
