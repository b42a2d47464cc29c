/* TEST INFRASTRUCTURE ONLY -- CPU oracle (plain C) for the Gridap assembly hot path.
 *
 * Restates, loop by loop, the reference algorithm (Gridap.jl v0.20.8, /root/reference):
 *   per-cell values   src/Fields/FieldArrays.jl:342-376 (Jacobian = linear combination of node coords),
 *                     src/TensorValues/Operations.jl:875-934,989 (det / inv / meas closed forms),
 *                     src/Fields/ApplyOptimizations.jl:306-310 (grad phi = inv(Jt) . grad N),
 *                     src/Fields/FieldArrays.jl:675-696 (integrand broadcast aq[p,i,j]),
 *                     src/Fields/FieldsInterfaces.jl:737-776 (IntegrationMap: p innermost / vectors p outer),
 *                     src/CellData/AttachDirichlet.jl:76-84 (b_e -= K_e u_e on Dirichlet cells)
 *   symbolic + scatter src/FESpaces/SparseMatrixAssemblers.jl:174-405,
 *                     src/Algebra/AlgebraInterfaces.jl:145-218 (for j outer, for i inner, skip ids <= 0),
 *                     src/Arrays/AlgebraMaps.jl:173-182 (blocks: for bj, for bi, if touched),
 *                     src/Algebra/SparseMatrixCSC.jl:72-283 (CounterCSC, InserterCSC, create_from_nz, nz_index)
 * Nothing in the product links or calls this file.  Single-threaded like the reference.
 * All ids are 1-based on the interface, exactly as Gridap holds them.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

enum { ORC_MASS = 1, ORC_LAPLACIAN = 2, ORC_ELASTICITY = 3, ORC_STOKES = 4, ORC_NEOHOOKEAN_JAC = 5,
       ORC_SOURCE = 10, ORC_NEOHOOKEAN_RES = 11,
       /* terms on the facets of a BoundaryTriangulation that need the adjacent cell (FaceToCellGlue,
          src/Geometry/BoundaryTriangulations.jl:13-70): unit normal (get_facet_normal, :244-283, push_normal :310-318) and the
          cell basis / its gradient at the facet quadrature points.  params = {coef, test kind, trial | data kind}:
          kind 0 = value, 1 = normal derivative n.grad;  data kind 0 = g at the points (fq), 1 = u_h, 2 = n.grad(u_h) */
       ORC_FACET = 20, ORC_FACET_VEC = 21 };

typedef struct {
  int32_t D, nn, np;
  int64_t ncells, nnodes;
  const double *X;            /* [nnodes][D] */
  const int32_t *cell_nodes;  /* [ncells][nn], 1-based */
  const double *w;            /* [np] */
  const double *Ng;           /* [np][nn] */
  const double *dNg;          /* [np][nn][Dr] */
  int32_t Dr;                 /* dimension of the cell type; 0 or D: bulk cells; D-1: facets of a BoundaryTriangulation */
  /* facet-of-cell glue: the "cells" are the cells adjacent to the facets; np = quadrature points per facet; the tabulations
     (w, Ng, dNg and every field's N, dN) hold nlf blocks of np points: block lf = the facet rule mapped onto local face lf of the
     reference cell (compute_face_to_cell_reference_map).  nref[lf] = reference outward normal scaled by the ratio of the reference
     measures (face of the reference cell / facet reference polytope).  lface == NULL: ordinary cells. */
  const int32_t *lface;       /* [ncells], 0-based local face */
  const double *nref;         /* [nlf][D] */
} orc_geom_t;

typedef struct {
  int32_t nds, ncomp;
  const double *N;            /* [np][nds] */
  const double *dN;           /* [np][nds][D] */
  const int32_t *cell_dofs;   /* [ncells][nds*ncomp], signed, 1-based, offsets included */
  const double *free_values;  /* state (residual / jacobian), indexed by id-1-offset; may be NULL */
  const double *dirichlet_values; /* indexed by -id-1; may be NULL (=> zeros) */
  int64_t offset;             /* multi-field offset already added to positive ids */
  /* per-field source of a multi-field linear form l((v,q)) = int(v.f + q*g) (test/GridapTests/StokesTaylorHoodTests.jl:61):
     values at the quadrature points [ncells][np][ncomp] or a constant [ncomp]; both NULL: the problem-wide fq / params */
  const double *fq;
  const double *src;
} orc_field_t;

#define MAXD 3
#define MAXLD 128   /* max local dofs per field */
#define MAXQ 64

/* ---- small tensors, restating TensorValues/Operations.jl ---- */
static double det_t(int D, const double *a /* a[i*D+j] = a[i,j] */) {
  if (D == 1) return a[0];
  if (D == 2) return a[0] * a[3] - a[1] * a[2];
  double a11 = a[0], a12 = a[1], a13 = a[2], a21 = a[3], a22 = a[4], a23 = a[5], a31 = a[6], a32 = a[7], a33 = a[8];
  return a11 * a22 * a33 + a12 * a23 * a31 + a13 * a21 * a32 - (a11 * a23 * a32 + a12 * a21 * a33 + a13 * a22 * a31);
}

static void inv_t(int D, const double *a, double *r) {
  if (D == 1) { r[0] = 1.0 / a[0]; return; }
  if (D == 2) {
    double c = 1.0 / det_t(2, a);
    /* data (column-major) = (a22 c, -a21 c, -a12 c, a11 c) */
    r[0] = a[3] * c; r[2] = -a[2] * c; r[1] = -a[1] * c; r[3] = a[0] * c;
    return;
  }
  double a11 = a[0], a12 = a[1], a13 = a[2], a21 = a[3], a22 = a[4], a23 = a[5], a31 = a[6], a32 = a[7], a33 = a[8];
  double c = 1.0 / det_t(3, a);
  /* column-major data tuple of Operations.jl:912-934: r[i,j] = data[(j-1)*3+i] */
  double data[9] = { (a22 * a33 - a23 * a32) * c, -(a21 * a33 - a23 * a31) * c, (a21 * a32 - a22 * a31) * c,
                     -(a12 * a33 - a13 * a32) * c, (a11 * a33 - a13 * a31) * c, -(a11 * a32 - a12 * a31) * c,
                     (a12 * a23 - a13 * a22) * c, -(a11 * a23 - a13 * a21) * c, (a11 * a22 - a12 * a21) * c };
  for (int j = 0; j < 3; j++) for (int i = 0; i < 3; i++) r[i * 3 + j] = data[j * 3 + i];
}

/* per-cell geometry at every quadrature point: inv(Jt), dV = |det Jt| w, physical point */
typedef struct { double iJt[MAXQ][9]; double dV[MAXQ]; double xq[MAXQ][MAXD]; double nrm[MAXQ][MAXD]; int p0; } cellgeo_t;

static int geom_dr(const orc_geom_t *g) { return g->Dr > 0 ? g->Dr : g->D; }

/* meas(Jt) for a Dr x D Jacobian, Dr < D (src/TensorValues/Operations.jl:991-1007): |t| for a curve, |t1 x t2| for a surface in 3D */
static double meas_rect(int Dr, int D, const double *Jt) {
  if (Dr == 1) { double s = 0; for (int j = 0; j < D; j++) s += Jt[j] * Jt[j]; return sqrt(s); }
  double n1 = Jt[0 * D + 1] * Jt[1 * D + 2] - Jt[0 * D + 2] * Jt[1 * D + 1];
  double n2 = Jt[0 * D + 2] * Jt[1 * D + 0] - Jt[0 * D + 0] * Jt[1 * D + 2];
  double n3 = Jt[0 * D + 0] * Jt[1 * D + 1] - Jt[0 * D + 1] * Jt[1 * D + 0];
  return sqrt(n1 * n1 + n2 * n2 + n3 * n3);
}

static void cell_geometry(const orc_geom_t *g, int64_t cell, cellgeo_t *cg) {
  int D = g->D, Dr = geom_dr(g);
  const int32_t *nodes = g->cell_nodes + cell * g->nn;
  const int lf = g->lface ? g->lface[cell] : 0;
  const int p0 = lf * g->np;   /* first point of this local face's block of the tabulations */
  cg->p0 = p0;
  for (int p = 0; p < g->np; p++) {
    double Jt[9] = {0};
    for (int d = 0; d < D; d++) cg->xq[p][d] = 0.0;
    for (int a = 0; a < g->nn; a++) {
      const double *x = g->X + (int64_t)(nodes[a] - 1) * D;
      const double *dn = g->dNg + ((int64_t)(p0 + p) * g->nn + a) * Dr;
      for (int i = 0; i < Dr; i++) for (int j = 0; j < D; j++) Jt[i * D + j] += dn[i] * x[j]; /* outer(dN_a, x_a) */
      for (int d = 0; d < D; d++) cg->xq[p][d] += g->Ng[(int64_t)(p0 + p) * g->nn + a] * x[d];
    }
    if (Dr == D && g->lface) {
      /* facet of a cell: n = invJt . nref / |invJt . nref| (push_normal); the surface measure of the facet map equals
         |det Jt| |invJt . nref| (Nanson), nref carrying the ratio of the reference measures */
      inv_t(D, Jt, cg->iJt[p]);
      double v[3] = {0, 0, 0}, m = 0.0;
      for (int i = 0; i < D; i++) { for (int k = 0; k < D; k++) v[i] += cg->iJt[p][i * D + k] * g->nref[lf * D + k]; m += v[i] * v[i]; }
      m = sqrt(m);
      for (int i = 0; i < D; i++) cg->nrm[p][i] = v[i] / m;
      cg->dV[p] = fabs(det_t(D, Jt)) * m * g->w[p0 + p];
    } else if (Dr == D) {
      inv_t(D, Jt, cg->iJt[p]);
      cg->dV[p] = fabs(det_t(D, Jt)) * g->w[p];
    } else {  /* facets: no inverse / physical gradients (only mass and source integrands are evaluated there) */
      for (int i = 0; i < 9; i++) cg->iJt[p][i] = 0.0;
      cg->dV[p] = meas_rect(Dr, D, Jt) * g->w[p];
    }
  }
}

/* physical gradients of the scalar shape functions of a field: G[p][a][i] = sum_k iJt[i,k] dN[p][a][k] */
static void phys_grads(const orc_geom_t *g, const orc_field_t *f, const cellgeo_t *cg, double *G) {
  int D = g->D;
  if (geom_dr(g) != D) { memset(G, 0, sizeof(double) * (size_t)g->np * f->nds * D); return; }
  for (int p = 0; p < g->np; p++)
    for (int a = 0; a < f->nds; a++) {
      const double *dn = f->dN + ((int64_t)(cg->p0 + p) * f->nds + a) * D;
      for (int i = 0; i < D; i++) {
        double s = 0.0;
        for (int k = 0; k < D; k++) s += cg->iJt[p][i * D + k] * dn[k];
        G[((int64_t)p * f->nds + a) * D + i] = s;
      }
    }
}

/* gradient tensor of vector basis function (a,c): grad[i][j] = d_i N_a * delta_{j c} */
static void basis_grad_tensor(int D, const double *ga, int c, double *T) {
  for (int i = 0; i < D * D; i++) T[i] = 0.0;
  for (int i = 0; i < D; i++) T[i * D + c] = ga[i];
}

static double inner_t(int n, const double *a, const double *b) { double s = 0; for (int i = 0; i < n; i++) s += a[i] * b[i]; return s; }
static double trace_t(int D, const double *a) { double s = 0; for (int i = 0; i < D; i++) s += a[i * D + i]; return s; }
static void sym_t(int D, const double *a, double *e) { for (int i = 0; i < D; i++) for (int j = 0; j < D; j++) e[i * D + j] = 0.5 * (a[i * D + j] + a[j * D + i]); }
static void matmul_t(int D, const double *a, const double *b, double *c) {
  for (int i = 0; i < D; i++) for (int j = 0; j < D; j++) { double s = 0; for (int k = 0; k < D; k++) s += a[i * D + k] * b[k * D + j]; c[i * D + j] = s; }
}
static void transpose_t(int D, const double *a, double *b) { for (int i = 0; i < D; i++) for (int j = 0; j < D; j++) b[i * D + j] = a[j * D + i]; }

/* value of the FE function of field f at the local dof k of `cell` (PosNegReindex, UnconstrainedFESpaces.jl:65-75) */
static double dof_value(const orc_field_t *f, int64_t cell, int k) {
  int32_t id = f->cell_dofs[cell * (int64_t)(f->nds * f->ncomp) + k];
  if (id > 0) return f->free_values ? f->free_values[id - 1 - f->offset] : 0.0;
  return f->dirichlet_values ? f->dirichlet_values[-id - 1] : 0.0;
}

/* grad u_h at quadrature point p: (grad u)[i][j] = sum_a sum_c u_{a,c} d_i N_a delta_{jc} */
static void state_gradient(const orc_geom_t *g, const orc_field_t *f, const double *G, int64_t cell, int p, double *gu) {
  int D = g->D;
  for (int i = 0; i < D * D; i++) gu[i] = 0.0;
  for (int c = 0; c < f->ncomp; c++)
    for (int a = 0; a < f->nds; a++) {
      double u = dof_value(f, cell, a + f->nds * c);
      const double *ga = G + ((int64_t)p * f->nds + a) * D;
      for (int i = 0; i < D; i++) gu[i * D + c] += u * ga[i];
    }
}

/* neo-Hookean law (Gridap Tutorials "hyperelasticity"; not in /root/reference -> parity unpinned):
 * F = I + (grad u)^T, C = F^T F, J = sqrt(det C), S = mu (I - C^-1) + lambda ln(J) C^-1,
 * dE(gdu,gu) = 1/2 (gdu.F + (gdu.F)^T),  dS = lambda (C^-1 : dE) C^-1 + 2 (mu - lambda ln J) C^-1 . dE . C^-T */
typedef struct { double F[9], Cinv[9], S[9], lnJ; } nh_state_t;
static void nh_state(int D, const double *gu, double lambda, double mu, nh_state_t *s) {
  double guT[9] = {0}, FT[9] = {0}, C[9] = {0};
  transpose_t(D, gu, guT);
  for (int i = 0; i < D; i++) for (int j = 0; j < D; j++) s->F[i * D + j] = (i == j ? 1.0 : 0.0) + guT[i * D + j];
  transpose_t(D, s->F, FT);
  matmul_t(D, FT, s->F, C);
  double J = sqrt(det_t(D, C));
  s->lnJ = log(J);
  inv_t(D, C, s->Cinv);
  for (int i = 0; i < D; i++) for (int j = 0; j < D; j++)
    s->S[i * D + j] = mu * ((i == j ? 1.0 : 0.0) - s->Cinv[i * D + j]) + lambda * s->lnJ * s->Cinv[i * D + j];
}
static void nh_dE(int D, const double *gdu, const double *F, double *dE) {
  double t[9]; matmul_t(D, gdu, F, t); sym_t(D, t, dE);
}
static void nh_dS(int D, const double *dE, const nh_state_t *s, double lambda, double mu, double *dS) {
  double CinvT[9], t1[9], t2[9];
  transpose_t(D, s->Cinv, CinvT);
  matmul_t(D, s->Cinv, dE, t1);
  matmul_t(D, t1, CinvT, t2);
  double cd = inner_t(D * D, s->Cinv, dE);
  for (int i = 0; i < D * D; i++) dS[i] = lambda * cd * s->Cinv[i] + 2.0 * (mu - lambda * s->lnJ) * t2[i];
}

/* ---- local matrix of one block (test field ft = row, trial field fu = column) ----
 * Ke column-major [ni][nj] like a Julia Matrix: Ke[i + ni*j].  IntegrationMap order: for j, for i, sum over p. */
static void cell_block_matrix(int form, int bi, int bj, const orc_geom_t *g, const orc_field_t *ft, const orc_field_t *fu,
                              const cellgeo_t *cg, const double *Gt, const double *Gu, const double *params,
                              const orc_field_t *state, const double *Gs, int64_t cell, double *Ke) {
  int D = g->D, np = g->np;
  int ni = ft->nds * ft->ncomp, nj = fu->nds * fu->ncomp;
  double aq[MAXQ];        /* aq[p] for the current (i,j) */
  nh_state_t nh[MAXQ];
  if (form == ORC_NEOHOOKEAN_JAC)
    for (int p = 0; p < np; p++) { double gu[9]; state_gradient(g, state, Gs, cell, p, gu); nh_state(D, gu, params[0], params[1], &nh[p]); }
  for (int j = 0; j < nj; j++) {
    int b = j % fu->nds, cj = j / fu->nds;
    for (int i = 0; i < ni; i++) {
      int a = i % ft->nds, ci = i / ft->nds;
      for (int p = 0; p < np; p++) {
        const double *ga = Gt + ((int64_t)p * ft->nds + a) * D;
        const double *gb = Gu + ((int64_t)p * fu->nds + b) * D;
        double Na = ft->N[(int64_t)(cg->p0 + p) * ft->nds + a], Nb = fu->N[(int64_t)(cg->p0 + p) * fu->nds + b];
        double v = 0.0;
        switch (form) {
          case ORC_MASS: v = (ci == cj) ? Na * Nb : 0.0; break;
          case ORC_FACET: {
            /* coef * T_i * U_j on equal components; T, U = value or normal derivative (e.g. the Nitsche terms
               (gamma/h) v u - v (n.grad u) - (n.grad v) u of test/GridapTests/PoissonTests.jl:99-101) */
            if (ci == cj) {
              double T = (int)params[1] == 0 ? Na : inner_t(D, cg->nrm[p], ga);
              double U = (int)params[2] == 0 ? Nb : inner_t(D, cg->nrm[p], gb);
              v = params[0] * T * U;
            }
          } break;
          case ORC_LAPLACIAN: {
            if (ft->ncomp == 1) v = inner_t(D, ga, gb);
            else { double A[9], B[9]; basis_grad_tensor(D, ga, ci, A); basis_grad_tensor(D, gb, cj, B); v = inner_t(D * D, A, B); }
          } break;
          case ORC_ELASTICITY: {
            double A[9], B[9], ev[9], eu[9], sig[9];
            basis_grad_tensor(D, ga, ci, A); basis_grad_tensor(D, gb, cj, B);
            sym_t(D, A, ev); sym_t(D, B, eu);
            double tr = trace_t(D, eu);
            for (int k = 0; k < D * D; k++) sig[k] = 2.0 * params[1] * eu[k];
            for (int k = 0; k < D; k++) sig[k * D + k] += params[0] * tr;
            v = inner_t(D * D, ev, sig);
          } break;
          case ORC_STOKES: {
            /* a((u,p),(v,q)) = grad(v) : grad(u) - (div v) p + q (div u); field 0 = velocity, 1 = pressure */
            if (bi == 0 && bj == 0) { double A[9], B[9]; basis_grad_tensor(D, ga, ci, A); basis_grad_tensor(D, gb, cj, B); v = inner_t(D * D, A, B); }
            else if (bi == 0 && bj == 1) { double A[9]; basis_grad_tensor(D, ga, ci, A); v = -trace_t(D, A) * Nb; }
            else if (bi == 1 && bj == 0) { double B[9]; basis_grad_tensor(D, gb, cj, B); v = Na * trace_t(D, B); }
          } break;
          case ORC_NEOHOOKEAN_JAC: {
            /* jac(u,du,v) = dE(grad v,grad u) : dS(grad du,grad u) + grad(v) : (S . grad(du)) */
            double A[9], B[9], dEv[9], dEu[9], dS[9], SB[9];
            basis_grad_tensor(D, ga, ci, A); basis_grad_tensor(D, gb, cj, B);
            nh_dE(D, A, nh[p].F, dEv); nh_dE(D, B, nh[p].F, dEu);
            nh_dS(D, dEu, &nh[p], params[0], params[1], dS);
            matmul_t(D, nh[p].S, B, SB);
            v = inner_t(D * D, dEv, dS) + inner_t(D * D, A, SB);
          } break;
        }
        aq[p] = v;
      }
      double rij = 0.0;
      for (int p = 0; p < np; p++) rij += aq[p] * cg->dV[p];
      Ke[i + (int64_t)ni * j] = rij;
    }
  }
}

/* local vector of one test field */
static void cell_block_vector(int form, const orc_geom_t *g, const orc_field_t *ft, const cellgeo_t *cg, const double *Gt,
                              const double *params, const double *fq /* [ncells][np][ncomp] or NULL */,
                              const orc_field_t *state, const double *Gs, int64_t cell, double *be) {
  int D = g->D, np = g->np, ni = ft->nds * ft->ncomp;
  for (int i = 0; i < ni; i++) be[i] = 0.0;
  for (int p = 0; p < np; p++) {
    nh_state_t nh;
    if (form == ORC_NEOHOOKEAN_RES) { double gu[9]; state_gradient(g, state, Gs, cell, p, gu); nh_state(D, gu, params[0], params[1], &nh); }
    for (int i = 0; i < ni; i++) {
      int a = i % ft->nds, ci = i / ft->nds;
      double v = 0.0;
      if (form == ORC_SOURCE) {
        double f = ft->fq ? ft->fq[((int64_t)cell * np + p) * ft->ncomp + ci] : ft->src ? ft->src[ci]
                   : fq ? fq[((int64_t)cell * np + p) * ft->ncomp + ci] : params[ci];
        v = ft->N[(int64_t)(cg->p0 + p) * ft->nds + a] * f;
      } else if (form == ORC_FACET_VEC) {
        /* coef * T_i * data: l(v) terms of the Nitsche / Neumann kind (PoissonTests.jl:103-107): data = g at the points,
           u_h or n.grad(u_h) of the state field (same component as the test function) */
        const double *ga = Gt + ((int64_t)p * ft->nds + a) * D;
        double T = (int)params[1] == 0 ? ft->N[(int64_t)(cg->p0 + p) * ft->nds + a] : inner_t(D, cg->nrm[p], ga);
        double dat = 0.0;
        int dk = (int)params[2];
        if (dk == 0) dat = fq[((int64_t)cell * np + p) * ft->ncomp + ci];
        else
          for (int j = 0; j < state->nds; j++) {
            double uj = dof_value(state, cell, j + state->nds * ci);
            dat += uj * (dk == 1 ? state->N[(int64_t)(cg->p0 + p) * state->nds + j] : inner_t(D, cg->nrm[p], Gs + ((int64_t)p * state->nds + j) * D));
          }
        v = params[0] * T * dat;
      } else if (form == ORC_NEOHOOKEAN_RES) {
        double A[9], dEv[9];
        basis_grad_tensor(D, Gt + ((int64_t)p * ft->nds + a) * D, ci, A);
        nh_dE(D, A, nh.F, dEv);
        v = inner_t(D * D, dEv, nh.S);
      }
      be[i] += v * cg->dV[p];
    }
  }
}

/* ------------------------------------------------------------------ sparse builder (SparseMatrixCSC.jl) */
typedef struct {
  int64_t nrows, ncols;
  int64_t *colptr, *colnnz, *rowval; /* 1-based contents */
  double *nzval;
} inserter_t;

static int64_t searchsortedfirst(const int64_t *v /* 1-based view: v[k-1] */, int64_t x, int64_t lo, int64_t hi) {
  /* Base.searchsortedfirst(v,x,lo,hi,Forward): first index in lo:hi with v[k] >= x, else hi+1 */
  lo = lo - 1; hi = hi + 1;
  while (lo < hi - 1) {
    int64_t m = lo + ((hi - lo) >> 1);
    if (v[m - 1] < x) lo = m; else hi = m;
  }
  return hi;
}

static void inserter_add(inserter_t *a, int has_v, double v, int64_t i, int64_t j) {
  int64_t pini = a->colptr[j - 1];
  int64_t pend = pini + a->colnnz[j - 1] - 1;
  int64_t p = searchsortedfirst(a->rowval, i, pini, pend);
  if (p > pend) {
    a->colnnz[j - 1] += 1; a->rowval[p - 1] = i; if (has_v) a->nzval[p - 1] = v;
  } else if (a->rowval[p - 1] != i) {
    for (int64_t k = pend; k >= p; k--) { a->rowval[k] = a->rowval[k - 1]; a->nzval[k] = a->nzval[k - 1]; }
    a->colnnz[j - 1] += 1; a->rowval[p - 1] = i; a->nzval[p - 1] = has_v ? v : 0.0;
  } else if (has_v) {
    a->nzval[p - 1] += v;
  }
}

/* nz_index(A,i,j) (SparseMatrixCSC.jl:14-22) on a final CSC; returns 1-based position or -1 */
static int64_t nz_index(const int64_t *colptr, const int64_t *rowval, int64_t i, int64_t j) {
  int64_t r1 = colptr[j - 1], r2 = colptr[j] - 1;
  if (r1 > r2) return -1;
  r1 = searchsortedfirst(rowval, i, r1, r2);
  return (r1 > r2 || rowval[r1 - 1] != i) ? -1 : r1;
}

/* ------------------------------------------------------------------ exported entry points */
typedef struct {
  int32_t form_mat, form_vec;  /* 0 = none */
  int32_t nfields;
  const orc_field_t *fields;   /* test == trial spaces (Galerkin) */
  const uint8_t *touched;      /* [nfields][nfields] row-major (bi,bj); NULL => all */
  const double *params;
  const double *fq;            /* source at quadrature points or NULL */
  int32_t state_field;         /* index of the field carrying u_h (neo-Hookean) */
  int32_t lift_dirichlet;      /* matrix_and_vector: b_e -= K_e u_e on Dirichlet cells */
  int64_t nrows, ncols;
} orc_problem_t;

static int is_touched(const orc_problem_t *pb, int bi, int bj) { return pb->touched ? pb->touched[bi * pb->nfields + bj] : 1; }

/* symbolic loop: colnnzmax[j] += 1 for each admissible pair (SparseMatrixAssemblers.jl:174-210) */
void orc_symbolic_count(const orc_geom_t *g, const orc_problem_t *pb, int64_t *colnnzmax) {
  for (int64_t j = 0; j < pb->ncols; j++) colnnzmax[j] = 0;
  for (int64_t cell = 0; cell < g->ncells; cell++)
    for (int bj = 0; bj < pb->nfields; bj++)
      for (int bi = 0; bi < pb->nfields; bi++) {
        if (!is_touched(pb, bi, bj)) continue;
        const orc_field_t *fu = &pb->fields[bj], *ft = &pb->fields[bi];
        int nj = fu->nds * fu->ncomp, ni = ft->nds * ft->ncomp;
        for (int lj = 0; lj < nj; lj++) {
          int32_t j = fu->cell_dofs[cell * (int64_t)nj + lj];
          if (j <= 0) continue;
          for (int li = 0; li < ni; li++) {
            int32_t i = ft->cell_dofs[cell * (int64_t)ni + li];
            if (i > 0) colnnzmax[j - 1] += 1;
          }
        }
      }
}

/* Computes all local blocks of `cell` into Kblk[bi][bj] (column-major each) and bblk[bi]; applies lifting. */
typedef struct { double *K[4][4]; double *b[4]; double *G[4]; } cellwork_t;

static void cell_compute(const orc_geom_t *g, const orc_problem_t *pb, int64_t cell, cellwork_t *wk, cellgeo_t *cg) {
  cell_geometry(g, cell, cg);
  for (int f = 0; f < pb->nfields; f++) phys_grads(g, &pb->fields[f], cg, wk->G[f]);
  const orc_field_t *state = &pb->fields[pb->state_field];
  if (pb->form_mat)
    for (int bj = 0; bj < pb->nfields; bj++)
      for (int bi = 0; bi < pb->nfields; bi++)
        if (is_touched(pb, bi, bj))
          cell_block_matrix(pb->form_mat, bi, bj, g, &pb->fields[bi], &pb->fields[bj], cg, wk->G[bi], wk->G[bj], pb->params,
                            state, wk->G[pb->state_field], cell, wk->K[bi][bj]);
  if (pb->form_vec)
    for (int bi = 0; bi < pb->nfields; bi++)
      cell_block_vector(pb->form_vec, g, &pb->fields[bi], cg, wk->G[bi], pb->params, pb->fq, state, wk->G[pb->state_field], cell, wk->b[bi]);
  if (pb->form_mat && pb->form_vec && pb->lift_dirichlet) {
    /* AttachDirichletMap: only cells with a Dirichlet dof; vec = vec - mat*vals, vals = 0 on free dofs */
    int any = 0;
    for (int f = 0; f < pb->nfields; f++) {
      int n = pb->fields[f].nds * pb->fields[f].ncomp;
      for (int k = 0; k < n; k++) if (pb->fields[f].cell_dofs[cell * (int64_t)n + k] < 0) any = 1;
    }
    if (any)
      for (int bi = 0; bi < pb->nfields; bi++) {
        int ni = pb->fields[bi].nds * pb->fields[bi].ncomp;
        for (int bj = 0; bj < pb->nfields; bj++) {
          if (!is_touched(pb, bi, bj)) continue;
          const orc_field_t *fu = &pb->fields[bj];
          int nj = fu->nds * fu->ncomp;
          for (int j = 0; j < nj; j++) {
            int32_t id = fu->cell_dofs[cell * (int64_t)nj + j];
            double uj = (id < 0 && fu->dirichlet_values) ? fu->dirichlet_values[-id - 1] : 0.0;
            for (int i = 0; i < ni; i++) wk->b[bi][i] -= wk->K[bi][bj][i + (int64_t)ni * j] * uj;
          }
        }
      }
  }
}

static void work_alloc(const orc_geom_t *g, const orc_problem_t *pb, cellwork_t *wk) {
  memset(wk, 0, sizeof(*wk));
  for (int bi = 0; bi < pb->nfields; bi++) {
    int ni = pb->fields[bi].nds * pb->fields[bi].ncomp;
    wk->b[bi] = (double *)calloc(ni, sizeof(double));
    wk->G[bi] = (double *)calloc((size_t)g->np * pb->fields[bi].nds * g->D, sizeof(double));
    for (int bj = 0; bj < pb->nfields; bj++) {
      int nj = pb->fields[bj].nds * pb->fields[bj].ncomp;
      wk->K[bi][bj] = (double *)calloc((size_t)ni * nj, sizeof(double));
    }
  }
}
static void work_free(const orc_problem_t *pb, cellwork_t *wk) {
  for (int bi = 0; bi < pb->nfields; bi++) { free(wk->b[bi]); free(wk->G[bi]); for (int bj = 0; bj < pb->nfields; bj++) free(wk->K[bi][bj]); }
}

/* local matrices / vectors of one cell, for unit tests of the integrands */
void orc_cell_local(const orc_geom_t *g, const orc_problem_t *pb, int64_t cell, double **Kout /* [nf*nf] */, double **bout /* [nf] */) {
  cellwork_t wk; work_alloc(g, pb, &wk);
  cellgeo_t *cg = (cellgeo_t *)malloc(sizeof(cellgeo_t));
  cell_compute(g, pb, cell, &wk, cg);
  for (int bi = 0; bi < pb->nfields; bi++) {
    int ni = pb->fields[bi].nds * pb->fields[bi].ncomp;
    if (bout && bout[bi]) memcpy(bout[bi], wk.b[bi], ni * sizeof(double));
    for (int bj = 0; bj < pb->nfields; bj++) {
      int nj = pb->fields[bj].nds * pb->fields[bj].ncomp;
      if (Kout && Kout[bi * pb->nfields + bj]) memcpy(Kout[bi * pb->nfields + bj], wk.K[bi][bj], (size_t)ni * nj * sizeof(double));
    }
  }
  free(cg); work_free(pb, &wk);
}

/* CONTEXT ONLY, NOT THE REFERENCE ALGORITHM: the per-cell quadrature (a4-a8) of every cell, without the sparse insertion, spread over
 * `nthreads` POSIX threads (the reference's assembly loop is serial; this image has no OpenMP runtime).  Returns the sum of all
 * local-matrix entries of field block (0,0) as a checksum.  bench.py reports its throughput next to the serial baseline to show how
 * the CPU time splits between quadrature and the CSC insertion. */
typedef struct { const orc_geom_t *g; const orc_problem_t *pb; int64_t c0, c1; double total; } qonly_job_t;
static void *qonly_worker(void *arg) {
  qonly_job_t *job = (qonly_job_t *)arg;
  cellwork_t wk; work_alloc(job->g, job->pb, &wk);
  cellgeo_t *cg = (cellgeo_t *)malloc(sizeof(cellgeo_t));
  const int n0 = job->pb->fields[0].nds * job->pb->fields[0].ncomp;
  double total = 0.0;
  for (int64_t cell = job->c0; cell < job->c1; cell++) {
    cell_compute(job->g, job->pb, cell, &wk, cg);
    for (int e = 0; e < n0 * n0; e++) total += wk.K[0][0][e];
  }
  job->total = total;
  free(cg); work_free(job->pb, &wk);
  return NULL;
}
double orc_quadrature_only(const orc_geom_t *g, const orc_problem_t *pb, int32_t nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  qonly_job_t jobs[256];
  for (int t = 0; t < nthreads; t++) {
    jobs[t].g = g; jobs[t].pb = pb; jobs[t].total = 0.0;
    jobs[t].c0 = g->ncells * t / nthreads; jobs[t].c1 = g->ncells * (t + 1) / nthreads;
    pthread_create(&th[t], NULL, qonly_worker, &jobs[t]);
  }
  double total = 0.0;
  for (int t = 0; t < nthreads; t++) { pthread_join(th[t], NULL); total += jobs[t].total; }
  return total;
}

/* quadrature points in physical space xq[cell][p][D] */
void orc_quadrature_points(const orc_geom_t *g, double *xq) {
  cellgeo_t *cg = (cellgeo_t *)malloc(sizeof(cellgeo_t));
  for (int64_t cell = 0; cell < g->ncells; cell++) {
    cell_geometry(g, cell, cg);
    for (int p = 0; p < g->np; p++) for (int d = 0; d < g->D; d++) xq[((int64_t)cell * g->np + p) * g->D + d] = cg->xq[p][d];
  }
  free(cg);
}

/* assemble_matrix / assemble_matrix_and_vector from scratch:
 * nz_counter -> symbolic loop -> nz_allocation -> numeric loop -> create_from_nz (SparseMatrixAssemblers.jl:70-106).
 * Caller passes colptr[ncols+1]; rowval/nzval are malloc'ed here (size = final nnz) and returned. b may be NULL. */
int64_t orc_assemble(const orc_geom_t *g, const orc_problem_t *pb, int64_t *colptr, int64_t **rowval_out, double **nzval_out, double *b) {
  inserter_t a;
  a.nrows = pb->nrows; a.ncols = pb->ncols;
  a.colnnz = (int64_t *)calloc(pb->ncols + 1, sizeof(int64_t));
  a.colptr = (int64_t *)calloc(pb->ncols + 1, sizeof(int64_t));
  orc_symbolic_count(g, pb, a.colnnz);
  /* nz_allocation: colptr[i+1] = colnnzmax[i]; length_to_ptrs! */
  a.colptr[0] = 1;
  for (int64_t j = 0; j < pb->ncols; j++) a.colptr[j + 1] = a.colptr[j] + a.colnnz[j];
  int64_t ndata = a.colptr[pb->ncols] - 1;
  a.rowval = (int64_t *)malloc((ndata + 1) * sizeof(int64_t));
  a.nzval = (double *)calloc(ndata + 1, sizeof(double));
  for (int64_t j = 0; j < pb->ncols; j++) a.colnnz[j] = 0;
  if (b) for (int64_t i = 0; i < pb->nrows; i++) b[i] = 0.0;

  cellwork_t wk; work_alloc(g, pb, &wk);
  cellgeo_t *cg = (cellgeo_t *)malloc(sizeof(cellgeo_t));
  for (int64_t cell = 0; cell < g->ncells; cell++) {
    cell_compute(g, pb, cell, &wk, cg);
    for (int bj = 0; bj < pb->nfields; bj++)
      for (int bi = 0; bi < pb->nfields; bi++) {
        if (!is_touched(pb, bi, bj)) continue;
        const orc_field_t *fu = &pb->fields[bj], *ft = &pb->fields[bi];
        int nj = fu->nds * fu->ncomp, ni = ft->nds * ft->ncomp;
        for (int lj = 0; lj < nj; lj++) {
          int32_t j = fu->cell_dofs[cell * (int64_t)nj + lj];
          if (j <= 0) continue;
          for (int li = 0; li < ni; li++) {
            int32_t i = ft->cell_dofs[cell * (int64_t)ni + li];
            if (i > 0) inserter_add(&a, 1, wk.K[bi][bj][li + (int64_t)ni * lj], i, j);
          }
        }
      }
    if (b && pb->form_vec)
      for (int bi = 0; bi < pb->nfields; bi++) {
        const orc_field_t *ft = &pb->fields[bi];
        int ni = ft->nds * ft->ncomp;
        for (int li = 0; li < ni; li++) {
          int32_t i = ft->cell_dofs[cell * (int64_t)ni + li];
          if (i > 0) b[i - 1] += wk.b[bi][li];
        }
      }
  }
  free(cg); work_free(pb, &wk);
  /* create_from_nz: compact columns to the left */
  int64_t k = 1;
  for (int64_t j = 0; j < pb->ncols; j++) {
    int64_t pini = a.colptr[j], pend = pini + a.colnnz[j] - 1;
    for (int64_t p = pini; p <= pend; p++) { a.nzval[k - 1] = a.nzval[p - 1]; a.rowval[k - 1] = a.rowval[p - 1]; k++; }
  }
  colptr[0] = 1;
  for (int64_t j = 0; j < pb->ncols; j++) colptr[j + 1] = colptr[j] + a.colnnz[j];
  int64_t nnz = colptr[pb->ncols] - 1;
  *rowval_out = (int64_t *)realloc(a.rowval, (nnz + 1) * sizeof(int64_t));
  *nzval_out = (double *)realloc(a.nzval, (nnz + 1) * sizeof(double));
  free(a.colnnz); free(a.colptr);
  return nnz;
}

/* in-place variants on an existing pattern: assemble_matrix!/_add! (nz_index) and assemble_vector!/_add!
 * (SparseMatrixAssemblers.jl:32-40,60-68,88-97).  add_flag=0 => fillstored!(A,0)/fill!(b,0) first.
 * Returns the number of entries that were not found in the pattern (must be 0). */
int64_t orc_assemble_inplace(const orc_geom_t *g, const orc_problem_t *pb, const int64_t *colptr, const int64_t *rowval,
                             double *nzval, double *b, int add_flag) {
  int64_t missing = 0;
  int64_t nnz = colptr ? colptr[pb->ncols] - 1 : 0;
  if (!add_flag) {
    if (nzval) for (int64_t k = 0; k < nnz; k++) nzval[k] = 0.0;
    if (b) for (int64_t i = 0; i < pb->nrows; i++) b[i] = 0.0;
  }
  cellwork_t wk; work_alloc(g, pb, &wk);
  cellgeo_t *cg = (cellgeo_t *)malloc(sizeof(cellgeo_t));
  for (int64_t cell = 0; cell < g->ncells; cell++) {
    cell_compute(g, pb, cell, &wk, cg);
    if (nzval && pb->form_mat)
      for (int bj = 0; bj < pb->nfields; bj++)
        for (int bi = 0; bi < pb->nfields; bi++) {
          if (!is_touched(pb, bi, bj)) continue;
          const orc_field_t *fu = &pb->fields[bj], *ft = &pb->fields[bi];
          int nj = fu->nds * fu->ncomp, ni = ft->nds * ft->ncomp;
          for (int lj = 0; lj < nj; lj++) {
            int32_t j = fu->cell_dofs[cell * (int64_t)nj + lj];
            if (j <= 0) continue;
            for (int li = 0; li < ni; li++) {
              int32_t i = ft->cell_dofs[cell * (int64_t)ni + li];
              if (i <= 0) continue;
              int64_t k = nz_index(colptr, rowval, i, j);
              if (k < 0) { missing++; continue; }
              nzval[k - 1] += wk.K[bi][bj][li + (int64_t)ni * lj];
            }
          }
        }
    if (b && pb->form_vec)
      for (int bi = 0; bi < pb->nfields; bi++) {
        const orc_field_t *ft = &pb->fields[bi];
        int ni = ft->nds * ft->ncomp;
        for (int li = 0; li < ni; li++) {
          int32_t i = ft->cell_dofs[cell * (int64_t)ni + li];
          if (i > 0) b[i - 1] += wk.b[bi][li];
        }
      }
  }
  free(cg); work_free(pb, &wk);
  return missing;
}

/* assemble with one constant local matrix for every cell (the Fill(K_e,ncells) case of a
 * CartesianDiscreteModel: src/Arrays/LazyArrays.jl:302-322, src/Geometry/CartesianGrids.jl:271-276) */
int64_t orc_assemble_const(int64_t ncells, int32_t nd, const int32_t *cell_dofs, const double *Ke /* col-major */, int64_t nrows,
                           int64_t ncols, int64_t *colptr, int64_t **rowval_out, double **nzval_out) {
  inserter_t a;
  a.nrows = nrows; a.ncols = ncols;
  a.colnnz = (int64_t *)calloc(ncols + 1, sizeof(int64_t));
  a.colptr = (int64_t *)calloc(ncols + 1, sizeof(int64_t));
  for (int64_t cell = 0; cell < ncells; cell++)
    for (int lj = 0; lj < nd; lj++) {
      int32_t j = cell_dofs[cell * nd + lj];
      if (j <= 0) continue;
      for (int li = 0; li < nd; li++) if (cell_dofs[cell * nd + li] > 0) a.colnnz[j - 1] += 1;
    }
  a.colptr[0] = 1;
  for (int64_t j = 0; j < ncols; j++) a.colptr[j + 1] = a.colptr[j] + a.colnnz[j];
  int64_t ndata = a.colptr[ncols] - 1;
  a.rowval = (int64_t *)malloc((ndata + 1) * sizeof(int64_t));
  a.nzval = (double *)calloc(ndata + 1, sizeof(double));
  for (int64_t j = 0; j < ncols; j++) a.colnnz[j] = 0;
  for (int64_t cell = 0; cell < ncells; cell++)
    for (int lj = 0; lj < nd; lj++) {
      int32_t j = cell_dofs[cell * nd + lj];
      if (j <= 0) continue;
      for (int li = 0; li < nd; li++) {
        int32_t i = cell_dofs[cell * nd + li];
        if (i > 0) inserter_add(&a, 1, Ke[li + nd * lj], i, j);
      }
    }
  int64_t k = 1;
  for (int64_t j = 0; j < ncols; j++) {
    int64_t pini = a.colptr[j], pend = pini + a.colnnz[j] - 1;
    for (int64_t p = pini; p <= pend; p++) { a.nzval[k - 1] = a.nzval[p - 1]; a.rowval[k - 1] = a.rowval[p - 1]; k++; }
  }
  colptr[0] = 1;
  for (int64_t j = 0; j < ncols; j++) colptr[j + 1] = colptr[j] + a.colnnz[j];
  int64_t nnz = colptr[ncols] - 1;
  *rowval_out = (int64_t *)realloc(a.rowval, (nnz + 1) * sizeof(int64_t));
  *nzval_out = (double *)realloc(a.nzval, (nnz + 1) * sizeof(double));
  free(a.colnnz); free(a.colptr);
  return nnz;
}

void orc_free(void *p) { free(p); }

/* ---- raw builder protocol, for the golden test test/AlgebraTests/AlgebraInterfacesTests.jl:106-152 ---- */
typedef struct { inserter_t a; int64_t cap; } orc_builder_t;

orc_builder_t *orc_builder_from_counts(int64_t nrows, int64_t ncols, const int64_t *colnnzmax) {
  orc_builder_t *b = (orc_builder_t *)calloc(1, sizeof(*b));
  b->a.nrows = nrows; b->a.ncols = ncols;
  b->a.colptr = (int64_t *)calloc(ncols + 1, sizeof(int64_t));
  b->a.colnnz = (int64_t *)calloc(ncols + 1, sizeof(int64_t));
  b->a.colptr[0] = 1;
  for (int64_t j = 0; j < ncols; j++) b->a.colptr[j + 1] = b->a.colptr[j] + colnnzmax[j];
  b->cap = b->a.colptr[ncols] - 1;
  b->a.rowval = (int64_t *)calloc(b->cap + 1, sizeof(int64_t));
  b->a.nzval = (double *)calloc(b->cap + 1, sizeof(double));
  return b;
}
void orc_builder_add(orc_builder_t *b, int has_v, double v, int64_t i, int64_t j) { if (i > 0 && j > 0) inserter_add(&b->a, has_v, v, i, j); }
void orc_builder_state(const orc_builder_t *b, int64_t *colptr, int64_t *colnnz) {
  memcpy(colptr, b->a.colptr, (b->a.ncols + 1) * sizeof(int64_t));
  memcpy(colnnz, b->a.colnnz, b->a.ncols * sizeof(int64_t));
}
int64_t orc_builder_finish(orc_builder_t *b, int64_t *colptr, int64_t *rowval, double *nzval) {
  int64_t k = 1;
  for (int64_t j = 0; j < b->a.ncols; j++) {
    int64_t pini = b->a.colptr[j], pend = pini + b->a.colnnz[j] - 1;
    for (int64_t p = pini; p <= pend; p++) { nzval[k - 1] = b->a.nzval[p - 1]; rowval[k - 1] = b->a.rowval[p - 1]; k++; }
  }
  colptr[0] = 1;
  for (int64_t j = 0; j < b->a.ncols; j++) colptr[j + 1] = colptr[j] + b->a.colnnz[j];
  int64_t nnz = colptr[b->a.ncols] - 1;
  free(b->a.colptr); free(b->a.colnnz); free(b->a.rowval); free(b->a.nzval); free(b);
  return nnz;
}
