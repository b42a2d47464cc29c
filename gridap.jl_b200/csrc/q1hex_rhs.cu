// q1hex_rhs.cu -- local vectors for scalar Q1 hexahedra: b_e = int v f  with the Dirichlet lifting fused.
//
// Reference: IntegrationMap for vectors (src/Fields/FieldsInterfaces.jl:762-776), AttachDirichletMap
// `b_e <- b_e - K_e u_e` on cells that touch a Dirichlet DoF (src/CellData/AttachDirichlet.jl:76-84), scatter
// `b[i] += b_e[li]` for i > 0 (src/Algebra/AlgebraInterfaces.jl:202-210).
//
// One thread per cell, general (trilinear) geometry, the full quadrature loop.  The lifting never forms K_e: with
// t(p) = sum_{j Dirichlet} u_j grad(phi_j)(p) the Laplacian gives  (K_e u_e)_i = sum_p dV_p grad(phi_i)(p) . t(p)
// (mass: s(p) = sum_j u_j N_j(p),  (M_e u_e)_i = sum_p dV_p N_i(p) s(p)).  Scatter: 8 RED.ADD.F64 per cell.
#include "common.cuh"

namespace gb {

namespace {

struct RhsArgs {
  const double *X;
  const int32_t *cell_nodes, *row_ids, *col_ids;
  const double *w, *N, *dN;  // tabulation at the 8 quadrature points: N[p][a], dN[p][a][3]
  const double *dir_vals;
  const double *fq;          // [cell][p] or null
  double f0, coef;
  int lift_form;             // 0 none, GB200_FORM_LAPLACIAN or GB200_FORM_MASS
  int64_t ncells, row_off;
  double *bvec;
  const int32_t *cell_list;  // LIFT pass: the cells that touch a Dirichlet DoF
};

// LIFT = false: source term on every cell.  LIFT = true: only -K_e u_e, on the listed cells (those with a Dirichlet DoF; a few
// per cent of a large mesh -- inside one pass they made every fourth warp of an x-fastest mesh run the divergent lifting code).
// AFFINE (every cell map of the mesh is affine, checked once per mesh): the Jacobian is the same at all quadrature points.
template <bool LIFT, bool AFFINE>
__global__ void __launch_bounds__(128) q1hex_rhs_kernel(RhsArgs k) {
  __shared__ double s_w[8], s_N[64], s_dN[192];
  for (int i = threadIdx.x; i < 192; i += blockDim.x) {
    s_dN[i] = k.dN[i];
    if (i < 64) s_N[i] = k.N[i];
    if (i < 8) s_w[i] = k.w[i];
  }
  __syncthreads();
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= k.ncells) return;
  const int64_t c = LIFT ? (int64_t)k.cell_list[t] : t;
  const int4 *cn = reinterpret_cast<const int4 *>(k.cell_nodes + c * 8);
  const int4 n0 = cn[0], n1 = cn[1];
  const int ids[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
  const int4 *rp = reinterpret_cast<const int4 *>(k.row_ids + c * 8);
  const int4 r0 = rp[0], r1 = rp[1];
  const int rows[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
  double u[8];
  bool any_dir = false;
  if (LIFT) {
    const int4 *cp = reinterpret_cast<const int4 *>(k.col_ids + c * 8);
    const int4 c0 = cp[0], c1 = cp[1];
    const int cols[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
    for (int a = 0; a < 8; a++) {
      u[a] = (cols[a] < 0 && k.dir_vals) ? k.dir_vals[-cols[a] - 1] : 0.0;
      any_dir |= cols[a] < 0;
    }
  }
  double x[8][3];
#pragma unroll
  for (int a = 0; a < 8; a++) {
    const double *p = k.X + (int64_t)ids[a] * 3;
    x[a][0] = p[0]; x[a][1] = p[1]; x[a][2] = p[2];
  }
  double b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const bool need_grads = LIFT && any_dir && k.lift_form == GB200_FORM_LAPLACIAN;
  double J[9];
#pragma unroll 1
  for (int p = 0; p < 8; p++) {
    if (!AFFINE || p == 0) {
#pragma unroll
    for (int i = 0; i < 9; i++) J[i] = 0.0;
#pragma unroll
    for (int a = 0; a < 8; a++) {
      const double d0 = s_dN[(p * 8 + a) * 3], d1 = s_dN[(p * 8 + a) * 3 + 1], d2 = s_dN[(p * 8 + a) * 3 + 2];
#pragma unroll
      for (int j = 0; j < 3; j++) {
        J[j] += d0 * x[a][j];
        J[3 + j] += d1 * x[a][j];
        J[6 + j] += d2 * x[a][j];
      }
    }
    }
    const double det = J[0] * J[4] * J[8] + J[1] * J[5] * J[6] + J[2] * J[3] * J[7] - (J[0] * J[5] * J[7] + J[1] * J[3] * J[8] + J[2] * J[4] * J[6]);
    const double dV = fabs(det) * s_w[p];
    const double f = LIFT ? 0.0 : (k.fq ? k.fq[c * 8 + p] : k.f0);
    double lift_s = 0.0;
    if (LIFT && any_dir && k.lift_form == GB200_FORM_MASS) {
#pragma unroll
      for (int a = 0; a < 8; a++) lift_s += u[a] * s_N[p * 8 + a];
      lift_s *= k.coef;
    }
#pragma unroll
    for (int a = 0; a < 8; a++) b[a] += s_N[p * 8 + a] * (f - lift_s) * dV;
    if (need_grads) {
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
      double g[8][3], t0 = 0.0, t1 = 0.0, t2 = 0.0;
#pragma unroll
      for (int a = 0; a < 8; a++) {
        const double d0 = s_dN[(p * 8 + a) * 3], d1 = s_dN[(p * 8 + a) * 3 + 1], d2 = s_dN[(p * 8 + a) * 3 + 2];
        g[a][0] = I[0] * d0 + I[1] * d1 + I[2] * d2;
        g[a][1] = I[3] * d0 + I[4] * d1 + I[5] * d2;
        g[a][2] = I[6] * d0 + I[7] * d1 + I[8] * d2;
        t0 += u[a] * g[a][0]; t1 += u[a] * g[a][1]; t2 += u[a] * g[a][2];
      }
      const double s = k.coef * dV;
#pragma unroll
      for (int a = 0; a < 8; a++) b[a] -= s * (g[a][0] * t0 + g[a][1] * t1 + g[a][2] * t2);
    }
  }
#pragma unroll
  for (int a = 0; a < 8; a++)
    if (rows[a] > 0) atomicAdd(k.bvec + (rows[a] - 1 + k.row_off), b[a]);
}

// Source term on axis-aligned boxes (cells_are_boxes: every Cartesian mesh): |det Jt| = hx hy hz from 4 nodes / 6 coordinates instead
// of the 8-node Jacobian -- the kernel is left with its compulsory traffic (ids, f at the points, the REDs).
__global__ void __launch_bounds__(128) q1hex_rhs_box_kernel(RhsArgs k) {
  __shared__ double s_wN[64];   // w_p N[p][a]
  for (int i = threadIdx.x; i < 64; i += blockDim.x) s_wN[i] = k.w[i >> 3] * k.N[i];
  __syncthreads();
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= k.ncells) return;
  const int4 n = *reinterpret_cast<const int4 *>(k.cell_nodes + c * 8);
  const int n4 = k.cell_nodes[c * 8 + 4];
  const int4 *rp = reinterpret_cast<const int4 *>(k.row_ids + c * 8);
  const int4 r0 = rp[0], r1 = rp[1];
  const int rows[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
  const double *p0 = k.X + (int64_t)n.x * 3;
  const double det = fabs((k.X[(int64_t)n.y * 3] - p0[0]) * (k.X[(int64_t)n.z * 3 + 1] - p0[1]) * (k.X[(int64_t)n4 * 3 + 2] - p0[2]));
  double f[8];
  if (k.fq) {
    const double2 *fp = reinterpret_cast<const double2 *>(k.fq + c * 8);
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const double2 v = fp[q];
      f[2 * q] = v.x;
      f[2 * q + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int p = 0; p < 8; p++) f[p] = k.f0;
  }
#pragma unroll
  for (int a = 0; a < 8; a++) {
    double b = 0.0;
#pragma unroll
    for (int p = 0; p < 8; p++) b += s_wN[p * 8 + a] * f[p];
    if (rows[a] > 0) atomicAdd(k.bvec + (rows[a] - 1 + k.row_off), b * det);
  }
}

__global__ void dirichlet_cells_kernel(const int32_t *col_ids, int64_t ncells, int32_t *list, unsigned long long *count) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  const int4 *cp = reinterpret_cast<const int4 *>(col_ids + c * 8);
  const int4 a = cp[0], b = cp[1];
  if ((a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w) < 0) list[atomicAdd(count, 1ull)] = (int32_t)c;
}

}  // namespace

// Scalar Q1 hexahedra with the 8-point rule, atomic mode only (the deterministic mode keeps the coloured generic kernel).
bool launch_q1hex_rhs(gb200_plan plan, int form_vec, int lift_form, const double *params, const double *fq, double *bvec) {
  gb200_ctx ctx = plan->ctx;
  const ElemDesc &ed = plan->ed;
  if (ctx->deterministic() || form_vec != GB200_FORM_SOURCE) return false;
  if (plan->nfields != 1 || plan->mesh->celltype != GB200_HEX8 || plan->NL != 8 || ed.np != 8 || ed.f[0].ncomp != 1) return false;
  if (lift_form && lift_form != GB200_FORM_LAPLACIAN && lift_form != GB200_FORM_MASS) return false;
  static const bool disabled = getenv("GB200_NO_Q1_RHS") != nullptr;
  if (disabled) return false;
  RhsArgs k;
  k.X = ed.X; k.cell_nodes = ed.cell_nodes; k.row_ids = ed.f[0].row_ids; k.col_ids = ed.f[0].col_ids;
  k.w = ed.w; k.N = ed.f[0].N; k.dN = ed.f[0].dN; k.dir_vals = ed.f[0].dir_vals; k.fq = fq;
  k.f0 = params[4]; k.coef = params[0]; k.lift_form = lift_form; k.ncells = ed.ncells; k.row_off = ed.f[0].row_off; k.bvec = bvec;
  k.cell_list = nullptr;
  ScopedTimer t(ctx, "k:q1hex_rhs");
  static const bool no_box = getenv("GB200_NO_RHS_BOX") != nullptr;
  if (!no_box && mesh_check_affine(plan->mesh) && cells_are_boxes(plan)) q1hex_rhs_box_kernel<<<(int)((ed.ncells + 127) / 128), 128, 0, ctx->stream>>>(k);
  else if (mesh_check_affine(plan->mesh)) q1hex_rhs_kernel<false, true><<<(int)((ed.ncells + 127) / 128), 128, 0, ctx->stream>>>(k);
  else q1hex_rhs_kernel<false, false><<<(int)((ed.ncells + 127) / 128), 128, 0, ctx->stream>>>(k);
  check_launch(ctx, "q1hex_rhs_kernel");
  if (lift_form && k.dir_vals) {  // homogeneous Dirichlet data (no values set): nothing to lift
    if (plan->n_dir_cells < 0) {  // once per plan: the cells that touch a Dirichlet DoF (any order: the scatter is atomic)
      plan->dir_cells.alloc((size_t)ed.ncells);
      DevBuf<int64_t> cnt;
      cnt.alloc(1);
      cnt.zero(ctx->stream);
      dirichlet_cells_kernel<<<(int)((ed.ncells + 255) / 256), 256, 0, ctx->stream>>>(k.col_ids, ed.ncells, plan->dir_cells.p, (unsigned long long *)cnt.p);
      check_launch(ctx, "dirichlet_cells_kernel");
      cnt.download(&plan->n_dir_cells, ctx->stream);
      GB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (plan->n_dir_cells > 0) {
      RhsArgs kl = k;
      kl.cell_list = plan->dir_cells.p;
      kl.ncells = plan->n_dir_cells;
      q1hex_rhs_kernel<true, false><<<(int)((plan->n_dir_cells + 127) / 128), 128, 0, ctx->stream>>>(kl);
      check_launch(ctx, "q1hex_rhs_kernel");
    }
  }
  return true;
}

}  // namespace gb
