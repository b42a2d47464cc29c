/* gridap_b200.h -- C ABI of libgridap_b200.so, the B200-native FE assembly engine that sits behind
 * Gridap's `SparseMatrixAssembler` / `AssemblyStrategy` interface.
 *
 * The reference (Gridap.jl v0.20.8) has no FFI on this path: the boundary is the Julia abstract type
 * `SparseMatrixAssembler` (src/FESpaces/SparseMatrixAssemblers.jl:4, src/FESpaces/Assemblers.jl:155-257).
 * A Julia subtype `B200SparseMatrixAssembler` overrides the methods listed beside each
 * entry point below and `ccall`s it.  Conventions on the wire are Julia's:
 *   - arrays are caller-owned, contiguous, column-major; the library never keeps a host pointer after return;
 *   - ids are 1-based; DoF ids are signed Int32 (free > 0, Dirichlet < 0) exactly as `get_cell_dof_ids` returns
 *     them (src/FESpaces/UnconstrainedFESpaces.jl:54-75); CSC arrays are Int64 (`SparseMatrixCSC{Float64,Int}`);
 *   - jagged arrays are `Table(data,ptrs)` pairs (src/Arrays/Tables.jl:21-28), ptrs 1-based of length n+1;
 *   - every function returns 0 on success and a negative `gb200_status` otherwise; `gb200_last_error` gives the text;
 *   - one host thread per context; calls are synchronous.
 * Unsupported integrands / spaces return GB200_ERR_UNSUPPORTED: there is no CPU fallback anywhere.
 */
#ifndef GRIDAP_B200_H
#define GRIDAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gb200_ctx_s *gb200_ctx;
typedef struct gb200_mesh_s *gb200_mesh;
typedef struct gb200_refel_s *gb200_refel;
typedef struct gb200_space_s *gb200_space;
typedef struct gb200_plan_s *gb200_plan;

typedef enum {
  GB200_OK = 0,
  GB200_ERR_INVALID = -1,     /* bad argument (the @check / @assert failures of the reference) */
  GB200_ERR_UNSUPPORTED = -2, /* integrand / element / space outside the supported set (@notimplemented) */
  GB200_ERR_CUDA = -3,        /* CUDA runtime error, no device, or out of memory */
  GB200_ERR_STATE = -4        /* call made in the wrong order (e.g. pattern not built) */
} gb200_status;

/* Cell types (geometry map = first-order Lagrangian, node order as Gridap: first axis fastest). */
typedef enum { GB200_QUAD4 = 1, GB200_HEX8 = 2, GB200_TRI3 = 3, GB200_TET4 = 4, GB200_SEG2 = 5 } gb200_celltype;
/* A mesh may be passed with D = (dimension of the cell type) + 1: the facets of a BoundaryTriangulation
 * (src/Geometry/BoundaryTriangulations.jl:152-167; QUAD4 / TRI3 in 3D, SEG2 in 2D).  The measure is then
 * sqrt(det(Jt.J)) (src/TensorValues/Operations.jl:991-1007) and only mass / source integrands are defined
 * (Neumann and Robin terms); reference elements of such a mesh are tabulated in the facet's D-1 reference coordinates. */

/* Supported integrands.  Matrix forms: */
typedef enum {
  GB200_FORM_NONE = 0,
  GB200_FORM_MASS = 1,           /* int u.v            (benchmark/bm/bm_assembly.jl:7)  params: {coef}            */
  GB200_FORM_LAPLACIAN = 2,      /* int grad(u):grad(v) (bm_assembly.jl:8)              params: {coef}            */
  GB200_FORM_ELASTICITY = 3,     /* int eps(v):(lambda tr(eps(u)) I + 2 mu eps(u))      params: {lambda, mu}      */
  GB200_FORM_STOKES = 4,         /* int grad(v):grad(u) - (div v) p + q (div u); fields (u,p) (StokesTaylorHoodTests.jl:59) */
  GB200_FORM_NEOHOOKEAN_JAC = 5, /* Jacobian of the neo-Hookean residual at u_h          params: {lambda, mu}      */
  /* vector forms: */
  GB200_FORM_SOURCE = 10,        /* int v.f, f constant (params[0..ncomp)) or given at quadrature points          */
  GB200_FORM_NEOHOOKEAN_RES = 11, /* neo-Hookean residual at u_h                          params: {lambda, mu}      */
  /* facet-of-cell plans only (gb200_plan_set_facets); kind 0 = value, 1 = normal derivative n.grad:
   * matrix  int_Gamma coef T(v) U(u) on equal components          params: {coef, test kind, trial kind}
   *         (Nitsche, PoissonTests.jl:99-101: (gamma/h) v u  {c,0,0};  - v (n.grad u)  {-1,0,1};  - (n.grad v) u  {-1,1,0})
   * vector  int_Gamma coef T(v) d                                 params: {coef, test kind, data kind}
   *         data kind 0: d = g at the quadrature points (fq), 1: d = u_h, 2: d = n.grad(u_h)  (u_h: gb200_plan_set_state) */
  GB200_FORM_FACET = 20,
  GB200_FORM_FACET_VEC = 21,
  /* ---- interior facets of a SkeletonTriangulation (gb200_plan_set_skeleton): plus / minus traces of the cell bases
   * (src/Geometry/SkeletonTriangulations.jl:7-32, SkeletonPair; jump / mean: src/CellData/CellFields.jl, `jump(a) = a.plus - a.minus`,
   * `mean(a) = 0.5 (a.plus + a.minus)`, `jump(v n) = v+ n+ + v- n-`).
   * matrix  int_Lambda coef [w+ T(v+) + w- T(v-)] [z+ U(u+) + z- U(u-)]      params: {coef, T kind, w+, w-, U kind, z+, z-}
   *         kind 0: value, 1: derivative along the PLUS normal n+ (n- = -n+)
   *         (DG Poisson, test/GridapTests/PoissonDGTests.jl:42-45:  (gamma/h) jump(v n).jump(u n)  {c,0,1,-1,0,1,-1};
   *          - jump(v n).mean(grad u)  {-1,0,1,-1,1,.5,.5};   - mean(grad v).jump(u n)  {-1,1,.5,.5,0,1,-1}) */
  GB200_FORM_SKELETON = 22
} gb200_form;

/* Context flags */
#define GB200_FLAG_DETERMINISTIC 1u /* atomic-free scatter (owner-computes gather or cell colouring) */

/* ---- context -------------------------------------------------------------------------------------- */
int32_t gb200_init(int32_t device, uint32_t flags, gb200_ctx *ctx);
int32_t gb200_finalize(gb200_ctx ctx);
/* Text of the last error raised on `ctx` (or the last global error if ctx == NULL). Never NULL. */
const char *gb200_last_error(gb200_ctx ctx);
/* Library version / build info ("gridap_b200 <ver> sm_100a ..."). */
const char *gb200_version(void);
/* JSON with the device time (ms, CUDA events) of the kernels / copies of the last call, written into buf. */
int32_t gb200_get_timings(gb200_ctx ctx, char *buf, size_t len);
/* Page-locked host memory for the caller's arrays (H2D / D2H at full PCIe rate).  gb200_host_alloc gives a buffer the
 * host language can wrap (Julia: unsafe_wrap); gb200_host_register pins an existing array in place. */
int32_t gb200_host_alloc(gb200_ctx ctx, size_t bytes, void **p);
int32_t gb200_host_free(gb200_ctx ctx, void *p);
int32_t gb200_host_register(gb200_ctx ctx, void *p, size_t bytes);
int32_t gb200_host_unregister(gb200_ctx ctx, void *p);
/* Device blocks freed by the library stay in a per-stream cache so that repeated assemblies do not pay cudaMalloc /
 * cudaFree for the multi-GB transients of the symbolic phase; gb200_trim (and an out-of-memory allocation) returns them
 * to the driver. */
int32_t gb200_trim(gb200_ctx ctx);
/* Number of kernel launches issued by the library on this context since init (bench.py `gpu_launches`). */
int64_t gb200_launch_count(gb200_ctx ctx);
/* The CUDA stream (cudaStream_t) all work of this context is enqueued on. */
void *gb200_stream(gb200_ctx ctx);
int32_t gb200_synchronize(gb200_ctx ctx);

/* ---- geometry: get_node_coordinates / get_cell_node_ids (src/Geometry/UnstructuredGrids.jl:10-17) ----
 * coords: f64[D*nnodes], node-major (a Julia Vector{Point{D,Float64}});
 * cell_node_data/ptrs: Table{Int32}, all cells of type `celltype`. */
int32_t gb200_mesh_create(gb200_ctx ctx, int32_t D, int64_t nnodes, const double *coords, int64_t ncells,
                          const int32_t *cell_node_data, const int32_t *cell_node_ptrs, int32_t celltype,
                          gb200_mesh *mesh);
int32_t gb200_mesh_destroy(gb200_mesh mesh);
/* 1 if every cell map is affine (Jacobian constant per cell up to round-off), else 0. */
int32_t gb200_mesh_is_affine(gb200_mesh mesh, int32_t *is_affine);

/* ---- reference element tabulated at the quadrature points (host, once per reference element:
 * get_shapefuns / Quadrature, src/ReferenceFEs/ReferenceFEInterfaces.jl:563-583, Quadratures.jl:156-176).
 * Scalar Lagrangian shape functions; a vector-valued space has local DoF k = a + nd*(c-1)
 * (src/ReferenceFEs/LagrangianDofBases.jl:77-96).
 *   w  f64[np];  N f64[np*nd] (N[p + np*a], i.e. a Julia Matrix [np,nd]);  dN f64[D*np*nd] (dN[d + D*(p + np*a)],
 *   a Julia Matrix{VectorValue{D}} [np,nd]).
 * Limits: 1 <= D <= 3, 1 <= np <= 64 (the order-3 mass rule on a hexahedron: 4^3 points), 1 <= ncomp <= 3, nd free (any order:
 * the orders 1-3 of benchmark/bm/bm_assembly.jl are covered by tests/test_gpu_bm_protocol.py); elements without a dedicated kernel
 * run on the size-generic element kernel, whose per-cell scratch must fit 200 KB of shared memory (GB200_ERR_UNSUPPORTED else). */
int32_t gb200_refel_create(gb200_ctx ctx, int32_t D, int32_t np, int32_t nd, int32_t ncomp, const double *w,
                           const double *N, const double *dN, gb200_refel *refel);
int32_t gb200_refel_destroy(gb200_refel refel);

/* ---- FE space: get_cell_dof_ids (signed), num_free_dofs, num_dirichlet_dofs --------------------------- */
int32_t gb200_space_create(gb200_ctx ctx, gb200_mesh mesh, gb200_refel refel, const int32_t *cell_dof_data,
                           const int32_t *cell_dof_ptrs, int64_t nfree, int64_t ndirichlet, gb200_space *space);
int32_t gb200_space_destroy(gb200_space space);

/* ---- plan = symbolic phase (nz_counter -> symbolic_loop_matrix! -> nz_allocation -> create_from_nz,
 * src/FESpaces/SparseMatrixAssemblers.jl:51-58,174-210; src/Algebra/SparseMatrixCSC.jl:72-283) --------------
 * geo: reference element of the geometry map tabulated at the same quadrature points (nd = nodes per cell).
 * test/trial spaces: one per field (single field: ntest = ntrial = 1).  touched: u8[ntest*ntrial], column-major
 * (touched[bi + ntest*bj]); untouched blocks are absent from the pattern (src/Fields/FieldArrayBlocks.jl:488).
 * row/col_offsets: Int64 per field, added to positive ids (ConsecutiveMultiFieldStyle,
 * src/MultiField/MultiFieldFESpaces.jl:356-364).  nrows/ncols = size of the global system.
 * The pattern it builds equals the reference's bit for bit (canonical CSC: rows ascending, unique per column). */
int32_t gb200_plan_create(gb200_ctx ctx, gb200_mesh mesh, gb200_refel geo, int32_t ntest, const gb200_space *test_spaces,
                          int32_t ntrial, const gb200_space *trial_spaces, const uint8_t *touched,
                          const int64_t *row_offsets, const int64_t *col_offsets, int64_t nrows, int64_t ncols,
                          gb200_plan *plan);
int32_t gb200_plan_destroy(gb200_plan plan);
/* Facet-of-cell plans: terms on a BoundaryTriangulation that need the adjacent cell -- the unit normal (get_facet_normal,
 * src/Geometry/BoundaryTriangulations.jl:244-283, push_normal :310-318) and the cell basis / its gradient at the facet quadrature
 * points (FaceToCellGlue :13-70, compute_face_to_cell_reference_map :320-340), e.g. the Nitsche terms of
 * test/GridapTests/PoissonTests.jl:99-107.  The plan is created on the mesh of the cells ADJACENT to the facets (one "cell" per
 * facet) with the cell dof tables of those cells; every tabulation passed to gb200_refel_create holds nlfaces blocks of npf points:
 * block lf = the facet rule mapped onto local face lf of the reference cell (w repeated per block).  lface i32[ncells]: 1-based local
 * face of every facet; nref f64[D*nlfaces] (nref[d + D*lf]): outward reference normal of local face lf scaled by the ratio of the
 * reference measures (face of the reference cell / facet reference polytope: 1 except for the oblique faces of simplices).
 * Afterwards the plan has npf quadrature points per facet (fq arrays, gb200_quadrature_points). */
int32_t gb200_plan_set_facets(gb200_plan plan, const int32_t *lface, int32_t nlfaces, const double *nref);
/* Skeleton plans: terms on the interior facets (SkeletonTriangulation, src/Geometry/SkeletonTriangulations.jl:7-32,54-99).  The plan
 * has TWO "fields": field 0 = the FE space on the PLUS cells of the facets, field 1 = the same FE space on the MINUS cells (spaces
 * created on two meshes with one "cell" per facet: the plus / minus cell; both row / column offsets 0, all four blocks touched), so
 * that a facet's local matrix is the 2x2 block matrix [plus, minus] x [plus, minus] of the reference (BlockMap over SkeletonPair) and
 * the symbolic phase yields the union pattern of the cross couplings.  Tabulations as for gb200_plan_set_facets (one block of npf
 * points per local face).  lface_plus / lface_minus i32[nfacets]: 1-based local face of the facet in its plus / minus cell;
 * perm i32[npf*nfacets] (perm[p + npf*facet], 0-based): the point of the minus cell's local-face block that coincides with point p
 * of the plus side (the vertex permutation of FaceToCellGlue, cell_to_lface_to_pindex, :42-70, applied to the quadrature points). */
int32_t gb200_plan_set_skeleton(gb200_plan plan, const int32_t *lface_plus, const int32_t *lface_minus, const int32_t *perm,
                                int32_t nlfaces, const double *nref);
int32_t gb200_plan_nnz(gb200_plan plan, int64_t *nnz);
/* colptr Int64[ncols+1], rowval Int64[nnz], 1-based: the arrays of the SparseMatrixCSC `allocate_matrix` returns. */
int32_t gb200_plan_get_pattern(gb200_plan plan, int64_t *colptr, int64_t *rowval);
/* Same, but returns once the copy is enqueued (on a second stream): it completes inside the next call on this context that
 * synchronises (any assemble call given a host array, gb200_plan_download, gb200_synchronize).  For the
 * allocate_matrix + assemble_matrix! pair inside assemble_matrix (src/FESpaces/SparseMatrixAssemblers.jl:70-77): the 3.7 GB
 * pattern download of the 256^3 problem then overlaps the numeric phase.  colptr / rowval must stay valid until then. */
int32_t gb200_plan_get_pattern_async(gb200_plan plan, int64_t *colptr, int64_t *rowval);
/* Free (>0) and Dirichlet (<0) values of the FE function u_h used by residual / Jacobian forms and of the
 * Dirichlet lifting (PosNegReindex, src/FESpaces/UnconstrainedFESpaces.jl:65-75).  NULL => zeros. */
int32_t gb200_plan_set_state(gb200_plan plan, int32_t field, const double *free_values, const double *dirichlet_values);
/* Same, from DEVICE arrays (copied device-to-device on the context stream, no host round trip): the Newton update of a solver that
 * keeps the unknown on the GPU (src/Algebra/NLSolvers.jl:34-77 with a device linear solver).  Either pointer may be NULL: that vector is
 * left UNCHANGED (unlike gb200_plan_set_state, where NULL means zeros) -- a Newton loop sets the Dirichlet values once and then only
 * updates the free values.  gb200_plan_set_state / gb200_plan_set_state_space reset both. */
int32_t gb200_plan_set_state_device(gb200_plan plan, int32_t field, const double *d_free_values, const double *d_dirichlet_values);
/* The FE function u_h of a residual / Jacobian form lives on the GLOBAL trial space (EvaluationFunction(trial, x),
 * src/FESpaces/FEOperators.jl:154-176).  A plan whose trial ids are masked / renumbered (owned-column plans of the multi-GPU
 * path, AssemblyStrategy col_map / col_mask, src/FESpaces/Assemblers.jl:31-55) gathers u_h through the ids of `space` instead
 * (same mesh, same local DoF layout, unmasked global ids); gb200_plan_set_state then takes that space's full free / Dirichlet
 * vectors.  NULL restores the trial space.  Clears the current state values. */
int32_t gb200_plan_set_state_space(gb200_plan plan, int32_t field, gb200_space space);

/* ---- numeric phase ----------------------------------------------------------------------------------
 * nzval f64[nnz] / b f64[nrows] are host arrays; NULL keeps the result on the device only
 * (see gb200_plan_device_* below).  add_flag: 0 = assemble_*! (fillstored!/fill! first), 1 = assemble_*_add!
 * (src/FESpaces/SparseMatrixAssemblers.jl:32-40,60-68,88-97).  With add_flag=1 and a host array, the host
 * values are the starting point (uploaded first). */
int32_t gb200_assemble_matrix(gb200_plan plan, int32_t form, const double *params, int32_t nparams, double *nzval,
                              int32_t add_flag);
/* Every cell has the same local matrix Ke f64[ni*nj] column-major: the Fill(K_e,ncells) case of a
 * CartesianDiscreteModel (src/Arrays/LazyArrays.jl:302-322, src/Geometry/CartesianGrids.jl:271-276); scatter only. */
int32_t gb200_assemble_matrix_const(gb200_plan plan, const double *Ke, double *nzval, int32_t add_flag);
/* fq: f64[ncomp*np*ncells] values of f at the physical quadrature points (fq[c + ncomp*(p + np*cell)]) or NULL
 * (constant f = params[0..ncomp)). */
int32_t gb200_assemble_vector(gb200_plan plan, int32_t form, const double *params, int32_t nparams, const double *fq,
                              double *b, int32_t add_flag);
/* Fused a13+a16 with Dirichlet lifting b_e -= K_e u_e on Dirichlet cells (src/CellData/AttachDirichlet.jl:76-84,
 * src/FESpaces/SparseMatrixAssemblers.jl:365-405); Dirichlet values come from gb200_plan_set_state.
 * mat_params parameterise form_mat, vec_params form_vec (e.g. the constant source f). */
int32_t gb200_assemble_matrix_and_vector(gb200_plan plan, int32_t form_mat, const double *mat_params, int32_t nmat,
                                         int32_t form_vec, const double *vec_params, int32_t nvec, const double *fq,
                                         double *nzval, double *b, int32_t add_flag);
/* Physical quadrature points xq f64[D*np*ncells] (xq[d + D*(p + np*cell)]) so the host can evaluate f(x). */
int32_t gb200_quadrature_points(gb200_plan plan, double *xq);

/* ---- device-resident results (hand-off to a GPU solver, multi-GPU exchange, roofline timing) --------
 * The arrays are written on the context's stream (gb200_stream): call gb200_synchronize (or make the consumer's stream wait on
 * it) after a device-resident assembly before reading them, and let the consumer finish before the next assembly call on the
 * same plan overwrites them. */
int32_t gb200_plan_device_nzval(gb200_plan plan, void **dptr, int64_t *nnz);
/* The pattern as it lives on the device: colptr Int64[ncols+1] and rowval Int32[nnz], 0-based, rows ascending inside a column --
 * what a device-side consumer (cuSPARSE / AmgX-style solver replacing LUSolver, src/Algebra/LinearSolvers.jl; a SpMV) binds
 * together with gb200_plan_device_nzval, without any download.  The arrays belong to the plan. */
int32_t gb200_plan_device_pattern(gb200_plan plan, void **colptr, void **rowval);
int32_t gb200_plan_device_vector(gb200_plan plan, void **dptr, int64_t *nrows);
int32_t gb200_plan_download(gb200_plan plan, double *nzval, double *b);
/* ---- forms over several triangulations (a = int_Omega ... + int_Gamma ..., src/FESpaces/SparseMatrixAssemblers.jl:223-236: one
 * numeric loop per triangulation into the same matrix).  `src` is the plan of another triangulation of the same global system
 * whose pattern is contained in `dst`'s (boundary facets inside bulk cells): every stored value of src is added at the
 * slot of the same (row, column) in dst's device matrix.  GB200_ERR_INVALID if an entry of src is not in dst's pattern. */
int32_t gb200_plan_add_matrix_from(gb200_plan dst, gb200_plan src);
/* ---- linear constraints (FESpaceWithLinearConstraints, src/FESpaces/FESpacesWithLinearConstraints.jl:40-120; the reference applies
 * the cell-wise constraint matrices to every cell matrix / vector before the scatter: attach_constraints_rows / _cols,
 * src/FESpaces/FESpaceInterface.jl:361-387).  Summed over the cells that is  A_c = T^T A T,  b_c = T^T b - (T^T A T)[:, Dirichlet
 * masters] u_D  with the global constraint table T, applied here to the ASSEMBLED arrays: `src` is the plan of the unconstrained
 * space with its free and Dirichlet DoFs in one positive numbering (DOF = dof > 0 ? dof : nfree - dof, :356-372) -- every fast cell
 * kernel runs unchanged on it -- and `dst` is a plan whose pattern is that of the constrained space (cell tables of master DoFs,
 * get_cell_dof_ids(::FESpaceWithLinearConstraints) :280-283).  dof_ptrs Int64[nDOFs+1] (1-based), dof_mdofs i32 (signed master ids:
 * > 0 free master, < 0 Dirichlet master), dof_coeffs f64: the tables DOF_to_mDOFs / DOF_to_coeffs of the reference.
 * dirichlet_master_values f64[ndirichlet_masters] or NULL (no lifting).  with_matrix / with_vector: which device arrays of dst are
 * overwritten. */
/* The plan's device vector := b (host, f64[nrows]): a vector accumulated over several triangulations on the host side of the binding,
 * handed back for a device-side operation (gb200_plan_fold_constraints). */
int32_t gb200_plan_upload_vector(gb200_plan plan, const double *b);
int32_t gb200_plan_fold_constraints(gb200_plan dst, gb200_plan src, const int64_t *dof_ptrs, const int32_t *dof_mdofs, const double *dof_coeffs,
                                    const double *dirichlet_master_values, int64_t ndirichlet_masters, int32_t with_matrix, int32_t with_vector);
/* ---- SparseMatrixCSR{Bi,Float64,Int} output (src/Algebra/SparseMatrixCSR.jl:31-75; SparseMatricesCSR.jl): rowptr Int64[nrows+1],
 * colval Int64[nnz] with index base Bi in {0,1}, columns ascending inside a row -- what `create_from_nz(::NzAllocationCSR)` returns
 * (the CSC of the transpose, transposed).  The CSR view is derived once per plan on the device (count, scan, fill, per-row sort);
 * gb200_plan_download_csr copies the current values in CSR order (after any gb200_assemble_* with nzval == NULL). */
int32_t gb200_plan_get_csr_pattern(gb200_plan plan, int32_t index_base, int64_t *rowptr, int64_t *colval);
int32_t gb200_plan_download_csr(gb200_plan plan, double *nzval);
/* ---- BlockMultiFieldStyle (src/MultiField/BlockSparseMatrixAssemblers.jl:19-33,197-230; nz_counter / create_from_nz per block
 * at :197-230): the matrix as a BlockMatrix with one SparseMatrixCSC per field block (bi, bj), block-local 1-based ids.  The
 * blocks are views of the plan's single device matrix; an untouched block comes back empty (nnz 0).
 * gb200_plan_block_nnz -> allocate; gb200_plan_get_block_pattern fills colptr Int64[ncols_bj + 1], rowval Int64[nnz];
 * gb200_plan_download_block copies the block's current values (after any gb200_assemble_* with nzval == NULL). */
int32_t gb200_plan_block_nnz(gb200_plan plan, int32_t bi, int32_t bj, int64_t *nnz);
int32_t gb200_plan_get_block_pattern(gb200_plan plan, int32_t bi, int32_t bj, int64_t *colptr, int64_t *rowval);
int32_t gb200_plan_download_block(gb200_plan plan, int32_t bi, int32_t bj, double *nzval);
/* Name of the kernel path chosen for (form) on this plan, e.g. "q1hex_gather_affine", "generic_atomic". */
const char *gb200_plan_kernel_path(gb200_plan plan, int32_t form);

/* ---- multi-GPU (one context per GPU, one process per GPU) -------------------------------------------------
 * Cells are partitioned over ranks; each rank owns a contiguous range of CSC columns and assembles exactly
 * those columns from its own cells plus the ghost cells that touch them (owner computes, no exchange) -- the
 * column-mask of Gridap's AssemblyStrategy (src/FESpaces/Assemblers.jl:31-55).  On the wire this needs nothing
 * new: the host passes trial ids renumbered to the local column range with id 0 for masked (non-owned) columns;
 * id 0 is skipped everywhere (neither free nor Dirichlet).  The rank's result is its column slab of the global
 * CSC (global row ids), so the global colptr/rowval/nzval are the concatenation over ranks. */

/* Host helper for that contract (no device work, ctx may be NULL): applies a column ownership to a trial DoF table.
 * ids i32[n]: signed cell DoF ids (global numbering, multi-field offsets already added); owned u8[nfree]: 1 where this rank owns the
 * column; out i32[n]: owned ids renumbered 1..n_owned in ascending global order, other positive ids 0 (masked), negative ids
 * (Dirichlet) unchanged -- exactly `map_cols!` of an AssemblyStrategy whose col_mask is `owned` and whose col_map is the rank-local
 * numbering (src/FESpaces/Assemblers.jl:45-55).  n_owned returns the number of owned columns (= ncols of the rank's plan); the
 * rank's local column j is global column owned_ids[j] (optional output, Int64[n_owned], 1-based, may be NULL). */
int32_t gb200_owned_column_ids(const int32_t *ids, int64_t n, const uint8_t *owned, int64_t nfree, int32_t *out, int64_t *n_owned,
                               int64_t *owned_ids);

#ifdef __cplusplus
}
#endif
#endif /* GRIDAP_B200_H */
