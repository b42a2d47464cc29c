"""GPU parity through the host mirror of the Gridap API (reads like test/FESpacesTests/SparseMatrixAssemblersTests.jl)."""
import numpy as np
import pytest

import gridap_b200 as g
from gridap_b200 import lib
from oracle import capi, problems

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def test_sparse_matrix_assemblers_golden():
    # test/FESpacesTests/SparseMatrixAssemblersTests.jl:16-40,104-152
    model = g.CartesianDiscreteModel((0, 1, 0, 1), (2, 2))
    reffe = g.ReferenceFE(g.lagrangian, float, 1)
    V = g.FESpace(model, reffe, dirichlet_tags=[1, 2, 3, 4, 6, 5])
    U = V
    dO = g.Measure(g.get_triangulation(model), 2)
    v, u = g.get_fe_basis(V), g.get_trial_fe_basis(U)
    matdata = g.collect_cell_matrix(U, V, g.Integral(g.inner(g.grad(v), g.grad(u))) * dO)
    vecdata = g.collect_cell_vector(V, g.Integral(g.inner(v, lambda x: x[:, 1])) * dO)
    assem = g.SparseMatrixAssembler(U, V)
    mat = assem.assemble_matrix(matdata)
    vec = assem.assemble_vector(vecdata)
    x = np.linalg.solve(mat.toarray(), vec)
    assem.assemble_matrix_(mat, matdata)
    assem.assemble_vector_(vec, vecdata)
    assert np.allclose(np.linalg.solve(mat.toarray(), vec), x)
    assert np.allclose(vec, [0.0625, 0.125, 0.0625], rtol=0, atol=1e-14)
    assert abs(mat.getindex(1, 1) - 1.333333333333333) < 1e-13
    assert abs(mat.getindex(2, 1) + 0.33333333333333) < 1e-13
    assert abs(mat.getindex(1, 2) + 0.33333333333333) < 1e-13
    assert abs(mat.getindex(2, 2) - 2.666666666666666) < 1e-13
    assert abs(mat.getindex(3, 2) + 0.33333333333333) < 1e-13
    assert abs(mat.getindex(2, 3) + 0.33333333333333) < 1e-13
    assert abs(mat.getindex(3, 3) - 1.333333333333333) < 1e-13
    data = g.collect_cell_matrix_and_vector(U, V, g.Integral(g.inner(g.grad(v), g.grad(u))) * dO, g.Integral(g.inner(v, lambda x: x[:, 1])) * dO,
                                            g.zero(U))
    mat2, vec2 = assem.allocate_matrix_and_vector(data)
    assem.assemble_matrix_and_vector_(mat2, vec2, data)
    assem.assemble_matrix_and_vector_(mat2, vec2, data)
    assert np.allclose(vec2, [0.0625, 0.125, 0.0625], rtol=0, atol=1e-14) and abs(mat2.getindex(1, 1) - 1.333333333333333) < 1e-13
    # _add! accumulates
    assem.assemble_matrix_add_(mat2, matdata)
    assert abs(mat2.getindex(1, 1) - 2 * 1.333333333333333) < 1e-13


def test_config1_poisson_2d_100x100():
    # BASELINE.json configs[0]: 2D Poisson Q1 100x100, assemble_matrix + assemble_vector
    n = 100
    model = g.CartesianDiscreteModel((0, 1, 0, 1), (n, n))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags="boundary")
    U = g.TrialFESpace(V, 0.0)
    dO = g.Measure(g.Triangulation(model), 2)
    A = g.assemble_matrix(lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dO, U, V)
    b = g.assemble_vector(lambda v: g.Integral(v * 1.0) * dO, V)
    assert A.nnz() == (3 * 99 - 2) ** 2 and A.shape == (9801, 9801)
    pb = problems.single_field_problem((0, 1, 0, 1), (n, n), form_mat=capi.LAPLACIAN, form_vec=capi.SOURCE, params=[1.0])
    colptr, rowval, nzval, bo = pb.assemble(with_vector=True)
    assert np.array_equal(A.colptr, colptr) and np.array_equal(A.rowval, rowval)
    assert relerr(A.nzval, nzval) <= 1e-12 and relerr(b, bo) <= 1e-12


@pytest.mark.parametrize("n", [8, 20, 40])  # 40: some 32-column blocks are canonical 3x3x3 stencils (register path)
def test_config2_poisson_3d_q1(n):
    model = g.UnstructuredDiscreteModel(g.CartesianDiscreteModel((0, 1) * 3, (n, n, n)))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags="boundary")
    U = g.TrialFESpace(V, lambda x: np.sin(x[:, 0]) + x[:, 2])
    dO = g.Measure(g.Triangulation(model), 2)
    a = lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dO  # noqa: E731
    l = lambda v: g.Integral(v * (lambda x: x[:, 0] * x[:, 1])) * dO  # noqa: E731
    assem = g.SparseMatrixAssembler(U, V)
    op = g.AffineFEOperator(a, l, U, V, assem)
    A, b = op.get_matrix(), op.get_vector()
    assert assem.plan(dO).kernel_path(lib.FORM_LAPLACIAN) == "q1hex_gather_affine+diag"   # axis-aligned cells: diagonal metric
    pb0 = problems.single_field_problem((0, 1) * 3, (n, n, n), form_mat=capi.LAPLACIAN, form_vec=capi.SOURCE)
    xq = pb0.quadrature_points()
    pb = problems.single_field_problem((0, 1) * 3, (n, n, n), form_mat=capi.LAPLACIAN, form_vec=capi.SOURCE, fq=xq[:, :, 0] * xq[:, :, 1],
                                       dirichlet_values=U.dirichlet_values, lift=True)
    colptr, rowval, nzval, bo = pb.assemble(with_vector=True)
    assert A.nnz() == (3 * (n - 1) - 2) ** 3
    assert np.array_equal(A.colptr, colptr) and np.array_equal(A.rowval, rowval)
    assert relerr(A.nzval, nzval) <= 1e-12 and relerr(b, bo) <= 1e-12
    # mass through the same assembler (second form on the same plan)
    M = g.assemble_matrix(lambda u, v: g.Integral(u * v) * dO, assem, U, V)
    pbm = problems.single_field_problem((0, 1) * 3, (n, n, n), form_mat=capi.MASS)
    assert relerr(M.nzval, pbm.assemble()[2]) <= 1e-12


def test_deterministic_is_reproducible_and_matches_atomic():
    n = 10
    model = g.CartesianDiscreteModel((0, 1) * 3, (n, n, n))
    X = model.node_coordinates
    rng = np.random.default_rng(12345)
    inner = np.all((X > 1e-9) & (X < 1 - 1e-9), axis=1)
    X[inner] += 0.2 / n * rng.uniform(-1, 1, size=(inner.sum(), 3))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags="boundary")
    dO = g.Measure(g.Triangulation(model), 2)
    a = lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dO  # noqa: E731
    det = g.SparseMatrixAssembler(V, V, deterministic=True)
    A1 = g.assemble_matrix(a, det, V, V)
    A2 = g.assemble_matrix(a, det, V, V)
    assert det.plan(dO).kernel_path(lib.FORM_LAPLACIAN) == "q1hex_gather_general"  # owner-computes: deterministic by construction
    assert np.array_equal(A1.nzval, A2.nzval)  # bitwise self-reproducible
    A3 = g.assemble_matrix(a, g.SparseMatrixAssembler(V, V), V, V)
    assert np.array_equal(A3.nzval, A1.nzval)
    # vector-valued field on the same perturbed mesh: coloured (deterministic) vs atomic scatter of the node-pair kernel
    W = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 1), dirichlet_tags="boundary")
    av = lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dO  # noqa: E731
    detw = g.SparseMatrixAssembler(W, W, deterministic=True)
    B1 = g.assemble_matrix(av, detw, W, W)
    B2 = g.assemble_matrix(av, detw, W, W)
    assert detw.plan(dO).kernel_path(lib.FORM_LAPLACIAN) == "staged_gather+blocks"   # non-affine cells: staged node-pair blocks + owner gather
    assert np.array_equal(B1.nzval, B2.nzval)
    from parity_helpers import env
    with env(GB200_NO_STAGED_GATHER=1):
        detc = g.SparseMatrixAssembler(W, W, deterministic=True)
        B4 = g.assemble_matrix(av, detc, W, W)
        B5 = g.assemble_matrix(av, detc, W, W)
        assert detc.plan(dO).kernel_path(lib.FORM_LAPLACIAN) == "vector_coloured"
    assert np.array_equal(B4.nzval, B5.nzval)
    assert relerr(B4.nzval, B1.nzval) <= 1e-13
    B3 = g.assemble_matrix(av, g.SparseMatrixAssembler(W, W), W, W)
    assert relerr(B3.nzval, B1.nzval) <= 1e-13


def test_fill_local_matrix_scatter_only():
    # CartesianDiscreteModel + constant coefficients: the cell-matrix array is Fill(K_e) (test/GeometryTests/CartesianGridsTests.jl:98)
    n = 6
    model = g.CartesianDiscreteModel((0, 1) * 3, (n, n, n))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags="boundary")
    dO = g.Measure(g.Triangulation(model), 2)
    pb = problems.single_field_problem((0, 1) * 3, (n, n, n), form_mat=capi.LAPLACIAN)
    Ke = pb.cell_local(0)[0][0][0]
    assem = g.SparseMatrixAssembler(V, V)
    A = assem.assemble_matrix(g.fill_cell_matrix(Ke, dO))
    colptr, rowval, nzval = capi.assemble_const(V.cell_dof_ids, Ke, V.nfree, V.nfree)
    assert np.array_equal(A.colptr, colptr) and np.array_equal(A.rowval, rowval)
    assert relerr(A.nzval, nzval) <= 1e-13


def test_unsupported_integrand_raises():
    model = g.CartesianDiscreteModel((0, 1, 0, 1), (3, 3))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1))
    dO = g.Measure(g.Triangulation(model), 2)
    with pytest.raises(NotImplementedError):
        g.assemble_matrix(lambda u, v: g.Integral(g.inner(g.grad(v), u)) * dO, V, V)
    with pytest.raises(NotImplementedError):
        g.ReferenceFE("raviart_thomas", float, 1)
    with pytest.raises(NotImplementedError):
        g.FESpace(model, g.ReferenceFE(g.lagrangian, float, 1), constraint="periodic")


@pytest.mark.parametrize("bi", [0, 1])
def test_sparse_matrix_csr_output(bi):
    # SparseMatrixAssembler(SparseMatrixCSR{Bi,Float64,Int}, Vector{Float64}, U, V): src/Algebra/SparseMatrixCSR.jl:31-75
    n = 7
    model = g.CartesianDiscreteModel((0, 1) * 3, (n, n - 1, n - 2))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags=[21, 22])     # faces z = 0, z = 1 (interiors)
    W = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags="boundary")   # different trial space: non-square
    U = g.TrialFESpace(W, lambda x: x[:, 0])
    dO = g.Measure(g.Triangulation(model), 2)
    a = lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dO  # noqa: E731
    A = g.assemble_matrix(a, g.SparseMatrixAssembler(U, V), U, V)
    assem = g.SparseMatrixAssembler(g.SparseMatrixCSR[bi], np.ndarray, U, V)
    R = g.assemble_matrix(a, assem, U, V)
    assert isinstance(R, g.SparseMatrixCSR) and R.Bi == bi and R.shape == A.shape and R.nnz() == A.nnz()
    ref = A.to_scipy().tocsr()
    ref.sort_indices()
    assert np.array_equal(R.rowptr - bi, ref.indptr) and np.array_equal(R.colval - bi, ref.indices)   # pattern of transpose(CSC of A^T)
    assert np.array_equal(R.nzval, ref.data)       # owner-computes path: deterministic, so the values are the same numbers
    # AffineFEOperator on the CSR assembler: same vector, same matrix
    l = lambda v: g.Integral(v * 1.0) * dO  # noqa: E731
    op = g.AffineFEOperator(a, l, U, V, assem)
    op0 = g.AffineFEOperator(a, l, U, V)
    assert np.array_equal(op.get_matrix().nzval, ref.data) and relerr(op.get_vector(), op0.get_vector()) <= 1e-13
    # bulk + Robin boundary term, delivered in CSR order
    U2 = g.TrialFESpace(V, 0.0)                                        # square system: the face y = 0 carries free DoFs
    dG = g.Measure(g.BoundaryTriangulation(model, tags=[23]), 2)
    ar = lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dO + g.Integral(3.0 * (u * v)) * dG  # noqa: E731
    Rr = g.assemble_matrix(ar, g.SparseMatrixAssembler(g.SparseMatrixCSR[bi], np.ndarray, U2, V), U2, V)
    Ar = g.assemble_matrix(ar, U2, V).to_scipy().tocsr()
    Ar.sort_indices()
    A0 = g.assemble_matrix(a, U2, V).to_scipy().tocsr()
    A0.sort_indices()
    assert np.array_equal(Rr.colval - bi, Ar.indices) and relerr(Rr.nzval, Ar.data) <= 1e-13 and np.abs(Ar.data - A0.data).max() > 0


@pytest.mark.parametrize("ncomp", [1, 3])
def test_column_slab_partition_on_the_device(ncomp):
    # the N-GPU path of bench.py (one assembler per rank with its owned column range), run rank after rank on one GPU:
    # the slabs concatenate to the single-GPU matrix, the owned rows of the vectors to the single-GPU vector
    from gridap_b200 import distributed as gd
    n, world = 9, 3
    model = g.UnstructuredDiscreteModel(g.CartesianDiscreteModel((0, 1) * 3, (n, n, n - 2)))
    T = float if ncomp == 1 else g.VectorValue(3)
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, T, 1), dirichlet_tags="boundary")
    U = g.TrialFESpace(V, (lambda x: x[:, 0] + 2 * x[:, 1]) if ncomp == 1 else (lambda x: np.stack([x[:, 0], x[:, 1] ** 2, x[:, 2] + 1.0], axis=1)))
    dO = g.Measure(g.Triangulation(model), 2)
    if ncomp == 1:
        a = lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dO  # noqa: E731
        l = lambda v: g.Integral(v * 1.0) * dO  # noqa: E731
    else:
        sigma = g.IsotropicLinearElasticity(2.0, 1.0)
        a = lambda u, v: g.Integral(g.inner(g.eps(v), sigma(g.eps(u)))) * dO  # noqa: E731
        l = lambda v: g.Integral(g.inner(v, (1.0, 0.5, -2.0))) * dO  # noqa: E731
    A = g.assemble_matrix(a, U, V)
    b = g.assemble_vector(l, V)
    slabs, vecs, ranges = [], [], []
    for rank in range(world):
        part = gd.slab_partition(model, V, world, rank)
        asm = part.assembler(U, V)
        Ar = g.assemble_matrix(a, asm, U, V)
        assert Ar.shape == (V.num_free_dofs(), part.col_range[1] - part.col_range[0])
        slabs.append((Ar.colptr, Ar.rowval, Ar.nzval))
        vecs.append(g.assemble_vector(l, asm, V))
        ranges.append(part.col_range)
    G = gd.gather_csc(slabs, V.num_free_dofs())
    assert np.array_equal(G.colptr, A.colptr) and np.array_equal(G.rowval, A.rowval)
    if ncomp == 1:
        assert np.array_equal(G.nzval, A.nzval)   # owner-computes gather: bitwise
    else:
        assert relerr(G.nzval, A.nzval) <= 1e-13  # atomic scatter: summation order
    assert relerr(gd.gather_vector(vecs, ranges), b) <= 1e-13


@pytest.mark.parametrize("order", [1, 2])
def test_neumann_and_robin_boundary_terms(order):
    # a(u,v) = int_Omega grad v.grad u + int_Gamma 2.5 u v,  l(v) = int_Omega v f + int_Gamma v g  (Neumann / Robin terms on a
    # BoundaryTriangulation, test/GridapTests/PoissonTests.jl:101-108), inhomogeneous Dirichlet data on the face z = 0
    from test_host_logic import _facet_problem
    part = (4, 3, 3)
    model = g.CartesianDiscreteModel((0, 1) * 3, part)
    X = model.node_coordinates
    rng = np.random.default_rng(3)
    inner = np.all((X > 1e-9) & (X < 1 - 1e-9), axis=1)
    X[inner] += 0.03 * rng.uniform(-1, 1, size=(int(inner.sum()), 3))
    top = np.isclose(X[:, 2], 1.0) & (X[:, 0] > 1e-9) & (X[:, 0] < 1 - 1e-9) & (X[:, 1] > 1e-9) & (X[:, 1] < 1 - 1e-9)
    X[top, :2] += 0.03 * rng.uniform(-1, 1, size=(int(top.sum()), 2))     # non-affine facets on the Robin face
    dtags = [21, 1, 2, 3, 4, 9, 10, 13, 14]                                # face z = 0 with its edges and corners
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, order), dirichlet_tags=dtags)
    U = g.TrialFESpace(V, lambda x: 1.0 + x[:, 0] * x[:, 1])
    dO = g.Measure(g.Triangulation(model), 2 * order)
    G = g.BoundaryTriangulation(model, tags=[22, 25])                      # faces z = 1 and x = 0
    dG = g.Measure(G, 2 * order)
    gfun = lambda x: np.sin(x[:, 0]) + x[:, 1] * x[:, 2]  # noqa: E731
    a = lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dO + g.Integral(2.5 * (u * v)) * dG  # noqa: E731
    l = lambda v: g.Integral(v * 1.5) * dO + g.Integral(v * gfun) * dG  # noqa: E731
    op = g.AffineFEOperator(a, l, U, V)
    A, b = op.get_matrix(), op.get_vector()
    # oracle: bulk problem, then the facet problem added in place on the bulk pattern (nz_index path, src/Algebra/SparseMatrixCSC.jl:14-22)
    pb = problems.single_field_problem((0, 1) * 3, part, order=order, degree=2 * order, dirichlet_tags=dtags, form_mat=capi.LAPLACIAN,
                                       form_vec=capi.SOURCE, params=[1.5], X=X, dirichlet_values=U.dirichlet_values, lift=True)
    assert np.array_equal(pb.cell_dofs, V.cell_dof_ids)
    colptr, rowval, nzval, bo = pb.assemble(with_vector=True)
    fp0 = _facet_problem(G, V, 2 * order)
    xq = fp0.quadrature_points()
    fq = gfun(xq.reshape(-1, 3)).reshape(xq.shape[:2]) / 2.5
    fp = _facet_problem(G, V, 2 * order, form_mat=capi.MASS, fq=fq, dirichlet_values=U.dirichlet_values, lift=True)
    nzG, bG = np.zeros_like(nzval), np.zeros_like(bo)
    fp.assemble_inplace(colptr, rowval, nzG, bG, add=True)
    assert np.array_equal(A.colptr, colptr) and np.array_equal(A.rowval, rowval)
    assert relerr(A.nzval, nzval + 2.5 * nzG) <= 1e-12 and relerr(b, bo + 2.5 * bG) <= 1e-12
    assert np.abs(nzG).max() > 0 and np.abs(bG).max() > 0
    # the pieces alone: assemble_vector with the Neumann term only, assemble_matrix with bulk + Robin
    bN = g.assemble_vector(lambda v: g.Integral(v * gfun) * dG, V)
    fpn = _facet_problem(G, V, 2 * order, fq=fq * 2.5)
    assert relerr(bN, fpn.assemble_vector()) <= 1e-12
    A2 = g.assemble_matrix(a, U, V)
    assert relerr(A2.nzval, nzval + 2.5 * nzG) <= 1e-12
    with pytest.raises(NotImplementedError):
        g.assemble_matrix(lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dG, U, V)   # only mass terms on facets
    with pytest.raises(NotImplementedError):
        g.assemble_matrix(lambda u, v: g.Integral(u * v) * dG, U, V)                           # boundary-only bilinear form


def test_neumann_term_in_2d_on_segments():
    # QUAD model -> SEG2 facets (D = 2, Dr = 1), vector-valued Q1 field, constant traction on the right edge (tag 8)
    model = g.CartesianDiscreteModel((0, 2, 0, 1), (5, 4))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(2), 1), dirichlet_tags=[7, 1, 3])   # left edge clamped
    G = g.BoundaryTriangulation(model, tags=[8])
    assert G.num_cells() == 4
    dG = g.Measure(G, 2)
    b = g.assemble_vector(lambda v: g.Integral(g.inner(v, (3.0, -1.0))) * dG, V)
    # int_Gamma N_i = h/2 at the two end nodes of the edge, h at the inner ones (h = 1/4); components interleaved per node
    fx, fc, _, _ = V.dof_coordinates()
    on = np.isclose(fx[:, 0], 2.0)
    wgt = np.where(np.isclose(fx[:, 1], 0.0) | np.isclose(fx[:, 1], 1.0), 0.125, 0.25)
    expect = np.where(on, wgt * np.where(fc == 0, 3.0, -1.0), 0.0)
    assert np.abs(b - expect).max() <= 1e-14


def test_neumann_term_on_a_simplex_boundary():
    # P2 tetrahedra, traction on the face x = 1 (tag 26): facets are TRI3 cells embedded in 3D
    from test_host_logic import _facet_problem
    model = g.simplexify(g.CartesianDiscreteModel((0, 1) * 3, (3, 2, 2)))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 2), dirichlet_tags=[25])
    G = g.BoundaryTriangulation(model, tags=[26])
    dG = g.Measure(G, 4)
    gfun = lambda x: 1.0 + x[:, 1] * x[:, 2]  # noqa: E731
    b = g.assemble_vector(lambda v: g.Integral(v * gfun) * dG, V)
    fp0 = _facet_problem(G, V, 4)
    xq = fp0.quadrature_points()
    fp = _facet_problem(G, V, 4, fq=gfun(xq.reshape(-1, 3)).reshape(xq.shape[:2]))
    assert relerr(b, fp.assemble_vector()) <= 1e-12 and abs(b.sum()) > 0


def test_device_resident_hand_off_to_a_gpu_consumer():
    # N1 of SURVEY 8(f): the assembled system never leaves the GPU -- a device-side consumer (here: torch's sparse CSC SpMV and a
    # few CG iterations) binds the plan's colptr / rowval / nzval / b through the CUDA array interface, no download in between
    import torch
    n = 12
    model = g.UnstructuredDiscreteModel(g.CartesianDiscreteModel((0, 1) * 3, (n, n, n)))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags="boundary")
    U = g.TrialFESpace(V, lambda x: x[:, 0])
    dO = g.Measure(g.Triangulation(model), 2)
    assem = g.SparseMatrixAssembler(U, V)
    plan = assem.plan(dO)
    plan.set_state(0, None, U.dirichlet_values)
    plan.assemble_matrix_and_vector(lib.FORM_LAPLACIAN, (), lib.FORM_SOURCE, (1.0,), None, None, None)   # device-resident
    assem.ctx.synchronize()
    cp, rv, nz, bv = (torch.as_tensor(x, device="cuda") for x in plan.device_arrays())
    assert nz.data_ptr() == plan.device_nzval()[0]                      # zero-copy views
    A = torch.sparse_csc_tensor(cp, rv.to(torch.int64), nz, size=(plan.nrows, plan.ncols))
    x = torch.zeros(plan.nrows, dtype=torch.float64, device="cuda")
    r = bv.clone()
    p = r.clone()
    rs = torch.dot(r, r)
    for _ in range(200):                                                # conjugate gradients on the device
        Ap = torch.mv(A, p)
        alpha = rs / torch.dot(p, Ap)
        x += alpha * p
        r -= alpha * Ap
        rs_new = torch.dot(r, r)
        if float(rs_new) < 1e-24:
            break
        p = r + (rs_new / rs) * p
        rs = rs_new
    # reference: the same system downloaded and solved on the host
    import scipy.sparse.linalg as spla
    op = g.AffineFEOperator(lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dO, lambda v: g.Integral(v * 1.0) * dO, U, V)
    xh = spla.spsolve(op.get_matrix().to_scipy().tocsc(), op.get_vector())
    assert np.abs(x.cpu().numpy() - xh).max() <= 1e-9 * np.abs(xh).max()


def test_reference_conformance_checker():
    # test_sparse_matrix_assembler(a, matdata, vecdata, data) of the reference (src/FESpaces/SparseMatrixAssemblers.jl:110-114),
    # as its own tests call it (test/FESpacesTests/SparseMatrixAssemblersTests.jl:55-60)
    for model, T in ((g.CartesianDiscreteModel((0, 1, 0, 1), (6, 5)), float),
                     (g.UnstructuredDiscreteModel(g.CartesianDiscreteModel((0, 1) * 3, (5, 4, 4))), float),
                     (g.CartesianDiscreteModel((0, 1) * 3, (3, 3, 4)), g.VectorValue(3))):
        V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, T, 1), dirichlet_tags="boundary")
        ncomp = 1 if T is float else 3
        U = g.TrialFESpace(V, (lambda x: x[:, 0] + 1.0) if ncomp == 1 else (lambda x: np.stack([x[:, 0], x[:, 1], x[:, 0] * 0 + 2.0], axis=1)))
        dO = g.Measure(g.Triangulation(model), 2)
        a = lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dO  # noqa: E731
        l = (lambda v: g.Integral(v * 2.0) * dO) if ncomp == 1 else (lambda v: g.Integral(g.inner(v, (1.0, 2.0, 3.0))) * dO)  # noqa: E731
        u, v = g.get_trial_fe_basis(U), g.get_fe_basis(V)
        matdata = g.collect_cell_matrix(U, V, a(u, v))
        vecdata = g.collect_cell_vector(V, l(v))
        data = g.collect_cell_matrix_and_vector(U, V, a(u, v), l(v), g.FEFunction(U, np.zeros(U.num_free_dofs())))
        assert g.test_sparse_matrix_assembler(g.SparseMatrixAssembler(U, V), matdata, vecdata, data)
