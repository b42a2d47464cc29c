"""Round-2 coverage on the device: multi-field vectors / AffineFEOperator on Stokes (consecutive and block style), loads that live
only on a BoundaryTriangulation, `_add!` on CSR / SymCSR / Block outputs, GenericAssemblyStrategy (row / column maps and masks),
the general multi-GPU partition (multi-field, state-carrying forms) run rank after rank on one GPU."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import gridap_b200 as g
from gridap_b200 import distributed as gd
from oracle import capi
from parity_helpers import check_csc, facet_problem, oracle_field, oracle_problem, perturb, relerr

pytestmark = pytest.mark.gpu


def _laplacian(dO):
    return lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dO


def test_affine_operator_with_a_boundary_only_load():
    # l(v) = int_Gamma g v only (same quadrature degree as the bilinear form): the load must NOT be integrated over the bulk
    model = g.CartesianDiscreteModel((0, 1) * 3, (5, 4, 3))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags=[25, 1, 3, 5, 7, 13, 15, 17, 19])   # face x = 0
    gD = lambda x: 1.0 + x[:, 1] - x[:, 2]   # noqa: E731
    U = g.TrialFESpace(V, gD)
    dO = g.Measure(g.Triangulation(model), 2)
    Gam = g.BoundaryTriangulation(model, tags=[26])     # face x = 1
    dG = g.Measure(Gam, 2)
    op = g.AffineFEOperator(_laplacian(dO), lambda v: g.Integral(v * 3.0) * dG, U, V)
    pb = oracle_problem(model, [oracle_field(model, V, 2, dirichlet_values=U.dirichlet_values)], 2, capi.LAPLACIAN, capi.SOURCE, params=[0.0], lift=True,
                        nrows=V.nfree, ncols=V.nfree)
    colptr, rowval, nzval, b = pb.assemble(with_vector=True)
    b = b + facet_problem(Gam, V, 2, params=[3.0]).assemble_vector()
    check_csc(op.get_matrix(), (colptr, rowval, nzval))
    assert relerr(op.get_vector(), b) <= 1e-12
    assert abs(op.get_vector().sum() - b.sum()) <= 1e-12 * abs(b).max()
    # FEOperator-style pairing check as well: matrix + vector on different triangulations through assemble_matrix_and_vector
    A2, b2 = g.assemble_matrix_and_vector(_laplacian(dO), lambda v: g.Integral(v * 3.0) * dG, U, V)
    assert relerr(b2, b) <= 1e-12 and relerr(A2.nzval, nzval) <= 1e-12


def _stokes_2d_reference_problem(style=None):
    """test/GridapTests/StokesTaylorHoodTests.jl:6-63 through the public API of the package"""
    model = g.CartesianDiscreteModel((0, 2, 0, 2), (3, 3))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(2), 2), dirichlet_tags=[1, 2, 5])
    Q = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1))
    u = lambda x: np.stack([x[:, 0] ** 2 + 2 * x[:, 1] ** 2, -x[:, 0] ** 2], axis=1)   # noqa: E731
    p = lambda x: x[:, 0] + 3 * x[:, 1]                                                   # noqa: E731
    Y = g.MultiFieldFESpace([V, Q], style=style)
    X = g.MultiFieldFESpace([g.TrialFESpace(V, u), g.TrialFESpace(Q)], style=style)
    dO = g.Measure(g.Triangulation(model), 2)
    Gam = g.BoundaryTriangulation(model, tags=[6, 7, 8])
    dG = g.Measure(Gam, 2)

    def t(x):   # n.grad(u) - p n on the three Neumann sides of (0,2)^2
        n = np.zeros_like(x)
        n[np.isclose(x[:, 0], 0.0)] = (-1.0, 0.0)
        n[np.isclose(x[:, 0], 2.0)] = (1.0, 0.0)
        n[np.isclose(x[:, 1], 2.0)] = (0.0, 1.0)    # (corners are not quadrature points)
        gu = np.stack([np.stack([2 * x[:, 0], -2 * x[:, 0]], axis=-1), np.stack([4 * x[:, 1], 0 * x[:, 1]], axis=-1)], axis=-2)
        return np.einsum("fi,fij->fj", n, gu) - p(x)[:, None] * n

    def a(up, vq):
        (uu, pp), (v, q) = up, vq
        return g.Integral(g.inner(g.grad(v), g.grad(uu)) - g.div(v) * pp + q * g.div(uu)) * dO

    def l(vq):
        v, q = vq
        return g.Integral(g.dot(v, (-5.0, 5.0)) + q * (lambda x: 2.0 * x[:, 0])) * dO + g.Integral(g.dot(v, t)) * dG

    return model, V, Q, X, Y, a, l, u, p


@pytest.mark.parametrize("block", [False, True])
def test_stokes_affine_operator_manufactured_solution(block):
    # AffineFEOperator(a,l,X,Y) on the reference's own Stokes driver; the reference asserts errors < 1e-9 (the manufactured
    # solution lies in the discrete spaces, so u_h / p_h are its nodal values)
    model, V, Q, X, Y, a, l, u, p = _stokes_2d_reference_problem(g.BlockMultiFieldStyle() if block else None)
    op = g.AffineFEOperator(a, l, X, Y)
    A, b = op.get_matrix(), op.get_vector()
    if block:
        assert isinstance(A, g.BlockMatrix) and isinstance(b, g.BlockVector)
    S = A.to_scipy().tocsc()
    x = spla.spsolve(S, np.asarray(b))
    nfu = V.nfree
    fx, fc, _, _ = V.dof_coordinates()
    assert np.abs(x[:nfu] - u(fx)[np.arange(nfu), fc]).max() < 1e-9
    assert np.abs(x[nfu:] - p(Q.dof_coordinates()[0])).max() < 1e-9


def test_stokes_3d_matrix_and_vector_against_the_oracle():
    # config 4 style (P2/P1 tets, perturbed): sources on both fields, inhomogeneous velocity data -> lifting through all blocks
    model = g.simplexify(perturb(g.CartesianDiscreteModel((0, 1) * 3, (3, 4, 3)), 0.1, 5))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 2), dirichlet_tags="boundary")
    Q = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1))
    ud = lambda x: np.stack([x[:, 1] * x[:, 2], np.sin(x[:, 0]), x[:, 0] + x[:, 1]], axis=1)   # noqa: E731
    Y = g.MultiFieldFESpace([V, Q])
    Ut = g.TrialFESpace(V, ud)
    X = g.MultiFieldFESpace([Ut, g.TrialFESpace(Q)])
    dO = g.Measure(g.Triangulation(model), 4)
    gq = lambda x: 1.0 + x[:, 0] * x[:, 2]   # noqa: E731

    def a(up, vq):
        (uu, pp), (v, q) = up, vq
        return g.Integral(g.inner(g.grad(v), g.grad(uu)) - g.div(v) * pp + q * g.div(uu)) * dO

    def l(vq):
        v, q = vq
        return g.Integral(g.dot(v, (1.0, -2.0, 0.5)) + q * gq) * dO

    op = g.AffineFEOperator(a, l, X, Y)
    ids = Y.get_cell_dof_ids()
    n = V.nfree + Q.nfree
    geo = oracle_problem(model, [oracle_field(model, Q, 4)], 4, capi.MASS, nrows=Q.nfree, ncols=Q.nfree)
    xq = geo.quadrature_points()
    fu = oracle_field(model, V, 4, 0, ids=ids[0], dirichlet_values=Ut.dirichlet_values, src=[1.0, -2.0, 0.5])
    fp = oracle_field(model, Q, 4, V.nfree, ids=ids[1], fq=gq(xq.reshape(-1, 3)).reshape(xq.shape[0], xq.shape[1], 1))
    pb = oracle_problem(model, [fu, fp], 4, capi.STOKES, capi.SOURCE, touched=np.array([[1, 1], [1, 0]], dtype=np.uint8), lift=True, nrows=n, ncols=n)
    colptr, rowval, nzval, b = pb.assemble(with_vector=True)
    check_csc(op.get_matrix(), (colptr, rowval, nzval))
    assert relerr(op.get_vector(), b) <= 1e-12
    # assemble_vector alone on the multi-field space
    bv = g.assemble_vector(l, Y)
    pbv = oracle_problem(model, [oracle_field(model, V, 4, 0, ids=ids[0], src=[1.0, -2.0, 0.5]),
                                 oracle_field(model, Q, 4, V.nfree, ids=ids[1], fq=gq(xq.reshape(-1, 3)).reshape(xq.shape[0], xq.shape[1], 1))],
                         4, 0, capi.SOURCE, touched=np.array([[1, 1], [1, 0]], dtype=np.uint8), nrows=n, ncols=n)
    assert relerr(bv, pbv.assemble_vector()) <= 1e-12


def test_add_on_csr_symcsr_and_block_outputs():
    # assemble_matrix_add! on every matrix type (the stage loops of src/ODEs/ODEOpsFromTFEOps.jl:124-405 call it)
    model = g.CartesianDiscreteModel((0, 1) * 3, (4, 3, 3))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags=[25])
    dO = g.Measure(g.Triangulation(model), 2)
    a = lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u)) + 2.0 * (u * v)) * dO   # noqa: E731  (two terms: device-resident sum)
    A = g.assemble_matrix(a, V, V).to_scipy()
    matdata = g.collect_cell_matrix(V, V, a(g.get_trial_fe_basis(V), g.get_fe_basis(V)))
    for T in (g.SparseMatrixCSR[0], g.SparseMatrixCSR[1], g.SymSparseMatrixCSR[1], g.SymSparseMatrixCSR[0]):
        assem = g.SparseMatrixAssembler(T, np.ndarray, V, V)
        B = assem.assemble_matrix(matdata)
        assert abs(B.to_scipy() - A).max() <= 1e-13 * abs(A).max()
        assem.assemble_matrix_add_(B, matdata)
        assert abs(B.to_scipy() - 2.0 * A).max() <= 1e-13 * abs(A).max()
        if issubclass(T, g.SymSparseMatrixCSR):   # upper triangle only, columns ascending per row
            U = sp.triu(A).tocsr()
            U.sort_indices()
            assert np.array_equal(B.rowptr - T.Bi, U.indptr) and np.array_equal(B.colval - T.Bi, U.indices)
        data = g.collect_cell_matrix_and_vector(V, V, a(g.get_trial_fe_basis(V), g.get_fe_basis(V)), (g.Integral(g.get_fe_basis(V) * 1.0) * dO), g.zero(V))
        B2, b2 = assem.assemble_matrix_and_vector(data)
        assem.assemble_matrix_and_vector_add_(B2, b2, data)
        assert abs(B2.to_scipy() - 2.0 * A).max() <= 1e-13 * abs(A).max()
    # block style
    mt = g.simplexify(g.CartesianDiscreteModel((0, 1) * 3, (2, 2, 2)))
    Vv = g.TestFESpace(mt, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 2), dirichlet_tags="boundary")
    Q = g.TestFESpace(mt, g.ReferenceFE(g.lagrangian, float, 1))
    Yb = g.MultiFieldFESpace([Vv, Q], style=g.BlockMultiFieldStyle())
    dT = g.Measure(g.Triangulation(mt), 4)

    def ast(up, vq):
        (u, p), (v, q) = up, vq
        return g.Integral(g.inner(g.grad(v), g.grad(u)) - g.div(v) * p + q * g.div(u)) * dT
    assem = g.SparseMatrixAssembler(Yb, Yb)
    md = g.collect_cell_matrix(Yb, Yb, ast(g.get_trial_fe_basis(Yb), g.get_fe_basis(Yb)))
    Ab = assem.assemble_matrix(md)
    S1 = Ab.to_scipy().copy()
    assem.assemble_matrix_add_(Ab, md)
    assert abs(Ab.to_scipy() - 2.0 * S1).max() <= 1e-13 * abs(S1).max()


def test_generic_assembly_strategy_row_and_column_maps():
    # GenericAssemblyStrategy(row_map, col_map, row_mask, col_mask) (src/FESpaces/Assemblers.jl:31-55,134-150): rows permuted,
    # every third row masked, columns reversed and the first ten masked; compare with the same operation on the full matrix
    model = g.CartesianDiscreteModel((0, 1) * 3, (4, 4, 3))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags=[25])
    U = g.TrialFESpace(V, lambda x: 1.0 + x[:, 1])
    n = V.nfree
    dO = g.Measure(g.Triangulation(model), 2)
    a, l = _laplacian(dO), (lambda v: g.Integral(v * 2.0) * dO)
    A = g.assemble_matrix(a, U, V).to_scipy().tocsr()
    op0 = g.AffineFEOperator(a, l, U, V)
    perm = np.random.default_rng(3).permutation(n) + 1
    strat = g.GenericAssemblyStrategy(lambda r: perm[r - 1], lambda c: n + 1 - c, lambda r: r % 3 != 0, lambda c: c > 10)
    assem = g.SparseMatrixAssembler(g.SparseMatrixCSC, np.ndarray, U, V, strat)
    assert assem.get_assembly_strategy() is strat
    B = g.assemble_matrix(a, assem, U, V).to_scipy().toarray()
    rows = np.arange(1, n + 1)
    keep_r, keep_c = rows % 3 != 0, rows > 10
    ref = np.zeros((n, n))
    Ad = A.toarray()
    ref[np.ix_(perm[keep_r] - 1, (n + 1 - rows[keep_c]) - 1)] = Ad[np.ix_(keep_r, keep_c)]
    assert np.abs(B - ref).max() <= 1e-13 * np.abs(ref).max()
    # the vector (with lifting) follows the row map / mask; the lifting itself is unaffected by the column mask
    op = g.AffineFEOperator(a, l, U, V, assem)
    bref = np.zeros(n)
    bref[perm[keep_r] - 1] = op0.get_vector()[keep_r]
    assert np.abs(op.get_vector() - bref).max() <= 1e-12 * np.abs(bref).max()
    # the reference's AssemblyStrategyMock (identity maps, test/MultiFieldTests/MultiFieldSparseMatrixAssemblersTests.jl:96-100)
    mock = g.GenericAssemblyStrategy(lambda r: r, lambda c: c, lambda r: np.ones(len(r), bool), lambda c: np.ones(len(c), bool))
    Bm = g.assemble_matrix(a, g.SparseMatrixAssembler(g.SparseMatrixCSC, np.ndarray, U, V, mock), U, V)
    assert abs(Bm.to_scipy() - A).max() <= 1e-14 * abs(A).max()


def _run_partitioned(model, U, V, world, build_forms, uh=None, deterministic=False):
    """every rank's assembly run one after the other on this GPU; returns the gathered global matrix and vector"""
    slabs, vecs, owned = [], [], []
    for rank in range(world):
        part = gd.partition(model, U, V, world, rank)
        asm = part.assembler(deterministic=deterministic)
        A, b = build_forms(part.local_model, part.local_space, asm, None if uh is None else part.local_function(uh))
        assert A.shape == (V.num_free_dofs(), len(part.owned_ids))
        slabs.append((A.colptr, A.rowval, A.nzval))
        vecs.append(b)
        owned.append(part.owned_ids)
    n = V.num_free_dofs()
    return gd.gather_csc_owned(slabs, owned, n), gd.gather_vector_owned(vecs, owned, n)


def test_partition_stokes_multifield_on_the_device():
    # multi-GPU path for multi-field plans: per-field column sets from one cell partition; ranks run one after the other here
    model = g.simplexify(g.CartesianDiscreteModel((0, 1) * 3, (3, 3, 5)))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 2), dirichlet_tags="boundary")
    Q = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1))
    ud = lambda x: np.stack([x[:, 1], x[:, 2] ** 2, x[:, 0]], axis=1)   # noqa: E731
    Y = g.MultiFieldFESpace([V, Q])
    X = g.MultiFieldFESpace([g.TrialFESpace(V, ud), g.TrialFESpace(Q)])

    def forms(m, Xl, asm, _):
        dO = g.Measure(g.Triangulation(m), 4)

        def a(up, vq):
            (u, p), (v, q) = up, vq
            return g.Integral(g.inner(g.grad(v), g.grad(u)) - g.div(v) * p + q * g.div(u)) * dO

        def l(vq):
            v, q = vq
            return g.Integral(g.dot(v, (1.0, 0.0, -1.0)) + q * 0.5) * dO
        if asm is None:
            op = g.AffineFEOperator(a, l, X, Y)
        else:
            op = g.AffineFEOperator(a, l, Xl, Xl, asm)
        return op.get_matrix(), op.get_vector()

    A, b = forms(model, None, None, None)
    G, bg = _run_partitioned(model, X, Y, 3, forms)
    assert np.array_equal(G.colptr, A.colptr) and np.array_equal(G.rowval, A.rowval)
    assert relerr(G.nzval, A.nzval) <= 1e-13 and relerr(bg, b) <= 1e-12


@pytest.mark.parametrize("deterministic", [False, True])
def test_partition_neohookean_state_on_the_device(deterministic):
    # forms that carry u_h on a column-partitioned assembler: u_h is gathered through the unmasked global ids (state space)
    model = g.UnstructuredDiscreteModel(g.CartesianDiscreteModel((0, 1) * 3, (5, 4, 7)))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 1), dirichlet_tags="boundary")
    U = g.TrialFESpace(V, (0.0, 0.0, 0.0))
    nh = g.NeoHookean(100.0, 1.0)
    uh = g.interpolate(lambda x: 0.05 * (np.sin(np.pi * x[:, 0]) * np.sin(np.pi * x[:, 1]) * np.sin(np.pi * x[:, 2]))[:, None] * np.array([[1.0, 0.5, -1.0]]), U)

    def forms(m, Ul, asm, uhl):
        dO = g.Measure(g.Triangulation(m), 2)
        res = lambda u, v: g.Integral(nh.res(u, v)) * dO          # noqa: E731
        jac = lambda u, du, v: g.Integral(nh.jac(u, du, v)) * dO  # noqa: E731
        if asm is None:
            op = g.FEOperator(res, jac, U, V, g.SparseMatrixAssembler(U, V, deterministic=deterministic))
            return op.jacobian(uh), op.residual(uh)
        op = g.FEOperator(res, jac, Ul, Ul, asm)
        b, A = op.residual_and_jacobian(uhl)
        assert relerr(op.residual(uhl)[asm.strategy.owned_ids - 1], b[asm.strategy.owned_ids - 1]) <= 1e-12
        return A, b

    A, b = forms(model, None, None, None)
    G, bg = _run_partitioned(model, U, V, 3, forms, uh, deterministic)
    assert np.array_equal(G.colptr, A.colptr) and np.array_equal(G.rowval, A.rowval)
    assert relerr(G.nzval, A.nzval) <= 1e-13 and relerr(bg, b) <= 1e-12


def test_partition_q1_poisson_general_partition_is_bitwise():
    # the general partition on the headline element: owner-computes gather per rank, bitwise equal to the single-GPU matrix
    model = g.UnstructuredDiscreteModel(g.CartesianDiscreteModel((0, 1) * 3, (9, 8, 11)))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags="boundary")
    U = g.TrialFESpace(V, lambda x: x[:, 0] - x[:, 2])

    def forms(m, Ul, asm, _):
        dO = g.Measure(g.Triangulation(m), 2)
        a, l = _laplacian(dO), (lambda v: g.Integral(v * 1.0) * dO)
        op = g.AffineFEOperator(a, l, U, V) if asm is None else g.AffineFEOperator(a, l, Ul, Ul, asm)
        return op.get_matrix(), op.get_vector()

    A, b = forms(model, None, None, None)
    G, bg = _run_partitioned(model, U, V, 4, forms)
    assert np.array_equal(G.colptr, A.colptr) and np.array_equal(G.rowval, A.rowval) and np.array_equal(G.nzval, A.nzval)
    assert relerr(bg, b) <= 1e-13
