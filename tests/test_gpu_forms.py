"""GPU parity of every supported integrand against the CPU oracle (configs 3-5 of BASELINE.json at reduced size).

Tolerance: max|d nzval| / max|nzval| <= 1e-12 (north star); patterns bit-exact.
Linear elasticity / Stokes matrices and the neo-Hookean law are not pinned by the reference's own tests
("parity unpinned"): the oracle pins them by construction (see test_oracle_forms.py)."""
import numpy as np
import pytest

import gridap_b200 as g
from gridap_b200 import lib
from oracle import capi, problems

pytestmark = pytest.mark.gpu

E, NU = 2.1e4, 0.3
LAM, MU = E * NU / ((1 + NU) * (1 - 2 * NU)), E / (2 * (1 + NU))


def relerr(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def check_csc(A, ref):
    colptr, rowval, nzval = ref[:3]
    assert np.array_equal(A.colptr, colptr) and np.array_equal(A.rowval, rowval)
    assert relerr(A.nzval, nzval) <= 1e-12


@pytest.mark.parametrize("order,n", [(1, 5), (2, 3)])
@pytest.mark.parametrize("deterministic", [False, True])
def test_config3_linear_elasticity(order, n, deterministic):
    # config 3: 3D linear elasticity, vector-valued hex, Dirichlet on the face x = 0 (entity 25 and its closure)
    part = (n, n, n)
    tags = [25, 1, 3, 5, 7, 13, 15, 17, 19]  # face x=0 + its 4 corners... closure listed explicitly like a Gridap tag
    model = g.CartesianDiscreteModel((0, 1) * 3, part)
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), order), dirichlet_tags=tags)
    U = g.TrialFESpace(V, (0.0, 0.0, 0.0))
    dO = g.Measure(g.Triangulation(model), 2 * order)
    sigma = g.IsotropicLinearElasticity.from_E_nu(E, NU)
    A = g.assemble_matrix(lambda u, v: g.Integral(g.inner(g.eps(v), sigma(g.eps(u)))) * dO,
                          g.SparseMatrixAssembler(U, V, deterministic=deterministic), U, V)
    pb = problems.single_field_problem((0, 1) * 3, part, order=order, ncomp=3, degree=2 * order, dirichlet_tags=tags,
                                       form_mat=capi.ELASTICITY, params=[LAM, MU])
    assert np.array_equal(pb.cell_dofs, V.cell_dof_ids)
    check_csc(A, pb.assemble())
    # the Cartesian mesh is affine: owner-computes column-node gather (deterministic by construction)
    assem = g.SparseMatrixAssembler(U, V, deterministic=deterministic)
    A1 = g.assemble_matrix(lambda u, v: g.Integral(g.inner(g.eps(v), sigma(g.eps(u)))) * dO, assem, U, V)
    assert assem.plan(dO).kernel_path(lib.FORM_ELASTICITY).startswith("affine_gather")
    assert np.array_equal(A1.nzval, A.nzval)
    # the cell-centric kernels on the same mesh (Q2: the local contraction on the FP64 tensor cores, DMMA)
    from parity_helpers import env
    with env(GB200_NO_AFFINE_GATHER=1, GB200_NO_STAGED_GATHER=1):
        assem = g.SparseMatrixAssembler(U, V, deterministic=deterministic)
        A2 = g.assemble_matrix(lambda u, v: g.Integral(g.inner(g.eps(v), sigma(g.eps(u)))) * dO, assem, U, V)
        path = assem.plan(dO).kernel_path(lib.FORM_ELASTICITY)
    assert path == ("vector_coloured" if deterministic else "vector_atomic") + ("+dmma" if order == 2 else "")
    check_csc(A2, pb.assemble())
    # staged gather (the route of non-affine meshes and state-dependent forms): cell-centric node-pair blocks through HBM, summed per
    # stored block -- no atomics, deterministic by construction (Q2 elasticity only on request: it keeps the tensor-core kernel)
    with env(GB200_NO_AFFINE_GATHER=1, GB200_STAGED_Q2=1):
        assem = g.SparseMatrixAssembler(U, V, deterministic=deterministic)
        A3 = g.assemble_matrix(lambda u, v: g.Integral(g.inner(g.eps(v), sigma(g.eps(u)))) * dO, assem, U, V)
        A4 = g.assemble_matrix(lambda u, v: g.Integral(g.inner(g.eps(v), sigma(g.eps(u)))) * dO, assem, U, V)
        assert assem.plan(dO).kernel_path(lib.FORM_ELASTICITY) == "staged_gather+blocks"
    check_csc(A3, pb.assemble())
    assert np.array_equal(A3.nzval, A4.nzval)


def test_vector_laplacian_and_mass_q1():
    part = (4, 3, 3)
    model = g.CartesianDiscreteModel((0, 1) * 3, part)
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 1), dirichlet_tags="boundary")
    dO = g.Measure(g.Triangulation(model), 2)
    A = g.assemble_matrix(lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u)) + 2.0 * g.dot(u, v)) * dO, V, V)
    pbl = problems.single_field_problem((0, 1) * 3, part, ncomp=3, form_mat=capi.LAPLACIAN)
    pbm = problems.single_field_problem((0, 1) * 3, part, ncomp=3, form_mat=capi.MASS)
    cl, rl, nl = pbl.assemble()
    cm, rm, nm = pbm.assemble()
    assert np.array_equal(A.colptr, cl) and np.array_equal(A.rowval, rl)
    assert relerr(A.nzval, nl + 2.0 * nm) <= 1e-12


@pytest.mark.parametrize("simplex", [True, False])
def test_config4_stokes_taylor_hood(simplex):
    # config 4: Taylor-Hood P2/P1 on the simplexified Cartesian mesh (Q2/Q1 on hexes as well), consecutive multi-field
    part = (3, 2, 2)
    model = g.CartesianDiscreteModel((0, 1) * 3, part)
    if simplex:
        model = g.simplexify(model)
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 2), dirichlet_tags="boundary")
    Q = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1))
    Y = g.MultiFieldFESpace([V, Q])
    X = g.MultiFieldFESpace([g.TrialFESpace(V, (0.0, 0.0, 0.0)), g.TrialFESpace(Q)])
    dO = g.Measure(g.Triangulation(model), 4)

    def a(up, vq):
        (u, p), (v, q) = up, vq
        return g.Integral(g.inner(g.grad(v), g.grad(u)) - g.div(v) * p + q * g.div(u)) * dO

    A = g.assemble_matrix(a, X, Y)
    pb = problems.stokes_problem((0, 1) * 3, part, degree=4, simplex=simplex)
    assert A.shape == (pb.nrows, pb.ncols)
    check_csc(A, pb.assemble())
    # the (q,p) block is absent from the pattern (src/Fields/FieldArrayBlocks.jl:488)
    nfu = V.num_free_dofs()
    for j in range(nfu, A.n):
        rows = A.rowval[A.colptr[j] - 1:A.colptr[j + 1] - 1]
        assert (rows <= nfu).all()


def test_config5_neohookean_residual_and_jacobian():
    # config 5: Q1 vector neo-Hookean, u(x) = 0.05 sin(pi x) sin(pi y) sin(pi z) (1,1,1), lambda=100, mu=1
    n = 4
    part = (n, n, n)
    model = g.CartesianDiscreteModel((0, 1) * 3, part)
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 1), dirichlet_tags="boundary")
    U = g.TrialFESpace(V, (0.0, 0.0, 0.0))
    dO = g.Measure(g.Triangulation(model), 2)
    nh = g.NeoHookean(100.0, 1.0)
    ufun = lambda x: 0.05 * (np.sin(np.pi * x[:, 0]) * np.sin(np.pi * x[:, 1]) * np.sin(np.pi * x[:, 2]))[:, None] * np.ones((1, 3))  # noqa: E731
    uh = g.interpolate(ufun, U)
    op = g.FEOperator(lambda u, v: g.Integral(nh.res(u, v)) * dO, lambda u, du, v: g.Integral(nh.jac(u, du, v)) * dO, U, V)
    b = op.residual(uh)
    A = op.jacobian(uh)
    pb = problems.single_field_problem((0, 1) * 3, part, ncomp=3, form_mat=capi.NEOHOOKEAN_JAC, form_vec=capi.NEOHOOKEAN_RES,
                                       params=[100.0, 1.0], free_values=uh.free_values, dirichlet_values=uh.dirichlet_values)
    colptr, rowval, nzval, bo = pb.assemble(with_vector=True)
    check_csc(A, (colptr, rowval, nzval))
    assert relerr(b, bo) <= 1e-12
    # re-assembly on the existing pattern (Newton loop): jacobian! twice gives the same matrix
    A2 = op.jacobian_(A, uh)
    assert relerr(A2.nzval, nzval) <= 1e-12
    # residual_and_jacobian!: one fused pass (src/FESpaces/FEOperatorsFromWeakForm.jl:85-103)
    b3, A3 = op.residual_and_jacobian(uh)
    check_csc(A3, (colptr, rowval, nzval))
    assert relerr(b3, bo) <= 1e-12
    assert op.assem.plan(dO).kernel_path(lib.FORM_NEOHOOKEAN_JAC) == "staged_gather+blocks"
    assert np.array_equal(op.jacobian(uh).nzval, A.nzval)   # no atomics on the matrix: bitwise reproducible
    # the cell-centric scatter (RED) of the same kernels
    from parity_helpers import env
    with env(GB200_NO_STAGED_GATHER=1):
        op2 = g.FEOperator(lambda u, v: g.Integral(nh.res(u, v)) * dO, lambda u, du, v: g.Integral(nh.jac(u, du, v)) * dO, U, V)
        b4, A4 = op2.residual_and_jacobian(uh)
        assert op2.assem.plan(dO).kernel_path(lib.FORM_NEOHOOKEAN_JAC) == "vector_atomic"
    check_csc(A4, (colptr, rowval, nzval))
    assert relerr(b4, bo) <= 1e-12


@pytest.mark.parametrize("case", [
    ("HEX", 2, 1, (3, 3, 2), 4, False),   # scalar Q2 hex
    ("TET", 1, 1, (3, 3, 3), 2, True),    # scalar P1 tets (affine simplices)
    ("TET", 2, 1, (2, 2, 3), 4, True),    # scalar P2 tets
    ("TRI", 1, 1, (5, 4), 2, True),       # scalar P1 triangles (Witherden-Vincent rule of degree 2)
    ("TRI", 2, 1, (4, 3), 4, True),       # scalar P2 triangles
    ("QUAD", 2, 1, (5, 4), 4, False),     # scalar Q2 quads
    ("QUAD", 1, 2, (6, 5), 2, False),     # vector Q1 quads (2 components: generic kernel)
    ("HEX", 1, 1, (4, 4, 3), 4, False),   # scalar Q1 hex with the 27-point rule: still exact -> affine gather path
])
@pytest.mark.parametrize("form", ["laplacian", "mass"])
def test_scalar_and_2d_elements_on_the_generic_path(case, form):
    ptype, order, ncomp, part, degree, simplex = case
    D = len(part)
    model = g.CartesianDiscreteModel((0, 1) * D, part)
    if simplex:
        model = g.simplexify(model)
    T = float if ncomp == 1 else g.VectorValue(ncomp)
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, T, order), dirichlet_tags="boundary")
    dO = g.Measure(g.Triangulation(model), degree)
    if form == "laplacian":
        a = lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dO  # noqa: E731
        fid = capi.LAPLACIAN
    else:
        a = lambda u, v: g.Integral(g.inner(u, v)) * dO  # noqa: E731
        fid = capi.MASS
    A = g.assemble_matrix(a, V, V)
    pb = problems.single_field_problem((0, 1) * D, part, order=order, ncomp=ncomp, degree=degree, form_mat=fid, simplex=simplex)
    assert np.array_equal(pb.cell_dofs, V.cell_dof_ids)
    check_csc(A, pb.assemble())


def test_config4_stokes_block_multifield_style():
    # BlockMultiFieldStyle: BlockMatrix of CSCs == the consecutive matrix (test/MultiFieldTests/BlockSparseMatrixAssemblersTests.jl:53-56)
    part = (3, 2, 2)
    model = g.simplexify(g.CartesianDiscreteModel((0, 1) * 3, part))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 2), dirichlet_tags="boundary")
    Q = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1))
    dO = g.Measure(g.Triangulation(model), 4)

    def a(up, vq):
        (u, p), (v, q) = up, vq
        return g.Integral(g.inner(g.grad(v), g.grad(u)) - g.div(v) * p + q * g.div(u)) * dO

    Y = g.MultiFieldFESpace([V, Q])
    A1 = g.assemble_matrix(a, Y, Y)
    Yb = g.MultiFieldFESpace([V, Q], style=g.BlockMultiFieldStyle())
    assem = g.SparseMatrixAssembler(Yb, Yb)
    Ab = g.assemble_matrix(a, assem, Yb, Yb)
    assert isinstance(Ab, g.BlockMatrix) and Ab.blocksize() == (2, 2) and Ab.shape == A1.shape and Ab.nnz() == A1.nnz()
    nu, npr = V.num_free_dofs(), Q.num_free_dofs()
    assert Ab.blocks[0][1].shape == (nu, npr) and Ab.blocks[1][1].nnz() == 0   # (q,p) block untouched: empty
    assert (Ab.blocks[1][1].colptr == 1).all()
    S1 = A1.to_scipy()
    for i, rs in enumerate([slice(0, nu), slice(nu, nu + npr)]):
        for j, cs in enumerate([slice(0, nu), slice(nu, nu + npr)]):
            ref = S1[rs, cs].tocsc()
            ref.sort_indices()
            blk = Ab.blocks[i][j]
            assert np.array_equal(blk.colptr - 1, ref.indptr) and np.array_equal(blk.rowval - 1, ref.indices)  # canonical CSC per block
            assert len(ref.data) == 0 or relerr(blk.nzval, ref.data) <= 1e-13   # two atomic assemblies: same values up to summation order
    # in-place re-assembly and vectors
    A3 = assem.allocate_matrix(g.collect_cell_matrix(Yb, Yb, a(g.get_trial_fe_basis(Yb), g.get_fe_basis(Yb))))
    assem.assemble_matrix_(A3, g.collect_cell_matrix(Yb, Yb, a(g.get_trial_fe_basis(Yb), g.get_fe_basis(Yb))))
    assert abs(A3.to_scipy() - S1).max() <= 1e-13 * abs(S1).max()
    with pytest.raises(NotImplementedError):
        g.SparseMatrixAssembler(Y, Yb)
