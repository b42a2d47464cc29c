"""N4 of SURVEY 8(f), remainder: spaces with linear constraints (FESpaceWithLinearConstraints,
src/FESpaces/FESpacesWithLinearConstraints.jl; attach_constraints_rows / _cols, src/FESpaces/FESpaceInterface.jl:361-387).
The reference's own test (test/FESpacesTests/FESpacesWithLinearConstraintsTests.jl) replayed: numbering goldens, the cell DoF
values of a constrained FE function, and the constrained Poisson problem with a skeleton term solved to 1e-9; plus entry-wise parity
of the device result with the cell-wise restatement C_e K_e C_e^T of the oracle on a larger mesh with hanging-node-like constraints."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import gridap_b200 as g
from oracle import capi
from parity_helpers import oracle_field, oracle_problem

pytestmark = pytest.mark.gpu


def _reference_spaces():
    model = g.CartesianDiscreteModel((0, 1, 0, 1), (2, 2))
    V = g.FESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags=[1, 2, 5])   # "dirichlet" = tags [1,2,5]
    Vc = g.FESpaceWithLinearConstraints([1, 5, -2], [[-1, 4], [4, 6], [-1, -3]], [[0.5, 0.5]] * 3, V)
    return model, V, Vc


def test_reference_constrained_poisson_with_skeleton_term():
    # FESpacesWithLinearConstraintsTests.jl:30-95
    model, V, Vc = _reference_spaces()
    assert g.has_constraints(Vc) and Vc.n_fdofs == 6 and Vc.n_fmdofs == 4
    fv, dv = Vc.scatter_free_and_dirichlet_values(np.arange(1.0, Vc.num_free_dofs() + 1), -np.arange(1.0, Vc.num_dirichlet_dofs() + 1))
    ids = V.cell_dof_ids
    vals = np.concatenate([fv, dv])[np.where(ids > 0, ids - 1, V.nfree - ids - 1)]
    assert np.allclose(vals, [[-1.0, -1.5, 1.0, 1.0], [-1.5, -2.0, 1.0, 2.0], [1.0, 1.0, 3.0, 3.5], [1.0, 2.0, 3.5, 4.0]])   # :58-59
    u = lambda x: x[:, 0] + 2 * x[:, 1]   # noqa: E731
    Uc = g.TrialFESpace(Vc, u)
    assert g.has_constraints(Uc)
    dO = g.Measure(g.Triangulation(model), 2)
    Gam = g.BoundaryTriangulation(model, tags=[6, 7, 8])   # "neumann"
    Lam = g.SkeletonTriangulation(model)
    dG, dL = g.Measure(Gam, 2), g.Measure(Lam, 2)
    flux = lambda x: np.where(np.abs(x[:, 1] - 1) < 1e-12, 2.0, np.where(np.abs(x[:, 0]) < 1e-12, -1.0, 1.0))   # noqa: E731  n . grad u
    a = lambda uu, v: g.Integral(g.inner(g.grad(v), g.grad(uu))) * dO + g.Integral(g.jump(uu) * g.jump(v)) * dL   # noqa: E731
    l = lambda v: g.Integral(v * 0.0) * dO + g.Integral(v * flux) * dG   # noqa: E731  (f = -Laplace u = 0)
    op = g.AffineFEOperator(a, l, Uc, Vc)
    A, b = op.get_matrix().to_scipy().tocsc(), op.get_vector()
    assert A.shape == (4, 4)
    x = spla.spsolve(A, b)
    uex = g.interpolate(u, Uc).free_values
    assert np.abs(x - uex).max() <= 1e-9                       # the reference's tolerance (:88-91)


def _cellwise_reference(model, V, Vc, form, params, degree):
    """sum_e C_e K_e C_e^T on the masters, dense: K_e from the oracle (one cell at a time through the extended numbering)"""
    n = V.nfree + V.ndirichlet
    ext = Vc.extended
    pb = oracle_problem(model, [oracle_field(model, ext, degree)], degree, form, params=params, nrows=n, ncols=n)
    colptr, rowval, nzval = pb.assemble()
    A = sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(n, n)).toarray()
    T = np.zeros((n, Vc.num_free_dofs() + Vc.num_dirichlet_dofs()))
    p0 = Vc.DOF_to_mDOFs_ptrs - 1
    for D in range(n):
        for q in range(p0[D], p0[D + 1]):
            m = Vc.DOF_to_mdofs[q]
            T[D, m - 1 if m > 0 else Vc.n_fmdofs - m - 1] += Vc.DOF_to_coeffs[q]
    return (1.0 if form != capi.MASS else params[0]) * (T.T @ A @ T)   # (the oracle's mass integrand carries no coefficient)


@pytest.mark.parametrize("ptype", ["QUAD", "HEX"])
def test_constrained_assembly_against_the_cellwise_restatement(ptype):
    D = 2 if ptype == "QUAD" else 3
    part = (6, 5) if D == 2 else (4, 3, 3)
    model = g.CartesianDiscreteModel((0, 1) * D, part)
    V = g.FESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags=[1, 2, 5] if D == 2 else [21, 22])
    rng = np.random.default_rng(11)
    free = rng.permutation(V.nfree)
    slaves, masters = free[:V.nfree // 5] + 1, free[V.nfree // 5:] + 1
    dofs = [sorted(rng.choice(masters, size=2 + (k % 2), replace=False).tolist()) + ([-1] if k % 3 == 0 else []) for k in range(len(slaves))]
    coeffs = [(np.ones(len(r)) / len(r)).tolist() for r in dofs]
    Vc = g.FESpaceWithLinearConstraints(slaves.tolist(), dofs, coeffs, V)
    Uc = g.TrialFESpace(Vc, lambda x: 1.0 + x[:, 0])
    dO = g.Measure(g.Triangulation(model), 2)
    f = lambda x: np.sin(3 * x[:, 0]) + x[:, 1]   # noqa: E731
    op = g.AffineFEOperator(lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u)) + 2.5 * (v * u)) * dO, lambda v: g.Integral(v * f) * dO, Uc, Vc)
    Ac = op.get_matrix().to_scipy().toarray()
    full = _cellwise_reference(model, V, Vc, capi.LAPLACIAN, [1.0], 2) + _cellwise_reference(model, V, Vc, capi.MASS, [2.5], 2)
    nfm = Vc.n_fmdofs
    assert np.abs(Ac - full[:nfm, :nfm]).max() <= 1e-12 * np.abs(full).max()
    # vector: T^T b - A_c[:, Dirichlet masters] u_D
    n = V.nfree + V.ndirichlet
    ext = Vc.extended
    pbv = oracle_problem(model, [oracle_field(model, ext, 2)], 2, 0, capi.SOURCE, nrows=n, ncols=n,
                         fq=f(oracle_problem(model, [oracle_field(model, ext, 2)], 2, capi.MASS, nrows=n, ncols=n).quadrature_points().reshape(-1, D)).reshape(model.num_cells(), -1, 1))
    bext = pbv.assemble_vector()
    T = np.zeros((n, nfm + Vc.num_dirichlet_dofs()))
    p0 = Vc.DOF_to_mDOFs_ptrs - 1
    for Dd in range(n):
        for q in range(p0[Dd], p0[Dd + 1]):
            m = Vc.DOF_to_mdofs[q]
            T[Dd, m - 1 if m > 0 else nfm - m - 1] += Vc.DOF_to_coeffs[q]
    bref = (T.T @ bext)[:nfm] - full[:nfm, nfm:] @ Uc.dirichlet_values
    assert np.abs(op.get_vector() - bref).max() <= 1e-12 * max(np.abs(bref).max(), 1.0)
    # the pattern: all pairs of masters of a cell (get_cell_dof_ids of the constrained space drives the symbolic loop)
    A = op.get_matrix()
    stored = np.zeros((nfm, nfm), dtype=bool)
    stored[A.rowval - 1, np.repeat(np.arange(nfm), np.diff(A.colptr))] = True
    want = np.zeros((nfm, nfm), dtype=bool)
    for row in Vc.get_cell_dof_ids():
        m = row[row > 0] - 1
        want[np.ix_(m, m)] = True
    assert np.array_equal(stored, want)


def test_zero_mean_space_and_stokes_with_zero_mean_pressure():
    # test/FESpacesTests/ZeroMeanFESpacesTests.jl: constraint = :zeromean = FESpaceWithConstantFixed(space, true, num_free_dofs(space))
    # + the mean shift of FE functions (src/FESpaces/ZeroMeanFESpaces.jl:11-75); (0,1)^2, partition (4,4), order 2
    model = g.CartesianDiscreteModel((0, 1, 0, 1), (4, 4))
    order = 2
    dO = g.Measure(g.Triangulation(model), order)
    V0 = g.FESpace(model, g.ReferenceFE(g.lagrangian, float, order), conformity="L2", constraint="zeromean")
    assert V0.num_dirichlet_dofs() == 1 and V0.num_free_dofs() == 16 * 9 - 1 and V0.cell_dof_ids[-1, -1] == -1
    U0 = g.TrialFESpace(V0)
    mean_of = lambda vh: float(np.dot(np.concatenate([vh.free_values, vh.dirichlet_values]), V0._vol_i))   # noqa: E731  int(vh) dOmega
    f = lambda x: np.sin(4 * np.pi * (x[:, 0] + x[:, 1] ** 2)) + 3   # noqa: E731  non-zero mean (:36-38)
    uh = g.interpolate(f, U0)
    assert abs(mean_of(uh)) < 1e-10
    gm = 1.0 / 3.0 + 0.5                                             # mean of x^2 + y over the unit square
    gz = lambda x: x[:, 0] ** 2 + x[:, 1] - gm                       # noqa: E731  zero mean, in the space (:41-47)
    vh = g.interpolate(gz, U0)
    fx, _, dx, _ = V0.dof_coordinates()
    assert abs(mean_of(vh)) < 1e-10
    assert np.abs(vh.free_values - gz(fx)).max() < 1e-10 and np.abs(vh.dirichlet_values - gz(dx)).max() < 1e-10
    # Stokes with a zero-mean pressure (:51-83), Taylor-Hood Q2/Q1 here: u_ex = (y, -x), p_ex = x + 2y; l = a((u_ex, p_ex), .) reduces
    # to int v.grad(p_ex) for test functions vanishing on the boundary (u_ex is harmonic and divergence-free)
    V = g.FESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(2), 2), dirichlet_tags="boundary")
    U = g.TrialFESpace(V, lambda x: np.stack([x[:, 1], -x[:, 0]], axis=1))
    Q0 = g.FESpace(model, g.ReferenceFE(g.lagrangian, float, 1), constraint="zeromean")
    X, Y = g.MultiFieldFESpace([U, g.TrialFESpace(Q0)]), g.MultiFieldFESpace([V, Q0])
    dO4 = g.Measure(g.Triangulation(model), 4)

    def a(up, vq):
        (u, p), (v, q) = up, vq
        return g.Integral(g.inner(g.grad(v), g.grad(u)) - g.div(v) * p + q * g.div(u)) * dO4

    def l(vq):
        v, q = vq
        return g.Integral(g.dot(v, (1.0, 2.0)) + q * 0.0) * dO4

    op = g.AffineFEOperator(a, l, X, Y)
    sol = spla.spsolve(op.get_matrix().to_scipy().tocsc(), op.get_vector())
    nu = V.num_free_dofs()
    uex = g.interpolate(lambda x: np.stack([x[:, 1], -x[:, 0]], axis=1), U).free_values
    assert np.abs(sol[:nu] - uex).max() < 1e-10
    ph = g.FEFunction(g.TrialFESpace(Q0), sol[nu:])                  # the mean shift of FEFunction(::ZeroMeanFESpace, fv, dv)
    ph_i = g.interpolate(lambda x: x[:, 0] + 2 * x[:, 1], g.TrialFESpace(Q0))
    assert np.abs(ph.free_values - ph_i.free_values).max() < 1e-9 and abs(ph.dirichlet_values[0] - ph_i.dirichlet_values[0]) < 1e-9
    assert abs(np.dot(np.concatenate([ph.free_values, ph.dirichlet_values]), Q0._vol_i)) < 1e-10   # (:80-82)
