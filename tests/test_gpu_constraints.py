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
