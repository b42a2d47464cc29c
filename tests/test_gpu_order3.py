"""Order-3 Lagrangian elements (Q3 on QUAD / HEX, P3 on TRI / TET) on the CUDA path -- the reference's assembly benchmark sweeps
orders 1-3 (benchmark/bm/bm_assembly.jl:35).

* Manufactured solutions in the spirit of test/GridapTests/PoissonTests.jl: a cubic u lies in the order-3 space of an AFFINE mesh, so
  the discrete solution of -Laplace(u) = f, u = g on the boundary is the interpolant of u -- this checks the face-frame numbering,
  the tabulation, the fused Dirichlet lifting and the scatter in one go, against an answer that does not come from the oracle.
* Entry-wise parity against the oracle on perturbed (non-affine) hexahedra and on tetrahedra, matrix and right-hand side with lifting.

Tolerance of the entry-wise comparisons: 1e-10 relative (see tests/test_gpu_bm_protocol.py: the oracle inverts the monomial
Vandermonde matrix as the reference does)."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

import gridap_b200 as g
from oracle import capi
from parity_helpers import check_csc, hex_model, oracle_field, oracle_problem, perturb, relerr, shear

pytestmark = pytest.mark.gpu


def _cubic(D):
    c = np.arange(1, D + 1, dtype=np.float64)

    def u(x):
        return (x ** 3) @ c + x[:, 0] * x[:, -1] ** 2 - 2.0 * x[:, 0] * x[:, -1] + 1.0

    def minus_laplace(x):
        return -(6.0 * (x @ c) + 2.0 * x[:, 0])

    return u, minus_laplace


@pytest.mark.parametrize("kind", ["QUAD", "TRI", "HEX", "TET"])
def test_order3_poisson_reproduces_a_cubic(kind):
    D = 2 if kind in ("QUAD", "TRI") else 3
    part = (4, 3) if D == 2 else (3, 2, 2)
    model = g.CartesianDiscreteModel((0, 1) * D, part)
    A_shear = np.array([[1.0, 0.3], [0.1, 0.8]]) if D == 2 else np.array([[1.0, 0.3, 0.1], [0.0, 0.8, 0.25], [0.2, 0.0, 1.3]])
    shear(model, A_shear, (0.5, -1.0, 2.0)[:D])                    # cells stay affine, the Jacobian is full
    if kind in ("TRI", "TET"):
        model = g.simplexify(model)
    u, f = _cubic(D)
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 3), dirichlet_tags="boundary")
    U = g.TrialFESpace(V, u)
    dO = g.Measure(g.Triangulation(model), 6 if kind in ("QUAD", "HEX") else 4)
    a = lambda du, v: g.Integral(g.inner(g.grad(v), g.grad(du))) * dO   # noqa: E731
    l = lambda v: g.Integral(v * f) * dO                                  # noqa: E731
    op = g.AffineFEOperator(a, l, U, V)
    x = spla.spsolve(op.get_matrix().to_scipy().tocsc(), op.get_vector())
    exact = V.interpolate_free_values(u)
    assert V.nfree > 0 and np.abs(x - exact).max() <= 1e-10 * np.abs(exact).max()


@pytest.mark.parametrize("ncomp", [1, 3])
def test_q3_perturbed_hexahedra_entrywise(ncomp):
    model = perturb(hex_model((3, 2, 3)), 0.2, 77)
    T = float if ncomp == 1 else g.VectorValue(3)
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, T, 3), dirichlet_tags=[21, 22])
    U = g.TrialFESpace(V, (lambda x: 1.0 + x[:, 0] * x[:, 1]) if ncomp == 1 else (lambda x: np.stack([x[:, 0], x[:, 1] ** 2, 1.0 + x[:, 2]], axis=1)))
    dO = g.Measure(g.Triangulation(model), 6)
    lap = lambda u, v: g.Integral(g.inner(g.grad(u), g.grad(v))) * dO   # noqa: E731
    A = g.assemble_matrix(lap, U, V)
    pb = oracle_problem(model, [oracle_field(model, V, 6)], 6, capi.LAPLACIAN, nrows=V.nfree, ncols=V.nfree)
    check_csc(A, pb.assemble(), tol=1e-10)
    if ncomp == 1:   # right-hand side with the Dirichlet lifting (a 64 x 64 local matrix per cell in shared memory)
        op = g.AffineFEOperator(lap, lambda v: g.Integral(v * 2.0) * dO, U, V)
        fld = oracle_field(model, V, 6, dirichlet_values=U.dirichlet_values)
        ref = oracle_problem(model, [fld], 6, capi.LAPLACIAN, capi.SOURCE, params=(2.0,), lift=True, nrows=V.nfree, ncols=V.nfree).assemble(with_vector=True)
        check_csc(op.get_matrix(), ref, tol=1e-10)
        assert relerr(op.get_vector(), ref[3]) <= 1e-10
    else:
        E = g.IsotropicLinearElasticity(2.0, 1.5)
        Ae = g.assemble_matrix(lambda u, v: g.Integral(g.inner(g.eps(v), E(g.eps(u)))) * dO, U, V)
        pbe = oracle_problem(model, [oracle_field(model, V, 6)], 6, capi.ELASTICITY, params=(2.0, 1.5), nrows=V.nfree, ncols=V.nfree)
        check_csc(Ae, pbe.assemble(), tol=1e-10)


@pytest.mark.parametrize("ncomp", [1, 3])
def test_p3_tetrahedra_entrywise(ncomp):
    model = g.simplexify(perturb(g.CartesianDiscreteModel((0, 1) * 3, (3, 2, 2)), 0.15, 5))
    T = float if ncomp == 1 else g.VectorValue(3)
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, T, 3), dirichlet_tags=[23])
    dO = g.Measure(g.Triangulation(model), 5)
    for form, a in ((capi.MASS, lambda u, v: g.Integral(g.dot(u, v)) * dO), (capi.LAPLACIAN, lambda u, v: g.Integral(g.inner(g.grad(u), g.grad(v))) * dO)):
        A = g.assemble_matrix(a, V, V)
        pb = oracle_problem(model, [oracle_field(model, V, 5)], 5, form, nrows=V.nfree, ncols=V.nfree)
        check_csc(A, pb.assemble(), tol=1e-10)
    b = g.assemble_vector(lambda v: g.Integral(g.dot(v, (lambda x: x[:, :ncomp] * 1.0) if ncomp == 3 else (lambda x: x[:, 0] + x[:, 2]))) * dO, V)
    geo = oracle_problem(model, [oracle_field(model, V, 5)], 5, capi.MASS, nrows=V.nfree, ncols=V.nfree)
    xq = geo.quadrature_points()
    fq = xq[:, :, :3] if ncomp == 3 else (xq[:, :, 0] + xq[:, :, 2])[:, :, None]
    rb = oracle_problem(model, [oracle_field(model, V, 5)], 5, 0, capi.SOURCE, fq=np.ascontiguousarray(fq), nrows=V.nfree, ncols=V.nfree).assemble_vector()
    assert relerr(b, rb) <= 1e-10


@pytest.mark.parametrize("kind", ["QUAD", "TRI", "HEX", "TET"])
def test_order3_neumann_boundary_reproduces_a_cubic(kind):
    """facet-wise DoF tables of an order-3 space (read from the adjacent cells): -Laplace(u) = f, u given on every side but the last
    one, n.grad(u) = g there (test/GridapTests/PoissonTests.jl:101-107 style); the cubic lies in the space"""
    D = 2 if kind in ("QUAD", "TRI") else 3
    model = g.CartesianDiscreteModel((0, 1) * D, (4, 3) if D == 2 else (3, 2, 2))
    neumann = 6 if D == 2 else 22                                  # the side x_D = 1 (its interior; its boundary is Dirichlet)
    dirichlet = [t for t in range(1, 9 if D == 2 else 27) if t != neumann]
    if kind in ("TRI", "TET"):
        model = g.simplexify(model)
    c = np.arange(1, D + 1, dtype=np.float64)
    u = lambda x: (x ** 3) @ c + x[:, 0] * x[:, -1] ** 2 + 1.0            # noqa: E731
    f = lambda x: -(6.0 * (x @ c) + 2.0 * x[:, 0])                         # noqa: E731
    gN = lambda x: 3.0 * c[-1] * x[:, -1] ** 2 + 2.0 * x[:, 0] * x[:, -1]  # noqa: E731  (d u / d x_D; on the side: 3 c_D + 2 x_1)
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 3), dirichlet_tags=dirichlet)
    U = g.TrialFESpace(V, u)
    deg = 6 if kind in ("QUAD", "HEX") else 4
    dO = g.Measure(g.Triangulation(model), deg)
    dG = g.Measure(g.BoundaryTriangulation(model, tags=[neumann]), deg)
    a = lambda du, v: g.Integral(g.inner(g.grad(v), g.grad(du))) * dO      # noqa: E731
    l = lambda v: g.Integral(v * f) * dO + g.Integral(v * gN) * dG         # noqa: E731
    op = g.AffineFEOperator(a, l, U, V)
    x = spla.spsolve(op.get_matrix().to_scipy().tocsc(), op.get_vector())
    exact = V.interpolate_free_values(u)
    fx = V.dof_coordinates()[0]
    assert np.any(np.isclose(fx[:, -1], 1.0))                               # free DoFs on the Neumann side
    assert np.abs(x - exact).max() <= 1e-10 * np.abs(exact).max()
