"""N3 of SURVEY 8(f): boundary terms that need the cell adjacent to a facet -- the unit normal (get_normal_vector,
src/Geometry/BoundaryTriangulations.jl:244-283) and the cell basis / its gradient at the facet quadrature points (FaceToCellGlue,
:13-70) -- i.e. the Nitsche and normal-flux terms of test/GridapTests/PoissonTests.jl:99-107.  Entry-wise parity with the oracle's
restatement on perturbed hexahedral / tetrahedral meshes, and the reference's own manufactured-solution driver."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

import gridap_b200 as g
from oracle import capi
from oracle import ref_tabulation as rt
from parity_helpers import glued_facet_problem, oracle_field, oracle_problem, perturb, relerr

pytestmark = pytest.mark.gpu


def _to_dense(ref, n):
    import scipy.sparse as sp
    return sp.csc_matrix((ref[2], ref[1] - 1, ref[0] - 1), shape=(n, n))


@pytest.mark.parametrize("ptype,order,ncomp", [("HEX", 1, 1), ("HEX", 2, 1), ("TET", 2, 1), ("TET", 1, 3), ("HEX", 1, 3)])
def test_facet_of_cell_terms_against_the_oracle(ptype, order, ncomp):
    model = perturb(g.CartesianDiscreteModel((0, 1) * 3, (4, 3, 3)), 0.15, 17)
    if ptype == "TET":
        model = g.simplexify(model)
    T = float if ncomp == 1 else g.VectorValue(ncomp)
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, T, order), dirichlet_tags=[25])
    deg = 2 * order
    dO = g.Measure(g.Triangulation(model), deg)
    Gam = g.BoundaryTriangulation(model, tags=[26, 22])   # faces x = 1 and z = 1
    dG = g.Measure(Gam, deg)
    n = g.get_normal_vector(Gam)
    gam = 7.5
    bulk = (lambda u, v: g.inner(g.grad(v), g.grad(u)))
    mul = (lambda a, b: a * b) if ncomp == 1 else g.inner

    def a(u, v):
        return g.Integral(bulk(u, v)) * dO + g.Integral(gam * mul(v, u) - mul(v, g.dot(n, g.grad(u))) - mul(g.dot(n, g.grad(v)), u)) * dG

    A = g.assemble_matrix(a, V, V)
    nf = V.nfree
    ref = _to_dense(oracle_problem(model, [oracle_field(model, V, deg)], deg, capi.LAPLACIAN, nrows=nf, ncols=nf).assemble(), nf)
    for prm in ([gam, 0, 0], [-1.0, 0, 1], [-1.0, 1, 0]):
        ref = ref + _to_dense(glued_facet_problem(Gam, V, deg, form_mat=capi.FACET, params=prm).assemble(), nf)
    S = A.to_scipy()
    assert abs(S - ref).max() <= 1e-12 * abs(ref).max()
    # vector terms: (n.grad v) g,  v u_h,  (n.grad v) u_h,  v (n.grad u_h)   (PoissonTests.jl:103-107)
    U = g.TrialFESpace(V, (lambda x: 1.0 + x[:, 1]) if ncomp == 1 else (lambda x: np.stack([x[:, 0], 1.0 + x[:, 1], x[:, 2] ** 2], axis=1)))
    ufun = (lambda x: np.sin(x[:, 0]) + x[:, 1] * x[:, 2]) if ncomp == 1 else (lambda x: np.stack([np.sin(x[:, 0]), x[:, 1] * x[:, 2], x[:, 0] + x[:, 2]], axis=1))
    gfun = (lambda x: 1.0 + x[:, 0] * x[:, 1]) if ncomp == 1 else (lambda x: np.stack([1.0 + x[:, 0], x[:, 1], -x[:, 2]], axis=1))
    uh = g.interpolate(ufun, U)

    def l(v):
        return g.Integral(gam * mul(v, uh) - mul(g.dot(n, g.grad(v)), uh) + mul(v, g.dot(n, g.grad(uh))) - 0.5 * mul(g.dot(n, g.grad(v)), gfun)) * dG

    b = g.assemble_vector(l, V)
    npf = len(rt.quadrature("TRI" if ptype == "TET" else "QUAD", deg)[1])
    kw = dict(free_values=uh.free_values, dirichlet_values=uh.dirichlet_values)
    bo = glued_facet_problem(Gam, V, deg, form_vec=capi.FACET_VEC, params=[gam, 0, 1], **kw).assemble_vector()
    bo += glued_facet_problem(Gam, V, deg, form_vec=capi.FACET_VEC, params=[-1.0, 1, 1], **kw).assemble_vector()
    bo += glued_facet_problem(Gam, V, deg, form_vec=capi.FACET_VEC, params=[1.0, 0, 2], **kw).assemble_vector()
    xq = glued_facet_problem(Gam, V, deg).quadrature_points()
    fq = np.asarray(gfun(xq.reshape(-1, 3))).reshape(xq.shape[0], npf, ncomp)
    bo += glued_facet_problem(Gam, V, deg, form_vec=capi.FACET_VEC, params=[-0.5, 1, 0], fq=fq).assemble_vector()
    assert relerr(b, bo) <= 1e-12


@pytest.mark.parametrize("vector_valued", [False, True])
def test_poisson_nitsche_manufactured_solution(vector_valued):
    # test/GridapTests/PoissonTests.jl:9-124 (scalar and vector-valued data sets): Q2 on the 4x4 mesh of (0,1)^2, Dirichlet on tags
    # [1,2,5], Neumann on [7,8], Nitsche on 6, gamma = 10, degree = order; the reference asserts el2/ul2 < 1e-8, eh1/uh1 < 1e-7
    # (u is in the FE space: u_h is its interpolant)
    model = g.CartesianDiscreteModel((0, 1, 0, 1), (4, 4))
    order, h, gam = 2, 0.25, 10.0
    if vector_valued:
        u = lambda x: np.stack([x[:, 0] ** 2 + x[:, 1], 4 * x[:, 0] - x[:, 1] ** 2], axis=1)   # noqa: E731
        f = lambda x: np.stack([-2.0 + 0 * x[:, 0], 2.0 + 0 * x[:, 0]], axis=1)                  # noqa: E731  (-Laplace u)
        T, mul = g.VectorValue(2), g.inner
    else:
        u = lambda x: x[:, 0] ** 2 + x[:, 1]   # noqa: E731
        f = lambda x: -2.0 + 0 * x[:, 0]       # noqa: E731
        T, mul = float, (lambda a, b: a * b)
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, T, order), dirichlet_tags=[1, 2, 5])
    U = g.TrialFESpace(V, u)
    uh = g.interpolate(u, U)
    dO = g.Measure(g.Triangulation(model), order)
    Gn, Gd = g.BoundaryTriangulation(model, tags=[7, 8]), g.BoundaryTriangulation(model, tags=[6])
    dGn, dGd = g.Measure(Gn, order), g.Measure(Gd, order)
    nn, nd = g.get_normal_vector(Gn), g.get_normal_vector(Gd)

    def a(uu, v):
        return g.Integral(g.inner(g.grad(v), g.grad(uu))) * dO + \
            g.Integral((gam / h) * mul(v, uu) - mul(v, g.dot(nd, g.grad(uu))) - mul(g.dot(nd, g.grad(v)), uu)) * dGd

    def l(v):
        return g.Integral(mul(v, f)) * dO + g.Integral(mul(v, g.dot(nn, g.grad(uh)))) * dGn + \
            g.Integral((gam / h) * mul(v, uh) - mul(g.dot(nd, g.grad(v)), uh)) * dGd

    op = g.AffineFEOperator(a, l, U, V)
    x = spla.spsolve(op.get_matrix().to_scipy().tocsc(), op.get_vector())
    assert np.abs(x - uh.free_values).max() <= 1e-9 * np.abs(uh.free_values).max()
