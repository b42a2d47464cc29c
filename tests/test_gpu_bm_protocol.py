"""The reference's own assembly benchmark protocol (benchmark/bm/bm_assembly.jl:7-57) on the CUDA path, ENTRY BY ENTRY against the oracle:

    for (D, n) in [(2, 10), (3, 6)], model = UnstructuredDiscreteModel(CartesianDiscreteModel(domain, partition))
      for trian in [Triangulation(model), Triangulation(model, collect(1:div(n^D, 2)))]          # "bulk", "view"
        for order in [1, 2, 3]
          for T in [Float64, VectorValue{D,Float64}]                                             # (raviart_thomas: out of scope)
            for (biform, qdegree) in [(mass, 2*order), (laplacian, 2*(order-1)), (graddiv, 2*(order-1))]
              A = assemble_matrix((u, v) -> biform(u, v, Measure(trian, qdegree)), V, V),   V = TestFESpace(model, reffe)

graddiv = (div u)(div v) needs a vector-valued field.  The 3-D vector-valued order-3 cases run on n = 3 (27 cells; view = the first 13) so that the
single-threaded oracle stays within seconds (a Q3 vector hexahedron has 192 x 192 local entries at 64 points).

Tolerance: pattern bit-exact; max|d nzval| / max|nzval| <= 1e-12 for orders 1 and 2, 1e-10 for order 3 -- the oracle obtains the
shape functions as the reference does, by inverting the monomial Vandermonde matrix at the nodes (src/ReferenceFEs/ReferenceFEInterfaces.jl:
563-583), which costs ~1e-12 absolute on the Q3 hexahedron (64 x 64, tests/test_host_logic.py::test_order3_tabulation), while the
product tabulates the same polynomials in product form."""
import numpy as np
import pytest

import gridap_b200 as g
from oracle import capi
from parity_helpers import check_csc, oracle_field, oracle_problem

pytestmark = pytest.mark.gpu


def bm_cases():
    out = []
    for D, n in ((2, 10), (3, 6)):
        for trian_name in ("bulk", "view"):
            for order in (1, 2, 3):
                for vector in (False, True):
                    for biform, qdegree in (("mass", 2 * order), ("laplacian", 2 * (order - 1)), ("graddiv", 2 * (order - 1))):
                        if biform == "graddiv" and not vector:
                            continue
                        out.append((D, 3 if (D == 3 and order == 3 and vector) else n, trian_name, order, vector, biform, qdegree))
    return out


def bm_model(D, n):
    return g.UnstructuredDiscreteModel(g.CartesianDiscreteModel((0, 1) * D, (n,) * D))


def bm_view_cells(D, n):
    return np.arange(1, n ** D // 2 + 1)      # collect(1:div(n^D,2)), 1-based


def bm_oracle(model, V, trian_name, n, order, biform, qdegree):
    D = model.D
    if biform == "graddiv":
        form, params = capi.ELASTICITY, (1.0, 0.0)        # sigma = lambda tr(eps) I, mu = 0: (div u)(div v)
    else:
        form, params = (capi.MASS if biform == "mass" else capi.LAPLACIAN), (1.0,)
    ids = V.cell_dof_ids
    sub = model
    if trian_name == "view":
        cells = bm_view_cells(D, n) - 1
        sub = g.DiscreteModel(model.node_coordinates, model.cell_node_ids[cells], model.ptype)
        ids = ids[cells]
    pb = oracle_problem(sub, [oracle_field(sub, V, qdegree, ids=ids)], qdegree, form, params=params, nrows=V.nfree, ncols=V.nfree)
    return pb.assemble()


def bm_form(biform, dO):
    if biform == "mass":
        return lambda u, v: g.Integral(g.dot(u, v)) * dO
    if biform == "laplacian":
        return lambda u, v: g.Integral(g.inner(g.grad(u), g.grad(v))) * dO
    return lambda u, v: g.Integral(g.div(u) * g.div(v)) * dO


@pytest.mark.parametrize("case", bm_cases(), ids=lambda c: "%dD_n%d_%s_o%d_%s_%s" % (c[0], c[1], c[2], c[3], "vec" if c[4] else "sca", c[5]))
def test_bm_assembly_case(case):
    D, n, trian_name, order, vector, biform, qdegree = case
    model = bm_model(D, n)
    T = g.VectorValue(D) if vector else float
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, T, order))
    trian = g.Triangulation(model) if trian_name == "bulk" else g.Triangulation(model, bm_view_cells(D, n))
    dO = g.Measure(trian, qdegree)
    A = g.assemble_matrix(bm_form(biform, dO), V, V)
    assert A.shape == (V.nfree, V.nfree)
    ref = bm_oracle(model, V, trian_name, n, order, biform, qdegree)
    check_csc(A, ref, tol=1e-10 if order == 3 else 1e-12)


def test_view_plus_bulk_and_vector_on_a_view():
    """a form over the bulk AND a view (the view's matrix is merged into the bulk pattern), and assemble_vector on a view"""
    model = bm_model(3, 5)
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 2), dirichlet_tags=[22])
    U = g.TrialFESpace(V, 0.0)
    cells = np.array([3, 4, 17, 60, 61, 62, 100, 125])
    dO, dS = g.Measure(g.Triangulation(model), 4), g.Measure(g.Triangulation(model, cells), 4)
    A = g.assemble_matrix(lambda u, v: g.Integral(g.inner(g.grad(u), g.grad(v))) * dO + g.Integral(3.0 * (u * v)) * dS, U, V)
    b = g.assemble_vector(lambda v: g.Integral(v * 2.0) * dS, V)
    lap = oracle_problem(model, [oracle_field(model, V, 4)], 4, capi.LAPLACIAN, nrows=V.nfree, ncols=V.nfree).assemble()
    sub = g.DiscreteModel(model.node_coordinates, model.cell_node_ids[cells - 1], model.ptype)
    fld = oracle_field(sub, V, 4, ids=V.cell_dof_ids[cells - 1])
    mass = oracle_problem(sub, [fld], 4, capi.MASS, nrows=V.nfree, ncols=V.nfree).assemble()
    import scipy.sparse as sp
    to_sp = lambda r: sp.csc_matrix((r[2], r[1] - 1, r[0] - 1), shape=(V.nfree, V.nfree))   # noqa: E731
    refA = (to_sp(lap) + 3.0 * to_sp(mass)).toarray()
    assert np.array_equal(A.colptr, lap[0]) and np.array_equal(A.rowval, lap[1])         # the bulk pattern
    assert np.abs(A.to_scipy().toarray() - refA).max() <= 1e-12 * np.abs(refA).max()
    rb = oracle_problem(sub, [oracle_field(sub, V, 4, ids=V.cell_dof_ids[cells - 1])], 4, 0, capi.SOURCE, params=(2.0,),
                        nrows=V.nfree, ncols=V.nfree).assemble_vector()
    assert np.abs(b - rb).max() <= 1e-12 * np.abs(rb).max()


def test_affine_operator_on_a_view_only():
    """AffineFEOperator whose forms live on a view only: the view's plan owns the pattern (columns of DoFs outside the view stay
    empty), the Dirichlet lifting runs on the view's cells"""
    model = bm_model(3, 4)
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags=[25])
    U = g.TrialFESpace(V, lambda x: 1.0 + x[:, 1] * x[:, 2])
    cells = bm_view_cells(3, 4)
    dS = g.Measure(g.Triangulation(model, cells), 2)
    op = g.AffineFEOperator(lambda u, v: g.Integral(g.inner(g.grad(u), g.grad(v))) * dS, lambda v: g.Integral(v * 2.0) * dS, U, V)
    sub = g.DiscreteModel(model.node_coordinates, model.cell_node_ids[cells - 1], model.ptype)
    fld = oracle_field(sub, V, 2, ids=V.cell_dof_ids[cells - 1], dirichlet_values=U.dirichlet_values)
    ref = oracle_problem(sub, [fld], 2, capi.LAPLACIAN, capi.SOURCE, params=(2.0,), lift=True, nrows=V.nfree, ncols=V.nfree).assemble(with_vector=True)
    check_csc(op.get_matrix(), ref)
    assert np.abs(op.get_vector() - ref[3]).max() <= 1e-12 * np.abs(ref[3]).max()
    assert np.any(np.diff(ref[0]) == 0)      # (there are empty columns)
