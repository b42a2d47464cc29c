"""CUDA path vs CPU oracle, ENTRY BY ENTRY, at sizes that exercise the real control flow of the kernels:

* config 2 at 96^3 (857 375 columns = 26 793 column blocks > 11 trips of every persistent warp: next-block metadata prefetch,
  run-length rows), axis-aligned (3-factor diagonal metric and the general 6-factor instance) / sheared-affine / perturbed
  geometry, Laplacian and mass, right-hand side with lifting;
* configs 3 / 4 / 5 on PERTURBED meshes whose cell counts are not multiples of the CTA batches (8 hexes / 7 tets), atomic and
  coloured scatter: Q2 elasticity 9x8x7 (DMMA kernel), Taylor-Hood on 5x6x6x6 tets, neo-Hookean on 17x15x15 hexes.

Tolerance: pattern bit-exact, max|d nzval| / max|nzval| <= 1e-12 (north star)."""
import numpy as np
import pytest

import gridap_b200 as g
from gridap_b200 import lib
from oracle import capi
from parity_helpers import LAM, MU, check_csc, hex_model, oracle_field, oracle_problem, perturb, relerr, shear
from parity_helpers import env as _env

pytestmark = pytest.mark.gpu


PIPELINES = {
    "diag": {},                                  # axis-aligned cells: diagonal metric, 3 factors per cell
    "full": dict(GB200_GATHER_DIAG=0),           # the same mesh through the general affine instance (6 factors per cell)
    "diag5": dict(GB200_GATHER_DIAG_MINB5=1),    # (5 CTAs per SM instance of the diagonal kernel)
}

_oracle_cache = {}


def _q1_oracle(n, form, geometry):
    key = (n, form, geometry)
    if key not in _oracle_cache:
        model = _q1_model(n, geometry)
        V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags="boundary")
        pb = oracle_problem(model, [oracle_field(model, V, 2)], 2, form, nrows=V.nfree, ncols=V.nfree)
        _oracle_cache[key] = pb.assemble()
    return _oracle_cache[key]


def _q1_model(n, geometry):
    model = hex_model((n, n, n))
    if geometry == "perturbed":
        perturb(model, 0.2, 12345)
    elif geometry == "sheared":
        shear(model, [[1.0, 0.3, 0.1], [0.0, 0.8, 0.25], [0.2, 0.0, 1.3]], (0.5, -1.0, 2.0))
    return model


@pytest.mark.parametrize("form", ["laplacian", "mass"])
@pytest.mark.parametrize("pipeline", list(PIPELINES))
def test_config2_96_affine_entrywise(form, pipeline):
    n = 96
    fid = capi.LAPLACIAN if form == "laplacian" else capi.MASS
    ref = _q1_oracle(n, fid, "affine")
    model = _q1_model(n, "affine")
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags="boundary")
    U = g.TrialFESpace(V, 0.0)
    dO = g.Measure(g.Triangulation(model), 2)
    a = (lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dO) if form == "laplacian" else (lambda u, v: g.Integral(u * v) * dO)
    with _env(**PIPELINES[pipeline]):
        assem = g.SparseMatrixAssembler(U, V)
        A = g.assemble_matrix(a, assem, U, V)
        assert assem.plan(dO).kernel_path(fid) == ("q1hex_gather_affine+diag" if form == "laplacian" and pipeline != "full" else "q1hex_gather_affine")
        check_csc(A, ref)
        # in-place re-assembly and _add! through the same kernels
        matdata = g.collect_cell_matrix(U, V, a(g.get_trial_fe_basis(U), g.get_fe_basis(V)))
        assem.assemble_matrix_(A, matdata)
        assert relerr(A.nzval, ref[2]) <= 1e-12
        assem.assemble_matrix_add_(A, matdata)
        assert relerr(A.nzval, 2.0 * ref[2]) <= 1e-12


@pytest.mark.parametrize("geometry,path", [("sheared", "q1hex_gather_affine"), ("perturbed", "q1hex_gather_general")])
@pytest.mark.parametrize("form", ["laplacian", "mass"])
def test_config2_non_cartesian_geometry_entrywise(geometry, path, form):
    n = 96 if geometry == "perturbed" else 64
    fid = capi.LAPLACIAN if form == "laplacian" else capi.MASS
    ref = _q1_oracle(n, fid, geometry)
    model = _q1_model(n, geometry)
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags="boundary")
    dO = g.Measure(g.Triangulation(model), 2)
    a = (lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dO) if form == "laplacian" else (lambda u, v: g.Integral(u * v) * dO)
    assem = g.SparseMatrixAssembler(V, V)
    A = g.assemble_matrix(a, assem, V, V)
    assert assem.plan(dO).kernel_path(fid) == path
    check_csc(A, ref)


@pytest.mark.parametrize("geometry", ["affine", "sheared", "perturbed"])
def test_config2_96_matrix_and_rhs_with_lifting(geometry):
    # AffineFEOperator: f(x) at the quadrature points, inhomogeneous Dirichlet data, fused lifting (config 2 is "matrix + RHS");
    # affine = axis-aligned boxes (box instances of the geometry / RHS kernels), sheared = affine cells with a full Jacobian,
    # perturbed = general trilinear cells
    n = 96 if geometry != "sheared" else 48
    model = _q1_model(n, geometry)
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags="boundary")
    gfun = lambda x: np.sin(3.0 * x[:, 0]) + x[:, 1] * x[:, 2]   # noqa: E731
    ffun = lambda x: 1.0 + x[:, 0] - 2.0 * x[:, 1] * x[:, 2]     # noqa: E731
    U = g.TrialFESpace(V, gfun)
    dO = g.Measure(g.Triangulation(model), 2)
    op = g.AffineFEOperator(lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dO, lambda v: g.Integral(v * ffun) * dO, U, V)
    pb0 = oracle_problem(model, [oracle_field(model, V, 2)], 2, capi.LAPLACIAN, nrows=V.nfree, ncols=V.nfree)
    xq = pb0.quadrature_points()
    fq = ffun(xq.reshape(-1, 3)).reshape(xq.shape[:2])
    pb = oracle_problem(model, [oracle_field(model, V, 2, dirichlet_values=U.dirichlet_values)], 2, capi.LAPLACIAN, capi.SOURCE, fq=fq, lift=True,
                        nrows=V.nfree, ncols=V.nfree)
    colptr, rowval, nzval, b = pb.assemble(with_vector=True)
    check_csc(op.get_matrix(), (colptr, rowval, nzval))
    assert relerr(op.get_vector(), b) <= 1e-12
    # assemble_vector alone (no lifting)
    bv = g.assemble_vector(lambda v: g.Integral(v * ffun) * dO, V)
    pbv = oracle_problem(model, [oracle_field(model, V, 2)], 2, 0, capi.SOURCE, fq=fq, nrows=V.nfree, ncols=V.nfree)
    assert relerr(bv, pbv.assemble_vector()) <= 1e-12


X0_TAGS = [25, 1, 3, 5, 7, 13, 15, 17, 19]   # face x = 0 of the box and its closure


@pytest.mark.parametrize("deterministic", [False, True])
def test_config3_q2_elasticity_perturbed_9x8x7(deterministic):
    part = (9, 8, 7)
    model = perturb(g.CartesianDiscreteModel((0, 1) * 3, part), 0.15, 31)
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 2), dirichlet_tags=X0_TAGS)
    U = g.TrialFESpace(V, (0.0, 0.0, 0.0))
    dO = g.Measure(g.Triangulation(model), 4)
    sigma = g.IsotropicLinearElasticity(LAM, MU)
    assem = g.SparseMatrixAssembler(U, V, deterministic=deterministic)
    A = g.assemble_matrix(lambda u, v: g.Integral(g.inner(g.eps(v), sigma(g.eps(u)))) * dO, assem, U, V)
    assert assem.plan(dO).kernel_path(lib.FORM_ELASTICITY).endswith("+dmma")
    if "q2" not in _oracle_cache:
        pb = oracle_problem(model, [oracle_field(model, V, 4)], 4, capi.ELASTICITY, params=[LAM, MU], nrows=V.nfree, ncols=V.nfree)
        _oracle_cache["q2"] = pb.assemble()
    check_csc(A, _oracle_cache["q2"])


@pytest.mark.parametrize("deterministic", [False, True])
def test_config4_stokes_tets_perturbed_5x6x6(deterministic):
    part = (5, 6, 6)
    model = g.simplexify(perturb(g.CartesianDiscreteModel((0, 1) * 3, part), 0.15, 32))
    assert model.num_cells() % 7 != 0
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 2), dirichlet_tags="boundary")
    Q = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1))
    Y = g.MultiFieldFESpace([V, Q])
    X = g.MultiFieldFESpace([g.TrialFESpace(V, (0.0, 0.0, 0.0)), g.TrialFESpace(Q)])
    dO = g.Measure(g.Triangulation(model), 4)

    def a(up, vq):
        (u, p), (v, q) = up, vq
        return g.Integral(g.inner(g.grad(v), g.grad(u)) - g.div(v) * p + q * g.div(u)) * dO

    assem = g.SparseMatrixAssembler(X, Y, deterministic=deterministic)
    A = g.assemble_matrix(a, assem, X, Y)
    assert assem.plan(dO, assem._touched([type("T", (), {"form": lib.FORM_STOKES})()])).kernel_path(lib.FORM_STOKES).startswith("affine_gather")   # simplices: always affine
    with _env(GB200_NO_AFFINE_GATHER=1):   # the cell-centric node-pair kernel on the same mesh
        assem2 = g.SparseMatrixAssembler(X, Y, deterministic=deterministic)
        A_cell = g.assemble_matrix(a, assem2, X, Y)
    with _env(GB200_NO_BLOCK_GATHER=1):    # the column-node kernel
        A_col = g.assemble_matrix(a, g.SparseMatrixAssembler(X, Y, deterministic=deterministic), X, Y)
    assert relerr(A_col.nzval, A.nzval) <= 1e-13   # same closed form up to the association of the factors
    ids = Y.get_cell_dof_ids()
    fu = oracle_field(model, V, 4, 0, ids=ids[0])
    fp = oracle_field(model, Q, 4, V.nfree, ids=ids[1])
    n = V.nfree + Q.nfree
    if "stokes" not in _oracle_cache:
        pb = oracle_problem(model, [fu, fp], 4, capi.STOKES, touched=np.array([[1, 1], [1, 0]], dtype=np.uint8), nrows=n, ncols=n)
        _oracle_cache["stokes"] = pb.assemble()
    check_csc(A, _oracle_cache["stokes"])
    check_csc(A_cell, _oracle_cache["stokes"])


@pytest.mark.parametrize("deterministic,scatter", [(False, "staged"), (True, "staged"), (False, "cells"), (True, "cells")])
def test_config5_neohookean_perturbed_17x15x15(deterministic, scatter):
    # staged: node-pair blocks through HBM + block-owner gather (default); cells: RED / coloured scatter of the cell-centric kernel
    with _env(**({"GB200_NO_STAGED_GATHER": 1} if scatter == "cells" else {})):
        _config5_neohookean_perturbed(deterministic, "staged_gather+blocks" if scatter == "staged" else ("vector_coloured" if deterministic else "vector_atomic"))


def _config5_neohookean_perturbed(deterministic, path):
    part = (17, 15, 15)
    model = perturb(hex_model(part), 0.15, 33)
    assert model.num_cells() % 8 != 0
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 1), dirichlet_tags="boundary")
    U = g.TrialFESpace(V, (0.0, 0.0, 0.0))
    dO = g.Measure(g.Triangulation(model), 2)
    nh = g.NeoHookean(100.0, 1.0)
    ufun = lambda x: 0.05 * (np.sin(np.pi * x[:, 0]) * np.sin(np.pi * x[:, 1]) * np.sin(np.pi * x[:, 2]))[:, None] * np.array([[1.0, -0.5, 0.75]])  # noqa: E731
    uh = g.interpolate(ufun, U)
    assem = g.SparseMatrixAssembler(U, V, deterministic=deterministic)
    op = g.FEOperator(lambda u, v: g.Integral(nh.res(u, v)) * dO, lambda u, du, v: g.Integral(nh.jac(u, du, v)) * dO, U, V, assem)
    pb = oracle_problem(model, [oracle_field(model, V, 2, free_values=uh.free_values, dirichlet_values=uh.dirichlet_values)], 2, capi.NEOHOOKEAN_JAC,
                        capi.NEOHOOKEAN_RES, params=[100.0, 1.0], nrows=V.nfree, ncols=V.nfree)
    if "nh" not in _oracle_cache:
        _oracle_cache["nh"] = pb.assemble(with_vector=True)
    colptr, rowval, nzval, bo = _oracle_cache["nh"]
    check_csc(op.jacobian(uh), (colptr, rowval, nzval))
    assert relerr(op.residual(uh), bo) <= 1e-12
    b3, A3 = op.residual_and_jacobian(uh)
    check_csc(A3, (colptr, rowval, nzval))
    assert relerr(b3, bo) <= 1e-12
    assert assem.plan(dO).kernel_path(lib.FORM_NEOHOOKEAN_JAC) == path


def test_alternating_plans_of_different_sizes():
    # two live Q1 assemblers of different sizes used in turn: the dynamic shared memory opt-in of the gather kernel belongs to the
    # function, not to a plan (it used to be lowered by the smaller plan)
    big, small = _q1_model(40, "affine"), _q1_model(2, "affine")
    out = []
    for _ in range(2):
        for model in (big, small, big):
            V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags="boundary")
            dO = g.Measure(g.Triangulation(model), 2)
            out.append(g.assemble_matrix(lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dO, V, V))
    assert np.array_equal(out[0].nzval, out[2].nzval) and np.array_equal(out[0].nzval, out[5].nzval)
    assert out[1].nnz() == 1


SHEAR = [[1.0, 0.3, 0.1], [0.0, 0.8, 0.25], [0.2, 0.0, 1.3]]


@pytest.mark.parametrize("engine", ["pairs", "blocks", "columns"])   # thread per mirror pair of stored node-pair blocks / per block / warp per column node
@pytest.mark.parametrize("order,part", [(2, (7, 6, 5)), (1, (11, 9, 10))])
@pytest.mark.parametrize("form", ["elasticity", "laplacian", "mass"])
def test_affine_gather_vector_hexes_sheared(order, part, form, engine):
    with _env(**({"GB200_NO_BLOCK_GATHER": 1} if engine == "columns" and order == 1 else {"GB200_MIRROR": 1} if engine == "pairs" else {})):   # (order 2: automatic fallback)
        _affine_gather_vector_hexes_sheared(order, part, form, engine)


def _affine_gather_vector_hexes_sheared(order, part, form, engine):
    # owner-computes column-node gather on affine cells with a FULL Jacobian (parallelepipeds), vector-valued Q1 / Q2, with a
    # component-wise Dirichlet mask (only some components of the tagged nodes are constrained: partial column / row groups)
    model = shear(g.CartesianDiscreteModel((0, 1) * 3, part), SHEAR, (0.5, -1.0, 2.0))
    # blocks: the free components of every node are contiguous; columns: face x = 1 keeps components 0 and 2 free (a gap): the
    # block plan detects the irregular blocks and the column-node kernel takes over on its own
    masks = [(True, False, True), (True, True, False)] if engine != "columns" else [(True, False, True), (False, True, False)]
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), order), dirichlet_tags=[X0_TAGS[0], 26], dirichlet_masks=masks)
    dO = g.Measure(g.Triangulation(model), 2 * order)
    if form == "elasticity":
        sigma = g.IsotropicLinearElasticity(LAM, MU)
        a, fid, prm = (lambda u, v: g.Integral(g.inner(g.eps(v), sigma(g.eps(u)))) * dO), capi.ELASTICITY, [LAM, MU]
    elif form == "laplacian":
        a, fid, prm = (lambda u, v: g.Integral(1.5 * g.inner(g.grad(v), g.grad(u))) * dO), capi.LAPLACIAN, [1.5]
    else:
        a, fid, prm = (lambda u, v: g.Integral(g.dot(u, v)) * dO), capi.MASS, [1.0]
    assem = g.SparseMatrixAssembler(V, V)
    A = g.assemble_matrix(a, assem, V, V)
    assert assem.plan(dO).kernel_path(fid) == "affine_gather+" + engine
    pb = oracle_problem(model, [oracle_field(model, V, 2 * order)], 2 * order, fid, params=prm, nrows=V.nfree, ncols=V.nfree)
    ref = pb.assemble()
    if form == "laplacian":   # (the oracle's Laplacian carries no coefficient)
        ref = (ref[0], ref[1], 1.5 * ref[2])
    check_csc(A, ref)
    matdata = g.collect_cell_matrix(V, V, a(g.get_trial_fe_basis(V), g.get_fe_basis(V)))
    assem.assemble_matrix_add_(A, matdata)          # _add! through the gather (out += columns)
    assert relerr(A.nzval, 2.0 * ref[2]) <= 1e-12
    A2 = g.assemble_matrix(a, g.SparseMatrixAssembler(V, V), V, V)
    assert np.array_equal(A2.nzval, 0.5 * A.nzval) or relerr(A2.nzval, ref[2]) <= 1e-12
    A3 = g.assemble_matrix(a, g.SparseMatrixAssembler(V, V, deterministic=True), V, V)
    assert np.array_equal(A3.nzval, A2.nzval)       # bitwise reproducible, independent of the scatter mode


@pytest.mark.parametrize("ptype,order", [("TET", 2), ("TET", 1), ("HEX", 2)])
def test_affine_gather_scalar_elements(ptype, order):
    part = (5, 4, 6)
    model = shear(g.CartesianDiscreteModel((0, 1) * 3, part), SHEAR)
    if ptype == "TET":
        model = g.simplexify(model)
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, order), dirichlet_tags=[25, 22])
    dO = g.Measure(g.Triangulation(model), 2 * order)
    for fid, a in ((capi.LAPLACIAN, lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dO), (capi.MASS, lambda u, v: g.Integral(u * v) * dO)):
        assem = g.SparseMatrixAssembler(V, V)
        A = g.assemble_matrix(a, assem, V, V)
        assert assem.plan(dO).kernel_path(fid).startswith("affine_gather")
        pb = oracle_problem(model, [oracle_field(model, V, 2 * order)], 2 * order, fid, nrows=V.nfree, ncols=V.nfree)
        check_csc(A, pb.assemble())
