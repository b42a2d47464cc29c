"""Edge cases and error behaviour of the C ABI (the `@check` / `@notimplemented` failures of the reference become
status codes + gb200_last_error; nothing ever falls back to the CPU)."""
import ctypes as C

import numpy as np
import pytest

import gridap_b200 as g
from gridap_b200 import lib
from gridap_b200 import reffes as rf

pytestmark = pytest.mark.gpu


def _tab(ptype="HEX", order=1, degree=2):
    xq, w = rf.Quadrature(ptype, degree)
    N, dN = rf.tabulate_lagrangian(ptype, order, xq)
    return w, N, dN


def _raw_space(ctx, mesh, refel, data, ptrs, nfree, ndir):
    h = C.c_void_p()
    rc = lib.load().gb200_space_create(ctx.h, mesh.h, refel.h, data.ctypes.data_as(C.c_void_p), ptrs.ctypes.data_as(C.c_void_p), nfree, ndir, C.byref(h))
    return rc, h


def test_error_codes_and_messages():
    ctx = lib.default_context(0)
    model = g.CartesianDiscreteModel((0, 1) * 3, (2, 2, 2))
    # wrong cell type for the dimension (QUAD4 / TRI3 in 3D and SEG2 in 2D are legal: boundary facets; SEG2 in 3D is not)
    with pytest.raises(lib.GridapB200Error) as e:
        lib.DeviceMesh(ctx, model.node_coordinates, model.cell_node_ids[:, :2], lib.SEG2)
    assert e.value.code == lib.ERR_INVALID
    m2 = g.CartesianDiscreteModel((0, 1, 0, 1), (2, 2))
    with pytest.raises(lib.GridapB200Error) as e:
        lib.DeviceMesh(ctx, m2.node_coordinates, np.tile(m2.cell_node_ids, (1, 2)), lib.HEX8)
    assert e.value.code == lib.ERR_INVALID
    # cells that do not have the node count of the declared type
    with pytest.raises(lib.GridapB200Error) as e:
        lib.DeviceMesh(ctx, model.node_coordinates, model.cell_node_ids, lib.QUAD4)
    assert e.value.code == lib.ERR_UNSUPPORTED
    # node id out of range
    bad = model.cell_node_ids.copy()
    bad[3, 2] = 999
    with pytest.raises(lib.GridapB200Error) as e:
        lib.DeviceMesh(ctx, model.node_coordinates, bad, lib.HEX8)
    assert e.value.code == lib.ERR_INVALID and "node ids" in str(e.value)
    mesh = lib.DeviceMesh(ctx, model.node_coordinates, model.cell_node_ids, lib.HEX8)
    w, N, dN = _tab()
    refel = lib.DeviceRefEl(ctx, w, N, dN, 1)
    # ragged cell_dof table (a space with a varying number of DoFs per cell) -> unsupported
    data = np.ones(8 * 8, dtype=np.int32)
    ptrs = (1 + 8 * np.arange(9)).astype(np.int32)
    ptrs[4] += 1
    rc, _ = _raw_space(ctx, mesh, refel, data, ptrs, 1, 0)
    assert rc == lib.ERR_UNSUPPORTED and b"DoFs" in lib.load().gb200_last_error(ctx.h)
    # DoF id beyond nfree -> invalid
    ptrs = (1 + 8 * np.arange(9)).astype(np.int32)
    data[5] = 7
    rc, _ = _raw_space(ctx, mesh, refel, data, ptrs, 3, 0)
    assert rc == lib.ERR_INVALID
    # unsupported integrand id / form-space mismatch -> unsupported, never a fallback
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1))
    dO = g.Measure(g.Triangulation(model), 2)
    plan = g.SparseMatrixAssembler(V, V).plan(dO)
    for form in (99, lib.FORM_ELASTICITY, lib.FORM_STOKES, lib.FORM_NEOHOOKEAN_JAC):
        with pytest.raises(NotImplementedError):
            plan.assemble_matrix(form, (1.0, 1.0), np.zeros(plan.nnz))
    with pytest.raises(NotImplementedError):
        plan.assemble_vector(77, (), None, np.zeros(plan.nrows))
    # null handle
    assert lib.load().gb200_plan_nnz(None, None) == lib.ERR_INVALID


def test_empty_triangulation_gives_empty_matrix():
    # Triangulation(model, Int[]) in test/FESpacesTests/SparseMatrixAssemblersTests.jl:36-37: zero cells
    ctx = lib.default_context(0)
    X = np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1], [1, 1, 1]])
    mesh = lib.DeviceMesh(ctx, X, np.zeros((0, 8), dtype=np.int32), lib.HEX8)
    w, N, dN = _tab()
    refel = lib.DeviceRefEl(ctx, w, N, dN, 1)
    space = lib.DeviceSpace(ctx, mesh, refel, np.zeros((0, 8), dtype=np.int32), 8, 0)
    plan = lib.DevicePlan(ctx, mesh, refel, [space], [space], None, [0], [0], 8, 8)
    assert plan.nnz == 0
    colptr, rowval = plan.pattern()
    assert colptr.tolist() == [1] * 9 and len(rowval) == 0
    b = np.ones(8)
    plan.assemble_vector(lib.FORM_SOURCE, (1.0,), None, b)
    assert (b == 0).all()
    plan.assemble_matrix(lib.FORM_LAPLACIAN, (), None)


def test_all_dirichlet_and_single_cell():
    # every DoF Dirichlet: nfree = 0 is rejected (no rows), single free DoF works
    model = g.CartesianDiscreteModel((0, 1) * 3, (2, 2, 2))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags="boundary")
    assert V.nfree == 1
    dO = g.Measure(g.Triangulation(model), 2)
    A = g.assemble_matrix(lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dO, V, V)
    assert A.shape == (1, 1) and A.nnz() == 1
    assert abs(A.nzval[0] - 8 * (0.5 / 3)) < 1e-14  # 8 cells x h/3
    model1 = g.CartesianDiscreteModel((0, 1) * 3, (1, 1, 1))
    V1 = g.TestFESpace(model1, g.ReferenceFE(g.lagrangian, float, 1))
    M = g.assemble_matrix(lambda u, v: g.Integral(u * v) * g.Measure(g.Triangulation(model1), 2), V1, V1)
    assert M.nnz() == 64 and abs(M.nzval.sum() - 1.0) < 1e-14  # sum of the mass matrix = volume


def test_trial_and_test_spaces_with_different_dirichlet_sets():
    # rows from the test space, columns from the trial space (AssemblyStrategy semantics, src/FESpaces/Assemblers.jl:31-55)
    from oracle import capi, problems
    part = (4, 3, 3)
    model = g.CartesianDiscreteModel((0, 1) * 3, part)
    reffe = g.ReferenceFE(g.lagrangian, float, 1)
    V = g.TestFESpace(model, reffe, dirichlet_tags=[21])
    U = g.TestFESpace(model, reffe, dirichlet_tags=[22, 25])
    dO = g.Measure(g.Triangulation(model), 2)
    A = g.assemble_matrix(lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dO, U, V)
    assert A.shape == (V.nfree, U.nfree)
    # reference loop with two id tables
    pbv = problems.single_field_problem((0, 1) * 3, part, dirichlet_tags=[21])
    b = capi.Builder(V.nfree, U.nfree)
    for c in range(model.num_cells()):
        for j in U.cell_dof_ids[c]:
            for i in V.cell_dof_ids[c]:
                b.count(i, j)
    b.allocate()
    for c in range(model.num_cells()):
        Ke = pbv.cell_local(c)[0][0][0]
        for lj, j in enumerate(U.cell_dof_ids[c]):
            for li, i in enumerate(V.cell_dof_ids[c]):
                if i > 0 and j > 0:
                    b.add(Ke[li, lj], int(i), int(j))
    colptr, rowval, nzval = b.finish()
    assert np.array_equal(A.colptr, colptr) and np.array_equal(A.rowval, rowval)
    assert np.abs(A.nzval - nzval).max() <= 1e-12 * np.abs(nzval).max()
