"""GPU parity of the C ABI against the CPU oracle on oracle-generated inputs (no host mirror involved)."""
import numpy as np
import pytest

import gridap_b200  # noqa: F401
from gridap_b200 import lib
from oracle import capi, problems

pytestmark = pytest.mark.gpu

CELLTYPE = {"QUAD": lib.QUAD4, "HEX": lib.HEX8, "TRI": lib.TRI3, "TET": lib.TET4}


def device_problem(pb, deterministic=False):
    ctx = lib.default_context(0, deterministic)
    xq, w, N, dN, Ng, dNg = pb.tab[:6]
    mesh = lib.DeviceMesh(ctx, pb.X, pb.cell_nodes, CELLTYPE[pb.ptype])
    geo = lib.DeviceRefEl(ctx, w, Ng, dNg, 1)
    spaces = []
    for f in pb.fields:
        refel = lib.DeviceRefEl(ctx, w, f.N, f.dN, f.ncomp)
        ids = f.cell_dofs.copy()
        ids[ids > 0] -= f.offset
        nfree = int(ids.max(initial=0))
        ndir = int(-ids.min(initial=0))
        spaces.append(lib.DeviceSpace(ctx, mesh, refel, ids, nfree, ndir))
    offs = [f.offset for f in pb.fields]
    plan = lib.DevicePlan(ctx, mesh, geo, spaces, spaces, pb.touched, offs, offs, pb.nrows, pb.ncols)
    return ctx, plan


def rel_err(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("deterministic", [False, True])
@pytest.mark.parametrize("partition", [(5, 4), (4, 3, 5)])
def test_poisson_q1_matrix_and_pattern(partition, deterministic):
    D = len(partition)
    domain = (0, 1) * D
    pb = problems.single_field_problem(domain, partition, form_mat=capi.LAPLACIAN)
    colptr, rowval, nzval = pb.assemble()
    ctx, plan = device_problem(pb, deterministic)
    cp, rv = plan.pattern()
    assert np.array_equal(cp, colptr) and np.array_equal(rv, rowval)  # bit-exact pattern
    nz = np.zeros(plan.nnz)
    plan.assemble_matrix(lib.FORM_LAPLACIAN, (), nz)
    assert rel_err(nz, nzval) <= 1e-12
    print(plan.kernel_path(lib.FORM_LAPLACIAN))


def test_poisson_q1_perturbed_mesh_general_geometry_path():
    partition = (4, 4, 4)
    X = problems.rn.cartesian_node_coordinates((0, 1) * 3, partition)
    rng = np.random.default_rng(12345)
    inner = np.all((X > 1e-9) & (X < 1 - 1e-9), axis=1)
    X[inner] += 0.2 * 0.25 * rng.uniform(-1, 1, size=(inner.sum(), 3))
    pb = problems.single_field_problem((0, 1) * 3, partition, form_mat=capi.LAPLACIAN, X=X)
    colptr, rowval, nzval = pb.assemble()
    ctx, plan = device_problem(pb)
    nz = np.zeros(plan.nnz)
    plan.assemble_matrix(lib.FORM_LAPLACIAN, (), nz)
    assert plan.kernel_path(lib.FORM_LAPLACIAN) == "q1hex_gather_general"  # non-affine cells: staged local matrices + gather
    assert rel_err(nz, nzval) <= 1e-12


def test_matrix_and_vector_with_lifting():
    partition = (4, 3, 3)
    pb0 = problems.single_field_problem((0, 1) * 3, partition, form_mat=capi.LAPLACIAN, form_vec=capi.SOURCE, params=[1.0])
    dv = np.sin(np.arange(pb0.ndiri) + 1.0)
    pb = problems.single_field_problem((0, 1) * 3, partition, form_mat=capi.LAPLACIAN, form_vec=capi.SOURCE, params=[1.0],
                                       dirichlet_values=dv, lift=True)
    colptr, rowval, nzval, b = pb.assemble(with_vector=True)
    ctx, plan = device_problem(pb)
    plan.set_state(0, None, dv)
    nz, bb = np.zeros(plan.nnz), np.zeros(plan.nrows)
    plan.assemble_matrix_and_vector(lib.FORM_LAPLACIAN, (), lib.FORM_SOURCE, (1.0,), None, nz, bb)
    assert rel_err(nz, nzval) <= 1e-12 and rel_err(bb, b) <= 1e-12


@pytest.mark.parametrize("form", ["laplacian", "mass"])
def test_vector_with_lifting_on_perturbed_mesh(form):
    # b_e - K_e u_e on non-affine cells, f given at the quadrature points (src/CellData/AttachDirichlet.jl:76-84)
    partition = (5, 4, 3)
    X = problems.rn.cartesian_node_coordinates((0, 1) * 3, partition)
    rng = np.random.default_rng(7)
    inner = np.all((X > 1e-9) & (X < 1 - 1e-9), axis=1)
    X[inner] += 0.04 * rng.uniform(-1, 1, size=(inner.sum(), 3))
    fm, lf = (capi.LAPLACIAN, lib.FORM_LAPLACIAN) if form == "laplacian" else (capi.MASS, lib.FORM_MASS)
    pb0 = problems.single_field_problem((0, 1) * 3, partition, form_mat=fm, form_vec=capi.SOURCE, X=X)
    dv = np.cos(0.3 * np.arange(pb0.ndiri)) + 2.0
    fq = rng.uniform(-1, 1, size=(len(pb0.cells), 8))
    pb = problems.single_field_problem((0, 1) * 3, partition, form_mat=fm, form_vec=capi.SOURCE, X=X, fq=fq, dirichlet_values=dv, lift=True)
    colptr, rowval, nzval, b = pb.assemble(with_vector=True)
    ctx, plan = device_problem(pb)
    plan.set_state(0, None, dv)
    nz, bb = np.zeros(plan.nnz), np.zeros(plan.nrows)
    plan.assemble_matrix_and_vector(lf, (), lib.FORM_SOURCE, (0.0,), fq, nz, bb)
    assert rel_err(nz, nzval) <= 1e-12 and rel_err(bb, b) <= 1e-12
    # vector alone (no lifting): assemble_vector
    pbv = problems.single_field_problem((0, 1) * 3, partition, form_mat=fm, form_vec=capi.SOURCE, X=X, fq=fq)
    bv = pbv.assemble(with_vector=True)[3]
    b2 = np.zeros(plan.nrows)
    plan.assemble_vector(lib.FORM_SOURCE, (0.0,), fq, b2)
    assert rel_err(b2, bv) <= 1e-12
