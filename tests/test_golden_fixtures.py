"""The committed fixtures under tests/golden/: (CPU) the oracle reproduces them and the reference's known answers;
(GPU) the CUDA path reproduces them through the C ABI.  Pattern bit-exact, values within 1e-12 relative (north star)."""
import json
import os

import numpy as np
import pytest

import golden_cases
from oracle import capi
from oracle import ref_numbering as rn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KNOWN = json.load(open(os.path.join(GOLDEN, "reference_known_answers.json")))


def rel_err(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.mark.parametrize("name", list(golden_cases.CASES))
def test_oracle_reproduces_fixture(name):
    pb, extra, with_vector = golden_cases.build(name)
    out = pb.assemble(with_vector=with_vector)
    gold = load(name)
    assert np.array_equal(out[0], gold["colptr"]) and np.array_equal(out[1], gold["rowval"])
    assert rel_err(out[2], gold["nzval"]) <= 1e-14
    if with_vector:
        assert rel_err(out[3], gold["b"]) <= 1e-14


def test_fixture_holds_the_reference_values():
    # test/FESpacesTests/SparseMatrixAssemblersTests.jl:104-152
    k = KNOWN["sparse_matrix_assembler_2x2"]
    gold = load("poisson_2x2_reference")
    assert len(gold["colptr"]) == k["nfree"] + 1
    assert np.allclose(gold["b"], k["vec"], rtol=0, atol=1e-14)
    for i, j, v in k["mat_entries"]:
        col = slice(gold["colptr"][j - 1] - 1, gold["colptr"][j] - 1)
        pos = list(gold["rowval"][col]).index(i)
        assert abs(gold["nzval"][col][pos] - v) < 1e-13


def test_known_answers_csc_builder_and_numbering():
    k = KNOWN["csc_builder"]
    a = capi.Builder(*k["shape"])
    for i, j in k["counted"]:
        a.count(i, j)
    assert a.colnnzmax.tolist() == k["colnnzmax"]
    a.allocate()
    assert a.state()[0].tolist() == k["colptr_after_allocation"]
    for v, i, j in k["added"]:
        a.add(v, i, j)
    assert a.state()[1].tolist() == k["colnnz_after_adds"]
    colptr, rowval, nzval = a.finish()
    assert rowval.tolist() == k["findnz"]["I"] and nzval.tolist() == k["findnz"]["V"]
    assert np.repeat(np.arange(1, 10), np.diff(colptr)).tolist() == k["findnz"]["J"]
    c = KNOWN["cartesian_grid_3x4"]
    x = rn.cartesian_node_coordinates(tuple(c["domain"]), tuple(c["partition"]))
    t = rn.cartesian_cell_node_ids(tuple(c["partition"]))
    assert list(x[12]) == c["node_13"] and list(x[3]) == c["node_4"] and list(t[0]) == c["cell_1"] and list(t[10]) == c["cell_11"]
    assert rn.ncube_face_vertices(2, 1) == KNOWN["polytopes"]["quad_edges"] and rn.HEX_TO_TETS == KNOWN["polytopes"]["hex_to_tets"]


@pytest.mark.gpu
@pytest.mark.parametrize("deterministic", [False, True])
@pytest.mark.parametrize("name", list(golden_cases.CASES))
def test_cuda_path_reproduces_fixture(name, deterministic):
    from gridap_b200 import lib
    from test_gpu_lowlevel import device_problem

    pb, extra, with_vector = golden_cases.build(name)
    gold = load(name)
    ctx, plan = device_problem(pb, deterministic)
    cp, rv = plan.pattern()
    assert np.array_equal(cp, gold["colptr"]) and np.array_equal(rv, gold["rowval"])
    if "free_values" in extra or "dirichlet_values" in extra:
        plan.set_state(0, extra.get("free_values"), extra.get("dirichlet_values"))
    form_mat = pb.pb.form_mat
    form_vec = pb.pb.form_vec
    mat_params = list(pb.params[:2]) if form_mat in (capi.ELASTICITY, capi.NEOHOOKEAN_JAC) else []
    nz = np.zeros(plan.nnz)
    if with_vector and form_vec == capi.SOURCE:
        b = np.zeros(plan.nrows)
        if pb.pb.lift_dirichlet:
            plan.assemble_matrix_and_vector(form_mat, mat_params, form_vec, [pb.params[0]], pb.fq, nz, b)
        else:
            plan.assemble_matrix(form_mat, mat_params, nz)
            plan.assemble_vector(form_vec, [pb.params[0]], pb.fq, b)
        assert rel_err(b, gold["b"]) <= 1e-12
    elif with_vector:  # neo-Hookean residual + Jacobian
        b = np.zeros(plan.nrows)
        plan.assemble_matrix_and_vector(form_mat, mat_params, form_vec, mat_params, None, nz, b)
        assert rel_err(b, gold["b"]) <= 1e-12
    else:
        plan.assemble_matrix(form_mat, mat_params, nz)
    assert rel_err(nz, gold["nzval"]) <= 1e-12
