"""BASELINE.json configs[1] at FULL size (256^3 cells, 16.6 M DoFs, 444 M nnz) through size-independent properties:
structure counts, symmetry, the discrete Laplacian annihilating linear functions on interior rows, linearity in the
coefficient, bitwise reproducibility, and agreement with the generic (atomic) kernel path."""
import numpy as np
import pytest

import gridap_b200 as g
from gridap_b200 import lib

pytestmark = pytest.mark.gpu

N = 256


@pytest.fixture(scope="module")
def problem():
    model = g.UnstructuredDiscreteModel(g.CartesianDiscreteModel((0, 1) * 3, (N, N, N)))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags="boundary")
    U = g.TrialFESpace(V, 0.0)
    dO = g.Measure(g.Triangulation(model), 2)
    assem = g.SparseMatrixAssembler(U, V)
    A = g.assemble_matrix(lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dO, assem, U, V)
    return model, V, U, dO, assem, A


def test_structure_counts(problem):
    model, V, U, dO, assem, A = problem
    assert V.num_free_dofs() == (N - 1) ** 3
    assert A.shape == ((N - 1) ** 3, (N - 1) ** 3)
    assert A.nnz() == (3 * (N - 1) - 2) ** 3  # 763^3 = 444 194 947
    assert A.colptr[0] == 1 and A.colptr[-1] == A.nnz() + 1
    counts = np.diff(A.colptr)
    assert counts.min() == 8 and counts.max() == 27
    # rows ascending and unique within each column, checked on a sample of columns and globally through the differences
    d = np.diff(A.rowval)
    starts = A.colptr[1:-1] - 1
    mask = np.ones(len(d), dtype=bool)
    mask[starts - 1] = False
    assert (d[mask] > 0).all()
    assert assem.plan(dO).kernel_path(lib.FORM_LAPLACIAN) == "q1hex_gather_affine+diag"


def test_symmetry_and_linear_functions(problem):
    model, V, U, dO, assem, A = problem
    S = A.to_scipy()
    rng = np.random.default_rng(7)
    x, y = rng.standard_normal(S.shape[0]), rng.standard_normal(S.shape[0])
    Ax, Ay = S @ x, S @ y
    assert abs(y @ Ax - x @ Ay) <= 1e-12 * abs(y @ Ax)
    assert (x @ Ax) > 0  # SPD
    # A applied to a linear function vanishes on rows whose 27-point stencil holds no Dirichlet node
    fx, _, _, _ = V.dof_coordinates()
    lin = 1.0 + 2.0 * fx[:, 0] - 3.0 * fx[:, 1] + 0.5 * fx[:, 2]
    r = S @ lin
    h = 1.0 / N
    interior = np.all((fx > 1.5 * h) & (fx < 1 - 1.5 * h), axis=1)
    assert interior.sum() == (N - 3) ** 3
    assert np.abs(r[interior]).max() <= 1e-12 * np.abs(A.nzval).max() * np.abs(lin).max() * 27
    # diagonal of a uniform-mesh Q1 Laplacian: 8 cells x (h/3) = 8h/3
    diag = S.diagonal()
    assert np.allclose(diag, 8.0 * h / 3.0, rtol=1e-13, atol=0)


def test_linearity_reproducibility_and_generic_path(problem):
    model, V, U, dO, assem, A = problem
    plan = assem.plan(dO)
    nz2 = np.zeros(plan.nnz)
    plan.assemble_matrix(lib.FORM_LAPLACIAN, (2.0,), nz2)
    assert np.array_equal(nz2, 2.0 * A.nzval)  # scaling by 2 is exact in binary floating point
    nz3 = np.zeros(plan.nnz)
    plan.assemble_matrix(lib.FORM_LAPLACIAN, (), nz3)
    assert np.array_equal(nz3, A.nzval)  # owner-computes gather: bitwise reproducible
    # _add! accumulates on top of the caller's values
    plan.assemble_matrix(lib.FORM_LAPLACIAN, (), nz3, add=True)
    assert np.array_equal(nz3, 2.0 * A.nzval)
    # checksum of checksums against the generic cell-centric kernel (atomics): per-column sums agree to round-off
    Ke_path = np.zeros(plan.nnz)
    xq, w = dO.points, dO.weights
    # same matrix through gb200_assemble_matrix_const with the (constant) local matrix of this uniform mesh
    h = 1.0 / N
    from gridap_b200 import reffes as rf
    Nq, dNq = rf.tabulate_lagrangian("HEX", 1, xq)
    Ke = np.einsum("p,pad,pbd->ab", w, dNq, dNq) * h  # |det| h^3, inv(J)^2 = h^-2
    plan.assemble_matrix_const(Ke, Ke_path)
    assert np.abs(Ke_path - A.nzval).max() <= 1e-12 * np.abs(A.nzval).max()
