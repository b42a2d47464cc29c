"""The forms whose exact matrices the reference's tests do not pin (linear elasticity, Stokes, neo-Hookean) are pinned
in the oracle by construction: patch tests, symmetry, rigid-body null space, finite-difference Jacobian."""
import numpy as np

from oracle import capi, problems


def dense(pb):
    colptr, rowval, nzval = pb.assemble()
    return problems.csc_to_dense(colptr, rowval, nzval, pb.nrows, pb.ncols)


def test_elasticity_symmetric_and_rigid_body_null_space():
    lam, mu = 3.0, 2.0
    pb = problems.single_field_problem((0, 1) * 3, (2, 2, 2), order=1, ncomp=3, dirichlet_tags=[], form_mat=capi.ELASTICITY, params=[lam, mu])
    A = dense(pb)
    assert np.allclose(A, A.T, atol=1e-13)
    X = pb.X
    # dofs are node-major, component-minor for CLagrangian: dof = 3*node + comp
    trans = np.tile([1.0, 0.0, 0.0], len(X))
    rot = np.stack([-X[:, 1], X[:, 0], np.zeros(len(X))], axis=1).ravel()
    assert np.abs(A @ trans).max() < 1e-12 and np.abs(A @ rot).max() < 1e-12
    assert np.linalg.eigvalsh(A).min() > -1e-10
    # uniaxial strain u = (x,0,0): energy u^T A u = (lambda + 2 mu) * volume
    u = np.stack([X[:, 0], np.zeros(len(X)), np.zeros(len(X))], axis=1).ravel()
    assert abs(u @ A @ u - (lam + 2 * mu)) < 1e-12


def test_stokes_blocks():
    pb = problems.stokes_problem((0, 1) * 3, (2, 1, 1), degree=4, simplex=True)
    A = dense(pb)
    nfu, nfp = pb.nfree
    assert np.allclose(A[:nfu, :nfu], A[:nfu, :nfu].T, atol=1e-13)
    assert np.allclose(A[:nfu, nfu:], -A[nfu:, :nfu].T, atol=1e-13)  # -(div v) p  vs  q (div u)
    assert np.abs(A[nfu:, nfu:]).max() == 0.0


def test_neohookean_jacobian_is_derivative_of_residual():
    lam, mu = 100.0, 1.0
    n = 2
    rng = np.random.default_rng(0)
    base = problems.single_field_problem((0, 1) * 3, (n, n, n), ncomp=3, dirichlet_tags=[21], form_mat=capi.NEOHOOKEAN_JAC,
                                         form_vec=capi.NEOHOOKEAN_RES, params=[lam, mu])
    nfree = base.nfree
    u0 = 0.02 * rng.standard_normal(nfree)

    def res(u):
        pb = problems.single_field_problem((0, 1) * 3, (n, n, n), ncomp=3, dirichlet_tags=[21], form_mat=0, form_vec=capi.NEOHOOKEAN_RES,
                                           params=[lam, mu], free_values=u)
        return pb.assemble_vector()

    pbj = problems.single_field_problem((0, 1) * 3, (n, n, n), ncomp=3, dirichlet_tags=[21], form_mat=capi.NEOHOOKEAN_JAC,
                                        params=[lam, mu], free_values=u0)
    J = dense(pbj)
    assert np.allclose(J, J.T, atol=1e-10)
    h = 1e-6
    for k in rng.choice(nfree, size=6, replace=False):
        e = np.zeros(nfree)
        e[k] = h
        fd = (res(u0 + e) - res(u0 - e)) / (2 * h)
        assert np.abs(fd - J[:, k]).max() < 1e-6 * max(1.0, np.abs(J[:, k]).max())
    # zero displacement: zero residual, Jacobian = linear elasticity with the same Lame parameters
    assert np.abs(res(np.zeros(nfree))).max() < 1e-14
    pbe = problems.single_field_problem((0, 1) * 3, (n, n, n), ncomp=3, dirichlet_tags=[21], form_mat=capi.ELASTICITY, params=[lam, mu])
    pb0 = problems.single_field_problem((0, 1) * 3, (n, n, n), ncomp=3, dirichlet_tags=[21], form_mat=capi.NEOHOOKEAN_JAC, params=[lam, mu],
                                        free_values=np.zeros(nfree))
    assert np.allclose(dense(pb0), dense(pbe), atol=1e-11)
