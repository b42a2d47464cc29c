"""Named small cases behind the committed fixtures in tests/golden/ (one .npz per case).

The reference is Julia and cannot be executed in the build container (no `julia`), so the fixtures are produced by the CPU
oracle (`tests/golden/make_golden.py`) AFTER the oracle has been pinned against the reference's own golden values
(tests/test_oracle_golden.py; `reference_known_answers.json` holds those values verbatim with their file:line).  The fixtures then
freeze the oracle's output: the CPU suite checks the oracle still reproduces them, the GPU suite checks the CUDA path against
them through the C ABI.  Inputs are deterministic (fixed seeds below)."""
import numpy as np

from oracle import capi, problems

E, NU = 2.1e4, 0.3
LAM, MU = E * NU / ((1 + NU) * (1 - 2 * NU)), E / (2 * (1 + NU))


def _perturbed(partition, seed, amp):
    X = problems.rn.cartesian_node_coordinates((0, 1) * len(partition), partition)
    rng = np.random.default_rng(seed)
    inner = np.all((X > 1e-9) & (X < 1 - 1e-9), axis=1)
    X[inner] += amp * rng.uniform(-1, 1, size=(int(inner.sum()), X.shape[1]))
    return X


def poisson_2x2_reference():
    # test/FESpacesTests/SparseMatrixAssemblersTests.jl:16-40
    kw = dict(order=1, degree=2, dirichlet_tags=[1, 2, 3, 4, 6, 5], form_mat=capi.LAPLACIAN, form_vec=capi.SOURCE)
    xq = problems.single_field_problem((0, 1, 0, 1), (2, 2), **kw).quadrature_points()
    return problems.single_field_problem((0, 1, 0, 1), (2, 2), fq=xq[:, :, 1].copy(), lift=True, **kw), {}


def poisson_q1_2d():
    # config 1 at 12x9: Laplacian + source f = 1, homogeneous Dirichlet boundary
    return problems.single_field_problem((0, 1, 0, 1), (12, 9), form_mat=capi.LAPLACIAN, form_vec=capi.SOURCE, params=[1.0]), {}


def poisson_q1_3d_lifting():
    # config 2 at 6x5x4: f at the quadrature points (seed 11), Dirichlet values sin(k), lifting
    part = (6, 5, 4)
    pb0 = problems.single_field_problem((0, 1) * 3, part, form_mat=capi.LAPLACIAN, form_vec=capi.SOURCE)
    dv = np.sin(np.arange(pb0.ndiri) + 1.0)
    fq = np.random.default_rng(11).uniform(-1, 1, size=(len(pb0.cells), 8))
    pb = problems.single_field_problem((0, 1) * 3, part, form_mat=capi.LAPLACIAN, form_vec=capi.SOURCE, fq=fq, dirichlet_values=dv, lift=True)
    return pb, {"dirichlet_values": dv}


def poisson_q1_3d_perturbed():
    # config 2, general-geometry variant at 5x5x4, seed 12345, amplitude 0.2*dx
    part = (5, 5, 4)
    return problems.single_field_problem((0, 1) * 3, part, form_mat=capi.LAPLACIAN, X=_perturbed(part, 12345, 0.2 * 0.2)), {}


def mass_q1_3d():
    return problems.single_field_problem((0, 1) * 3, (4, 4, 5), form_mat=capi.MASS), {}


def elasticity_q2():
    # config 3 at 2x2x3: Dirichlet on the face x = 0
    tags = [25, 1, 3, 5, 7, 13, 15, 17, 19]
    return problems.single_field_problem((0, 1) * 3, (2, 2, 3), order=2, ncomp=3, degree=4, dirichlet_tags=tags, form_mat=capi.ELASTICITY,
                                         params=[LAM, MU]), {}


def stokes_taylor_hood():
    # config 4 at 2x2x2 hexes -> 48 tets
    return problems.stokes_problem((0, 1) * 3, (2, 2, 2), degree=4, simplex=True), {}


def neohookean_q1():
    # config 5 at 3x3x3: u = 0.05 sin(pi x) sin(pi y) sin(pi z) (1,1,1) at the free nodes, lambda = 100, mu = 1
    part = (3, 3, 3)
    pb0 = problems.single_field_problem((0, 1) * 3, part, ncomp=3, form_mat=capi.NEOHOOKEAN_JAC, form_vec=capi.NEOHOOKEAN_RES, params=[100.0, 1.0])
    X = pb0.X
    s = 0.05 * np.sin(np.pi * X[:, 0]) * np.sin(np.pi * X[:, 1]) * np.sin(np.pi * X[:, 2])
    fv = np.zeros(pb0.nfree)
    for c, nodes in enumerate(pb0.cells):
        for comp in range(3):
            for a in range(8):
                d = pb0.cell_dofs[c][comp * 8 + a]
                if d > 0:
                    fv[d - 1] = s[nodes[a] - 1]
    dv = np.zeros(pb0.ndiri)
    pb = problems.single_field_problem((0, 1) * 3, part, ncomp=3, form_mat=capi.NEOHOOKEAN_JAC, form_vec=capi.NEOHOOKEAN_RES, params=[100.0, 1.0],
                                       free_values=fv, dirichlet_values=dv)
    return pb, {"free_values": fv, "dirichlet_values": dv}


def poisson_q3_2d_lifting():
    # order 3 (the reference's benchmark sweeps orders 1-3): Q3 on 3x2 quadrilaterals, f = 1, Dirichlet values cos(k), lifting
    kw = dict(order=3, degree=6, dirichlet_tags="boundary", form_mat=capi.LAPLACIAN, form_vec=capi.SOURCE, params=[1.0])
    pb0 = problems.single_field_problem((0, 1, 0, 1), (3, 2), **kw)
    dv = np.cos(np.arange(pb0.ndiri) + 1.0)
    return problems.single_field_problem((0, 1, 0, 1), (3, 2), dirichlet_values=dv, lift=True, **kw), {"dirichlet_values": dv}


def mass_p3_tet():
    # P3 on the 12 tetrahedra of 2x1x1 hexahedra, perturbed nodes are not needed: the mass matrix of a cubic space, degree 5
    return problems.single_field_problem((0, 1) * 3, (2, 1, 1), order=3, degree=5, dirichlet_tags=[25], form_mat=capi.MASS, simplex=True), {}


CASES = {
    "poisson_2x2_reference": (poisson_2x2_reference, True),
    "poisson_q1_2d": (poisson_q1_2d, True),
    "poisson_q1_3d_lifting": (poisson_q1_3d_lifting, True),
    "poisson_q1_3d_perturbed": (poisson_q1_3d_perturbed, False),
    "mass_q1_3d": (mass_q1_3d, False),
    "elasticity_q2": (elasticity_q2, False),
    "stokes_taylor_hood": (stokes_taylor_hood, False),
    "neohookean_q1": (neohookean_q1, True),
    "poisson_q3_2d_lifting": (poisson_q3_2d_lifting, True),
    "mass_p3_tet": (mass_p3_tet, False),
}


def build(name):
    """-> (oracle Problem, extra state dict, with_vector)"""
    fn, with_vector = CASES[name]
    pb, extra = fn()
    return pb, extra, with_vector
