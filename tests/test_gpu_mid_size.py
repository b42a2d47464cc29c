"""BASELINE.json configs 3-5 at sizes far beyond what the oracle can check entry by entry (10^8 .. 10^9 stored entries), through
size-independent properties evaluated ON THE DEVICE (the plan's CSC bound zero-copy into a torch sparse tensor):
rigid-body modes in the null space of the unconstrained elasticity operator, symmetry of the hyperelastic tangent and its
agreement with a finite difference of the residual along a direction, the saddle-point structure of the Stokes matrix."""
import numpy as np
import pytest

import gridap_b200 as g
from gridap_b200 import lib

pytestmark = pytest.mark.gpu


def device_csc(plan):
    import torch
    cp, rv, nz, bv = (torch.as_tensor(x, device="cuda") for x in plan.device_arrays())
    return torch.sparse_csc_tensor(cp, rv.to(torch.int64), nz, size=(plan.nrows, plan.ncols)), nz, bv


def test_config3_q2_elasticity_rigid_body_modes():
    import torch
    n = 36   # 46 656 Q2 hexahedra, 1.2 M DoFs, 1.9e8 stored entries; no Dirichlet boundary: K has the 6 rigid-body modes in its kernel
    model = g.CartesianDiscreteModel((0, 1) * 3, (n, n, n))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 2))
    dO = g.Measure(g.Triangulation(model), 4)
    assem = g.SparseMatrixAssembler(V, V)
    plan = assem.plan(dO)
    E, nu = 2.1e4, 0.3
    lam, mu = E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))
    plan.assemble_matrix(lib.FORM_ELASTICITY, (lam, mu), None)
    assert plan.kernel_path(lib.FORM_ELASTICITY).startswith("affine_gather")   # Cartesian mesh: affine cells, owner-computes column-node gather
    assem.ctx.synchronize()
    A, nz, _ = device_csc(plan)
    assert plan.nnz > 1.8e8 and bool(torch.isfinite(nz).all())
    fx, fc, _, _ = V.dof_coordinates()
    scale = float(nz.abs().max())
    for mode in range(6):
        if mode < 3:   # translation e_mode
            u = (fc == mode).astype(np.float64)
        else:          # rotation about axis k: u = e_k x (x - c)
            k = mode - 3
            i, j = (k + 1) % 3, (k + 2) % 3
            u = np.where(fc == i, -(fx[:, j] - 0.5), np.where(fc == j, fx[:, i] - 0.5, 0.0))
        r = torch.mv(A, torch.as_tensor(u, device="cuda"))
        assert float(r.abs().max()) <= 1e-11 * scale * 81, mode
    rng = np.random.default_rng(1)
    x, y = (torch.as_tensor(rng.standard_normal(plan.nrows), device="cuda") for _ in range(2))
    Ax, Ay = torch.mv(A, x), torch.mv(A, y)
    a, b = float(torch.dot(y, Ax)), float(torch.dot(x, Ay))
    assert abs(a - b) <= 1e-12 * float(y.norm() * Ax.norm()) and float(torch.dot(x, Ax)) > 0   # symmetric, positive semi-definite


def test_config5_neohookean_tangent_is_symmetric_and_consistent():
    import torch
    n = 72   # 373 248 cells, 1.1 M DoFs, 8.4e7 stored entries
    model = g.CartesianDiscreteModel((0, 1) * 3, (n, n, n))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 1), dirichlet_tags="boundary")
    U = g.TrialFESpace(V, (0.0, 0.0, 0.0))
    dO = g.Measure(g.Triangulation(model), 2)
    ufun = lambda x: 0.05 * (np.sin(np.pi * x[:, 0]) * np.sin(np.pi * x[:, 1]) * np.sin(np.pi * x[:, 2]))[:, None] * np.ones((1, 3))  # noqa: E731
    uh = g.interpolate(ufun, U)
    assem = g.SparseMatrixAssembler(U, V)
    plan = assem.plan(dO)
    prm = (100.0, 1.0)
    plan.set_state(0, uh.free_values, uh.dirichlet_values)
    plan.assemble_matrix_and_vector(lib.FORM_NEOHOOKEAN_JAC, prm, lib.FORM_NEOHOOKEAN_RES, prm, None, None, None)   # fused, device-resident
    assem.ctx.synchronize()
    A, nz, bv = device_csc(plan)
    r0 = bv.clone()
    rng = np.random.default_rng(2)
    x, y = (torch.as_tensor(rng.standard_normal(plan.nrows), device="cuda") for _ in range(2))
    Ax, Ay = torch.mv(A, x), torch.mv(A, y)
    a, b = float(torch.dot(y, Ax)), float(torch.dot(x, Ay))
    assert abs(a - b) <= 1e-11 * float(y.norm() * Ax.norm())             # hyperelastic tangent: symmetric
    # J(u) d = (r(u + eps d) - r(u - eps d)) / (2 eps) + O(eps^2) for a smooth direction d
    d = g.interpolate(lambda x: 0.3 * np.stack([np.sin(2 * x[:, 1]) * x[:, 0], np.cos(x[:, 2]), x[:, 0] * x[:, 1]], axis=1) *
                      (x[:, 0] * (1 - x[:, 0]) * x[:, 1] * (1 - x[:, 1]) * x[:, 2] * (1 - x[:, 2]))[:, None], U).free_values
    Jd = torch.mv(A, torch.as_tensor(d, device="cuda")).clone()
    torch.cuda.synchronize()   # torch reads the plan's arrays on its own stream: finish before the library overwrites them
    eps = 1e-4   # truncation ~ eps^2, round-off of the atomically summed residuals ~ 1e-16 / eps: both far below the tolerance
    rs = []
    for sgn in (+1.0, -1.0):
        plan.set_state(0, uh.free_values + sgn * eps * d, uh.dirichlet_values)
        plan.assemble_vector(lib.FORM_NEOHOOKEAN_RES, prm, None, None)
        assem.ctx.synchronize()
        rs.append(torch.as_tensor(plan.device_arrays()[3], device="cuda").clone())
        torch.cuda.synchronize()
    fd = (rs[0] - rs[1]) / (2 * eps)
    assert float((fd - Jd).abs().max()) <= 1e-4 * float(Jd.abs().max())   # (a wrong tangent gives an O(1) relative difference)
    assert float((r0).abs().max()) > 0


def test_config4_stokes_saddle_point_structure():
    import torch
    n = 20   # 48 000 P2/P1 tetrahedra, 1.9e7 stored entries
    model = g.simplexify(g.CartesianDiscreteModel((0, 1) * 3, (n, n, n)))
    Vv = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 2), dirichlet_tags="boundary")
    Q = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1))
    Y = g.MultiFieldFESpace([Vv, Q])
    dO = g.Measure(g.Triangulation(model), 4)
    assem = g.SparseMatrixAssembler(Y, Y)
    plan = assem.plan(dO, np.array([[1, 1], [1, 0]], dtype=np.uint8))
    plan.assemble_matrix(lib.FORM_STOKES, (), None)
    assem.ctx.synchronize()
    A, nz, _ = device_csc(plan)
    nu, npr = Vv.num_free_dofs(), Q.num_free_dofs()
    rng = np.random.default_rng(3)
    xu, yu = np.zeros(plan.nrows), np.zeros(plan.nrows)
    xp, yp = np.zeros(plan.nrows), np.zeros(plan.nrows)
    xu[:nu], yu[:nu] = rng.standard_normal(nu), rng.standard_normal(nu)
    xp[nu:], yp[nu:] = rng.standard_normal(npr), rng.standard_normal(npr)
    T = lambda v: torch.as_tensor(v, device="cuda")  # noqa: E731
    dot = lambda a, b: float(torch.dot(T(a), torch.mv(A, T(b))))  # noqa: E731
    nrm = lambda a, b: float(T(a).norm() * torch.mv(A, T(b)).norm())  # noqa: E731
    assert abs(dot(yu, xu) - dot(xu, yu)) <= 1e-12 * nrm(yu, xu) and dot(xu, xu) > 0         # velocity block: symmetric positive definite
    assert abs(dot(yp, xp)) == 0.0                                                           # (q,p) block: absent
    assert abs(dot(xu, xp) + dot(xp, xu)) <= 1e-12 * nrm(xu, xp)                              # [v,p] = -[q,u]^T
    # constant pressure is in the kernel of the gradient block when the velocity vanishes on the whole boundary: int (div v) 1 = 0
    one = np.zeros(plan.nrows)
    one[nu:] = 1.0
    r = torch.mv(A, T(one))[:nu]
    assert float(r.abs().max()) <= 1e-12 * float(nz.abs().max()) * 30


def test_newton_loop_entirely_on_the_device():
    # N2 + N1 of SURVEY 8(f): persistent plan, residual_and_jacobian re-assembled every Newton iteration, unknown and linear solve on the
    # GPU (gb200_plan_set_state_device, CG on the plan's device CSC).  A homogeneous neo-Hookean block with the affine boundary
    # displacement u_D = 0.01 x e_x has the affine field as its exact equilibrium.
    import torch
    n = 10
    model = g.CartesianDiscreteModel((0, 1) * 3, (n, n, n))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 1), dirichlet_tags="boundary")
    aff = lambda x: np.stack([0.01 * x[:, 0], np.zeros(len(x)), np.zeros(len(x))], axis=1)  # noqa: E731
    U = g.TrialFESpace(V, aff)
    dO = g.Measure(g.Triangulation(model), 2)
    assem = g.SparseMatrixAssembler(U, V)
    plan = assem.plan(dO)
    prm = (100.0, 1.0)
    u = torch.zeros(plan.nrows, dtype=torch.float64, device="cuda")
    plan.set_state(0, np.zeros(plan.nrows), U.dirichlet_values)
    norms = []
    for it in range(10):
        torch.cuda.synchronize()
        plan.set_state_device(0, u)
        plan.assemble_matrix_and_vector(lib.FORM_NEOHOOKEAN_JAC, prm, lib.FORM_NEOHOOKEAN_RES, prm, None, None, None)
        assem.ctx.synchronize()
        A, nz, bv = device_csc(plan)
        r = bv.clone()
        norms.append(float(r.norm()))
        if norms[-1] <= 1e-12 * norms[0]:
            break
        du = torch.zeros_like(u)   # CG for J du = -r
        res = -r
        p = res.clone()
        rs = torch.dot(res, res)
        for _ in range(600):
            Ap = torch.mv(A, p)
            alpha = rs / torch.dot(p, Ap)
            du += alpha * p
            res -= alpha * Ap
            rs_new = torch.dot(res, res)
            if float(rs_new) <= 1e-28 * norms[-1] ** 2 + 1e-300:
                break
            p = res + (rs_new / rs) * p
            rs = rs_new
        u += du
    assert norms[-1] <= 1e-9 * norms[0], norms
    assert norms[-1] < 1e-3 * norms[-2] or norms[-1] <= 1e-12 * norms[0], norms   # Newton: fast contraction once close
    expect = g.interpolate(aff, U).free_values
    assert np.abs(u.cpu().numpy() - expect).max() <= 1e-9 * 0.01 * 100
