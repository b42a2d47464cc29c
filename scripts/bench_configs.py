"""Device-resident assembly throughput of the generic element kernels on BASELINE.json configs 3-5 at reduced size
(and config 1).  Prints one JSON line per config.  Usage: python scripts/bench_configs.py [scale]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import gridap_b200 as g  # noqa: E402
from gridap_b200 import lib  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
import os  # noqa: E402
N3, N4, N5, N2B = (int(os.environ.get(k, 0)) for k in ("N3", "N4", "N5", "N2B"))  # explicit sizes override the scale
ONLY = [x for x in os.environ.get("ONLY", "").split(",") if x]  # e.g. ONLY=3,5 runs just those configs
ctx = lib.Context(0)


def want(k):
    return not ONLY or k in ONLY



def timed(plan, call, steps=5, warm=2):
    for _ in range(warm):
        call()
    ctx.synchronize()
    ctx.timings()
    t0 = time.perf_counter()
    for _ in range(steps):
        call()
    ctx.synchronize()
    dt = (time.perf_counter() - t0) / steps
    return dt, ctx.timings()


def report(name, ncells, ndofs, nnz, dt, timers, extra=None):
    out = {"config": name, "ncells": int(ncells), "free_dofs": int(ndofs), "nnz": int(nnz), "ms_per_assembly": dt * 1e3,
           "cells_per_s": ncells / dt, "dofs_per_s": ndofs / dt, "kernels_ms": timers}
    if extra:
        out.update(extra)
    print(json.dumps(out), flush=True)


# config 1: 2D Poisson Q1 100x100
if want("1"):
    model = g.CartesianDiscreteModel((0, 1, 0, 1), (100, 100))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags="boundary")
    dO = g.Measure(g.Triangulation(model), 2)
    assem = g.SparseMatrixAssembler(V, V, ctx=ctx)
    plan = assem.plan(dO)
    dt, tm = timed(plan, lambda: plan.assemble_matrix(lib.FORM_LAPLACIAN, (), None))
    report("1: 2D Poisson Q1 100x100 (generic_atomic)", model.num_cells(), V.nfree, plan.nnz, dt, tm)
    del plan, assem
    ctx.trim()

# config 3: 3D linear elasticity Q2 vector hex
if want("3"):
    n = N3 or max(4, int(round(24 * scale)))
    model = g.CartesianDiscreteModel((0, 1) * 3, (n, n, n))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 2), dirichlet_tags=[25, 1, 3, 5, 7, 13, 15, 17, 19])
    dO = g.Measure(g.Triangulation(model), 4)
    assem = g.SparseMatrixAssembler(V, V, ctx=ctx)
    t0 = time.perf_counter()
    plan = assem.plan(dO)
    tsym = time.perf_counter() - t0
    E, NU = 2.1e4, 0.3
    lam, mu = E * NU / ((1 + NU) * (1 - 2 * NU)), E / (2 * (1 + NU))
    dt, tm = timed(plan, lambda: plan.assemble_matrix(lib.FORM_ELASTICITY, (lam, mu), None), steps=3, warm=1)
    flops = 27 * (81 * 81 * 12.0) * 2 * model.num_cells()  # ~ what the closed-form integrand costs per (p,i,j)
    report("3: 3D linear elasticity Q2 vector hex %d^3 (%s)" % (n, plan.kernel_path(lib.FORM_ELASTICITY)), model.num_cells(), V.nfree, plan.nnz, dt, tm,
           {"plan_s": tsym, "approx_gflops": flops / dt / 1e9})
    del plan, assem
    ctx.trim()

# config 4: Stokes Taylor-Hood P2/P1 on tets
if want("4"):
    n = N4 or max(3, int(round(20 * scale)))
    model = g.simplexify(g.CartesianDiscreteModel((0, 1) * 3, (n, n, n)))
    Vv = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 2), dirichlet_tags="boundary")
    Q = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1))
    Y = g.MultiFieldFESpace([Vv, Q])
    dO = g.Measure(g.Triangulation(model), 4)
    assem = g.SparseMatrixAssembler(Y, Y, ctx=ctx)
    t0 = time.perf_counter()
    plan = assem.plan(dO, np.array([[1, 1], [1, 0]], dtype=np.uint8))
    tsym = time.perf_counter() - t0
    dt, tm = timed(plan, lambda: plan.assemble_matrix(lib.FORM_STOKES, (), None), steps=3, warm=1)
    report("4: Stokes Taylor-Hood P2/P1, %d tets (%s)" % (model.num_cells(), plan.kernel_path(lib.FORM_STOKES)), model.num_cells(), Y.num_free_dofs(), plan.nnz, dt, tm,
           {"plan_s": tsym})
    del plan, assem
    ctx.trim()

# config 5: neo-Hookean Q1 vector hex, residual + Jacobian
if want("5"):
    n = N5 or max(4, int(round(64 * scale)))
    model = g.CartesianDiscreteModel((0, 1) * 3, (n, n, n))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, g.VectorValue(3), 1), dirichlet_tags="boundary")
    U = g.TrialFESpace(V, (0.0, 0.0, 0.0))
    dO = g.Measure(g.Triangulation(model), 2)
    uh = g.interpolate(lambda x: 0.05 * (np.sin(np.pi * x[:, 0]) * np.sin(np.pi * x[:, 1]) * np.sin(np.pi * x[:, 2]))[:, None] * np.ones((1, 3)), U)
    assem = g.SparseMatrixAssembler(U, V, ctx=ctx)
    t0 = time.perf_counter()
    plan = assem.plan(dO)
    tsym = time.perf_counter() - t0
    plan.set_state(0, uh.free_values, uh.dirichlet_values)


    def newton_assembly():
        plan.assemble_vector(lib.FORM_NEOHOOKEAN_RES, (100.0, 1.0), None, None)
        plan.assemble_matrix(lib.FORM_NEOHOOKEAN_JAC, (100.0, 1.0), None)


    dt, tm = timed(plan, newton_assembly, steps=3, warm=1)
    report("5: neo-Hookean Q1 vector hex %d^3, residual + Jacobian (%s)" % (n, plan.kernel_path(lib.FORM_NEOHOOKEAN_JAC)), model.num_cells(), V.nfree, plan.nnz, dt, tm, {"plan_s": tsym})
    # residual_and_jacobian!: one fused pass (geometry and the constitutive state computed once)
    dt, tm = timed(plan, lambda: plan.assemble_matrix_and_vector(lib.FORM_NEOHOOKEAN_JAC, (100.0, 1.0), lib.FORM_NEOHOOKEAN_RES, (100.0, 1.0), None, None, None),
                   steps=3, warm=1)
    report("5f: neo-Hookean Q1 vector hex %d^3, fused residual_and_jacobian (%s)" % (n, plan.kernel_path(lib.FORM_NEOHOOKEAN_JAC)), model.num_cells(), V.nfree, plan.nnz, dt, tm)
    del plan, assem
    ctx.trim()

# config 2, general-geometry variant: same connectivity, interior nodes displaced by 0.2 dx U(-1,1)^3 -> non-affine cells
if want("2"):
    n = N2B or max(8, int(round(128 * scale)))
    model = g.CartesianDiscreteModel((0, 1) * 3, (n, n, n))
    X = model.node_coordinates
    rng = np.random.default_rng(12345)
    inner = np.all((X > 1e-9) & (X < 1 - 1e-9), axis=1)
    X[inner] += 0.2 / n * rng.uniform(-1, 1, size=(int(inner.sum()), 3))
    V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags="boundary")
    dO = g.Measure(g.Triangulation(model), 2)
    assem = g.SparseMatrixAssembler(V, V, ctx=ctx)
    plan = assem.plan(dO)
    dt, tm = timed(plan, lambda: plan.assemble_matrix(lib.FORM_LAPLACIAN, (), None), steps=3, warm=1)
    report("2b: 3D Poisson Q1 hex %d^3, perturbed (non-affine) mesh (%s)" % (n, plan.kernel_path(lib.FORM_LAPLACIAN)), model.num_cells(), V.nfree, plan.nnz, dt, tm)
    del plan, assem
    ctx.trim()

