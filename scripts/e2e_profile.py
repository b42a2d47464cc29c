"""Breakdown of the end-to-end `assemble_matrix(a,U,V)` call at n^3 cells (host wall clock per phase)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import gridap_b200 as g  # noqa: E402
from gridap_b200 import lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = lib.Context(0)
model = g.UnstructuredDiscreteModel(g.CartesianDiscreteModel((0, 1) * 3, (n, n, n)))
V = g.TestFESpace(model, g.ReferenceFE(g.lagrangian, float, 1), dirichlet_tags="boundary")
U = g.TrialFESpace(V, 0.0)
dO = g.Measure(g.Triangulation(model), 2)
a = lambda u, v: g.Integral(g.inner(g.grad(v), g.grad(u))) * dO  # noqa: E731


def T(label, t0):
    ctx.synchronize()
    t1 = time.perf_counter()
    print("  %-28s %8.1f ms" % (label, 1e3 * (t1 - t0)))
    return t1


for rep in range(3):
    print("rep", rep)
    model._device.clear()
    V._device.clear()
    t = t_start = time.perf_counter()
    assem = g.SparseMatrixAssembler(U, V, ctx=ctx)
    matdata = g.collect_cell_matrix(U, V, a(g.get_trial_fe_basis(U), g.get_fe_basis(V)))
    t = T("recognise form", t)
    mesh = model.device_mesh(ctx)
    t = T("mesh upload", t)
    plan = assem.plan(matdata.measure, None)
    t = T("space upload + symbolic", t)
    print("     device timers:", plan.symbolic_timings)
    colptr, rowval = plan.pattern(wait=False)
    t0p = time.perf_counter(); print('  %-28s %8.1f ms' % ('pattern D2H enqueue', 1e3 * (t0p - t))); t = t0p
    nzval = ctx.pinned_empty(plan.nnz, np.float64)
    t = T("alloc nzval", t)
    plan.assemble_matrix(matdata.terms[0].form, matdata.terms[0].params, nzval, False)
    t = T("numeric + nzval D2H", t)
    print("     device timers:", ctx.timings())
    del plan, assem, colptr, rowval, nzval, mesh
    t = T("free", t)
    print("  total %.1f ms" % (1e3 * (t - t_start)))

# the public call exactly as bench.py's e2e leg makes it, under cProfile (host-side view)
import cProfile  # noqa: E402
import pstats  # noqa: E402


def e2e_step():
    asm = g.SparseMatrixAssembler(U, V, ctx=ctx)
    model._device.clear()
    V._device.clear()
    return g.assemble_matrix(a, asm, U, V)


for _ in range(2):
    A = e2e_step()
    del A
ctx.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    A = None
    A = e2e_step()
    ctx.synchronize()
    print("public call: %.1f ms" % (1e3 * (time.perf_counter() - t0)), ctx.timings())
pr = cProfile.Profile()
pr.enable()
A = None
A = e2e_step()
ctx.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
