"""Config 2 with the RHS: device-resident `assemble_matrix_and_vector` (Laplacian + source with Dirichlet lifting) and
`assemble_vector` at n^3 cells.  Usage: python scripts/bench_rhs.py [n]   (GB200_NO_Q1_RHS=1 -> generic kernel for the vector)"""
import json
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
U = g.TrialFESpace(V, lambda x: x[:, 0] + x[:, 1])
dO = g.Measure(g.Triangulation(model), 2)
assem = g.SparseMatrixAssembler(U, V, ctx=ctx)
plan = assem.plan(dO)
plan.set_state(0, None, U.dirichlet_values)


def timed(call, steps=10, warm=3):
    for _ in range(warm):
        call()
    ctx.synchronize()
    ctx.timings()
    t0 = time.perf_counter()
    for _ in range(steps):
        call()
    ctx.synchronize()
    return (time.perf_counter() - t0) / steps, ctx.timings()


for name, call in [
    ("assemble_matrix", lambda: plan.assemble_matrix(lib.FORM_LAPLACIAN, (), None)),
    ("assemble_vector", lambda: plan.assemble_vector(lib.FORM_SOURCE, (1.0,), None, None)),
    ("assemble_matrix_and_vector (lifting)", lambda: plan.assemble_matrix_and_vector(lib.FORM_LAPLACIAN, (), lib.FORM_SOURCE, (1.0,), None, None, None)),
]:
    dt, tm = timed(call)
    print(json.dumps({"call": name, "n": n, "ms": dt * 1e3, "cells_per_s": model.num_cells() / dt, "kernels_ms": tm}), flush=True)
