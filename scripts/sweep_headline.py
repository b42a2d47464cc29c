"""Sweep of the chunk-pipeline tunables of the headline path (256^3 Q1 Poisson, device-resident step).
Usage: python scripts/sweep_headline.py [n] -> one JSON line per setting."""
import json
import os
import sys
import time

sys.path.insert(0, ".")
import torch  # noqa: E402

import gridap_b200 as g  # noqa: E402
from gridap_b200 import lib  # noqa: E402
from bench import Workload  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = lib.Context(0)
w = Workload("2", n)
stream = torch.cuda.ExternalStream(ctx.stream(), device=0)
SETTINGS = [
    ("diag", dict()),
    ("full6", dict(GB200_GATHER_DIAG=0)),
    ("diag_minb5", dict(GB200_GATHER_DIAG_MINB5=1)),
]
only = os.environ.get("ONLY")
for name, env in SETTINGS:
    if only and name not in only.split(","):
        continue
    for k in list(os.environ):
        if k.startswith("GB200_"):
            del os.environ[k]
    os.environ.update({k: str(v) for k, v in env.items()})
    assem = g.SparseMatrixAssembler(w.U, w.V, ctx=ctx)
    plan, form, step = w.make_step(assem)
    for _ in range(3):
        step()
    ctx.synchronize()
    ctx.timings()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 20
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"setting": name, "env": env, "ms_per_step": ms, "gcells_per_s": n ** 3 / ms / 1e6, "host_issue_ms_per_step": 1e3 * t_issue / steps,
                      "kernels": ctx.timings()}), flush=True)
    del plan, assem, step
    ctx.trim()
