"""Host-side breakdown of the end-to-end call of ONE rank of an N-rank run (default: rank 3 of 8), executed on one GPU:
the same call bench.py's e2e leg makes at N > 1 (part.assembler(ctx) + assemble_matrix).  Usage: python scripts/e2e_rank_profile.py [world] [rank]"""
import cProfile
import pstats
import sys
import time

sys.path.insert(0, ".")
import gridap_b200 as g  # noqa: E402
from gridap_b200 import distributed as gd  # noqa: E402
from gridap_b200 import lib  # noqa: E402
from bench import Workload, _fields  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = lib.Context(0)
w = Workload("2", 256)
part = gd.partition(w.model, w.U, w.V, world, rank)
wl = w.localized(part)


def e2e_step():
    asm = part.assembler(ctx)
    wl.model._device.clear()
    for sp in _fields(wl.V):
        getattr(sp, "space", sp)._device.clear()
    return wl.e2e_call(asm)


for _ in range(2):
    A, b = e2e_step()
    del A, b
ctx.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    A = b = None
    A, b = e2e_step()
    ctx.synchronize()
    print("public call: %.1f ms" % (1e3 * (time.perf_counter() - t0)), ctx.timings())
pr = cProfile.Profile()
pr.enable()
A = b = None
A, b = e2e_step()
ctx.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
