"""Three device-resident steps of a bench config, to be run under
   ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv
(cache control off: the chunk pipeline keeps its factor ring in L2 ACROSS kernels, which the default per-kernel flush would hide).
Usage: python scripts/traffic_headline.py [config] [n]"""
import sys

sys.path.insert(0, ".")
import gridap_b200 as g  # noqa: E402
from gridap_b200 import lib  # noqa: E402
from bench import DEFAULT_N, Workload  # noqa: E402

key = sys.argv[1] if len(sys.argv) > 1 else "2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else DEFAULT_N[key]
ctx = lib.Context(0)
w = Workload(key, n)
assem = g.SparseMatrixAssembler(w.U, w.V, ctx=ctx)
plan, form, step = w.make_step(assem)
for _ in range(3):
    step()
ctx.synchronize()
print("path", plan.kernel_path(form), "nnz", plan.nnz)
