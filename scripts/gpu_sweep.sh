#!/bin/bash
# Runs on the GPU box: gather-kernel variants at 256^3 (device-resident step time per variant).
mkdir -p gpurun_out
for pp in 1 0; do
    GB200_GATHER_PIPE=$pp python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/sweep_pipe_${pp}.json
    python - <<PY
import json
d = json.load(open("gpurun_out/sweep_pipe_${pp}.json"))
print("pipe", $pp, "ms/step %.4f" % d["ms_per_step"], d["roofline"]["all_kernels_ms"], "e2e ms", d["e2e"]["ms_per_step"])
PY
done
