#!/bin/bash
# Runs on the GPU box: gather-kernel variants at 256^3 (device-resident step time per variant).
mkdir -p gpurun_out
for fu in 1 0; do
    GB200_GATHER_FUSED=$fu python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/sweep_fused_${fu}.json
    python - <<PY
import json
d = json.load(open("gpurun_out/sweep_fused_${fu}.json"))
print("fused", $fu, "ms/step %.4f" % d["ms_per_step"], d["roofline"]["all_kernels_ms"])
PY
done
