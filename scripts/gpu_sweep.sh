#!/bin/bash
# Runs on the GPU box: gather-kernel variants at 256^3 (device-resident step time per variant).
mkdir -p gpurun_out
for mb in 4 3; do
    GB200_GATHER_MINB=$mb python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/sweep_meta_${mb}.json
    python - <<PY
import json
d = json.load(open("gpurun_out/sweep_meta_${mb}.json"))
print("minb", $mb, "ms/step %.4f" % d["ms_per_step"], d["roofline"]["all_kernels_ms"], "e2e ms", d["e2e"]["ms_per_step"])
PY
done
