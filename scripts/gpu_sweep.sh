#!/bin/bash
# Runs on the GPU box: gather-kernel variants at 256^3 (device-resident step time per variant).
mkdir -p gpurun_out
for mb in 4 6 8; do
  for ov in 1 2; do
    GB200_GATHER_MINB=$mb GB200_GATHER_OVERSUB=$ov python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/sweep_${mb}_${ov}.json
    python - <<PY
import json
d = json.load(open("gpurun_out/sweep_${mb}_${ov}.json"))
print("minb", $mb, "oversub", $ov, "ms/step %.4f" % d["ms_per_step"], d["roofline"]["all_kernels_ms"])
PY
  done
done
