#!/bin/bash
# Runs on the GPU box: gather-kernel variants at 256^3 (device-resident step time per variant).
mkdir -p gpurun_out
for mb in 4 3; do
  for pr in 1 0; do
    GB200_GATHER_MINB=$mb GB200_GATHER_PAIRS=$pr python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/sweep_pairs_${mb}_${pr}.json
    python - <<PY
import json
d = json.load(open("gpurun_out/sweep_pairs_${mb}_${pr}.json"))
print("minb", $mb, "pairs", $pr, "ms/step %.4f" % d["ms_per_step"], d["roofline"]["all_kernels_ms"])
PY
  done
done
