#!/bin/bash
# Runs on the GPU box: gather-kernel variants at 256^3 (device-resident step time per variant).
mkdir -p gpurun_out
for mb in 2 3 4; do
  for pf in 0 1 2; do
    GB200_GATHER_MINB=$mb GB200_GATHER_PREFETCH=$pf python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/sweep_${mb}_${pf}.json
    python - <<PY
import json
d = json.load(open("gpurun_out/sweep_${mb}_${pf}.json"))
print("minb", $mb, "prefetch", $pf, "ms/step %.4f" % d["ms_per_step"], d["roofline"]["all_kernels_ms"])
PY
  done
done
