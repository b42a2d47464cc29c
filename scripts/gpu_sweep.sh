#!/bin/bash
mkdir -p gpurun_out
for cfg in "1 3" "1 2" "2 3" "2 2"; do
    set -- $cfg
    GB200_FUSED_GEOW=$1 GB200_FUSED_MINB=$2 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/sweep_fused_$1_$2.json
    python - <<PY
import json
d = json.load(open("gpurun_out/sweep_fused_$1_$2.json"))
print("geow", $1, "minb", $2, "ms/step %.4f" % d["ms_per_step"], d["roofline"]["all_kernels_ms"])
PY
done
