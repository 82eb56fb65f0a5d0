#!/bin/bash
# BASELINE config 3: OU batch sweep 1k-64k trajectories on one B200 (AUTO kernel family), one JSON line per size
out=${1:-gpurun_out/sweep_ou.jsonl}
: > $out
for b in 1024 2048 4096 8192 16384 32768 65536; do
  timeout 300 python bench.py --workload ou_b${b}_t100 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline >> $out 2>/dev/null
done
python - <<PY
import json
for l in open("$out"):
    d = json.loads(l)
    st = {k: round(v["ms_per_step"], 3) for k, v in d["stages"].items()}
    print(d["config"]["batch_per_gpu"], round(d["ms_per_step"], 3), f'{d["value"]:.3e}', st)
PY
