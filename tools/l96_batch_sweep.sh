#!/bin/bash
# per-GPU shard sizes of the strong-scaled config 5 (8192 / N): tensor-core (wide) family vs the register-resident FAST-S family
for B in ${SWEEP_B:-1024 1536 2048 4096 8192}; do
  for V in fast tc; do
    python bench.py --workload l96_b${B}_t100 --variant $V --steps 10 --warmup 3 --no-extra --no-e2e --no-cpu-baseline 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$B $V', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['stages'].items()})"
  done
done
