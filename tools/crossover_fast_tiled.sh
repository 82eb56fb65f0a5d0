#!/bin/bash
# K1 / K2 of the one-trajectory-per-CTA family (fast) against the batch-tiled family (tiled) and what AUTO picks, OU T = 100:
# the measurements behind the cost model in tiled_batch_tile() (path_tiled.cu)
for b in 200 296 444 592 740 1184 2048; do for v in fast tiled auto; do
python bench.py --workload ou_b${b}_t100 --variant $v --steps 8 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stages']
print($b,'$v',round(d['ms_per_step'],3),'K1',round(s['K1_path_fwd']['ms_per_step'],3),'K2',round(s['K2_path_bwd']['ms_per_step'],3))"
done; done
