"""gpurun_out/parity_log.jsonl (written by tests/conftest.py on the GPU box) -> profiles/parity_r2.md: achieved normwise
errors per tensor per configuration, next to the fp32 reference's own rounding noise and the widening that was allowed.
    python tools/parity_table.py [log] [out]"""
import json
import sys
from collections import OrderedDict
from pathlib import Path

root = Path(__file__).resolve().parents[1]
log = Path(sys.argv[1]) if len(sys.argv) > 1 else root / "gpurun_out" / "parity_log.jsonl"
out = Path(sys.argv[2]) if len(sys.argv) > 2 else root / "profiles" / "parity_r2.md"
rows: "OrderedDict[str, dict]" = OrderedDict()
for line in open(log):
    r = json.loads(line)
    k = r["name"]
    if k not in rows or r["normwise_err"] > rows[k]["normwise_err"]:
        rows[k] = r
groups: "OrderedDict[str, list]" = OrderedDict()
for k, r in rows.items():
    cfg, _, tensor = k.rpartition("/")
    groups.setdefault(cfg or "(misc)", []).append((tensor or k, r))
DETAIL = ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5", "tc/l96s10", "fold/E128")
with open(out, "w") as f:
    f.write("# Achieved parity errors (round 2)\n\n"
            "Source: `tests/` run with `-m gpu` on a B200 (`tests/_util.assert_parity` logs every comparison; this file is "
            "`tools/parity_table.py` over that log). Error = max|cuda - oracle_fp64| / max|oracle_fp64| per tensor (normwise); "
            "`fp32 ref noise` = the same measure for the oracle's own fp32 run; `allowed widening` = min(3 x noise, cap x scale) that "
            "was added to the rtol 1e-4 + 1e-5 x scale elementwise bar (cap 1e-3; 2e-2 for the discontinuous diag-floor test).\n\n"
            "## All comparison groups (one line each)\n\n| group | tensors | worst tensor | worst normwise error | its fp32 ref noise |\n|---|---|---|---|---|\n")
    for cfg, items in groups.items():
        t, r = max(items, key=lambda it: it[1]["normwise_err"])
        f.write(f"| {cfg} | {len(items)} | {t} | {r['normwise_err']:.2e} | {r['fp32_ref_noise_normwise']:.2e} |\n")
    f.write("\n## Per-tensor tables of the BASELINE configurations at size (`tests/test_gpu_at_size.py`) and the wide tensor-core family\n\n")
    for cfg, items in groups.items():
        if not cfg.startswith(DETAIL):
            continue
        worst = max(r["normwise_err"] for _, r in items)
        f.write(f"### {cfg}  (worst {worst:.2e})\n\n| tensor | elements | normwise error | fp32 ref noise | allowed widening |\n|---|---|---|---|---|\n")
        for tensor, r in items:
            f.write(f"| {tensor} | {r['numel']} | {r['normwise_err']:.2e} | {r['fp32_ref_noise_normwise']:.2e} | {r['widening_normwise']:.2e} |\n")
        f.write("\n")
print(f"{len(rows)} tensors in {len(groups)} groups -> {out}")
