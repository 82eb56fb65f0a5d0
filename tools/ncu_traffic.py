"""ncu CSV (--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum over a kernel-by-kernel bench run)
-> per-stage DRAM bytes per iteration, merged into profiles/ncu_traffic.json[workload] (bench.py reads `traffic` from it).
    python tools/ncu_traffic.py <csv> <workload> <iterations captured>"""
import collections
import csv
import json
import re
import sys
from pathlib import Path

root = Path(__file__).resolve().parents[1]
path, workload, iters = sys.argv[1], sys.argv[2], int(sys.argv[3])
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr, data = rows[hi], rows[hi + 1:]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
kern = collections.OrderedDict()  # id -> {name, metrics}
for r in data:
    if len(r) <= vi:
        continue
    k = kern.setdefault(r[0], {"name": r[ki], "m": {}})
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
    k["m"][r[mi]] = v * scale


def stage_of(name: str, after_elbo_bwd: bool) -> str | None:
    n = name
    if "elbo_fwd" in n:
        return "K5_elbo_fwd"
    if "elbo_bwd" in n:
        return "K6_elbo_bwd"
    if re.search(r"path_fwd|tcw_expand|gth_kernel", n):
        return "K1_path_fwd" if "gth" not in n else "K0_ctx_gemm"
    if re.search(r"path_bwd|thin|fast_partials", n):
        return "K2_path_bwd"
    if re.search(r"tcw_images|tcw_tile", n):
        return "K2_path_bwd" if after_elbo_bwd else "K1_path_fwd"
    if re.search(r"tc_wgrad|gemm_tn|wgrad", n):
        return "K4_wgrad"
    if re.search(r"theta_grads|gemm_nn", n):
        return "K3_grad_ctx"
    if re.search(r"tc_rows_kernel|gemm_nt|split_weights|ctx_bf16", n):
        return "K3_grad_ctx" if after_elbo_bwd else "K0_ctx_gemm"
    return None  # PyTorch kernels (user SDE, adds, flush)


stages = collections.defaultdict(lambda: {"dram_bytes": 0.0, "us": 0.0, "launches": 0})
other = {"dram_bytes": 0.0, "us": 0.0, "launches": 0}
after = False
for k in kern.values():
    nm = k["name"]
    st = stage_of(nm, after)
    if st == "K4_wgrad":
        after = False  # K4 closes the iteration: what follows is the next iteration's K0 / K1
    if "elbo_bwd" in nm:
        after = True
    b = k["m"].get("dram__bytes_read.sum", 0.0) + k["m"].get("dram__bytes_write.sum", 0.0)
    t = k["m"].get("gpu__time_duration.sum", 0.0)
    tgt = stages[st] if st else other
    if st is None and "FillFunctor<unsigned char>" in nm:
        continue  # the bench's L2 flush
    tgt["dram_bytes"] += b
    tgt["us"] += t
    tgt["launches"] += 1
out_path = root / "profiles" / "ncu_traffic.json"
allw = json.loads(out_path.read_text()) if out_path.exists() else {}
rec = {s: v["dram_bytes"] / iters for s, v in stages.items()}
rec["_pytorch_user_sde_and_glue"] = other["dram_bytes"] / iters
rec["_detail"] = {s: {"dram_mb_per_iteration": v["dram_bytes"] / iters / 1e6, "ncu_us_per_iteration": v["us"] / iters,
                      "launches_per_iteration": v["launches"] / iters} for s, v in list(stages.items()) + [("pytorch", other)]}
rec["_source"] = f"{Path(path).name}: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum, {iters} iterations"
allw[workload] = rec
out_path.write_text(json.dumps(allw, indent=1))
for s, v in rec["_detail"].items():
    print(f"{s:14s} {v['dram_mb_per_iteration']:9.1f} MB  {v['ncu_us_per_iteration']:9.1f} us  {v['launches_per_iteration']:.1f} launches")
