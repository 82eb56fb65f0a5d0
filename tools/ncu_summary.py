"""Summarise ncu outputs into markdown for profiles/ (run in the build container)."""
import collections, csv, re, subprocess, sys

def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("visde::<unnamed>::", "")[:70]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = ["| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{n}` | {c} | {t:.1f} | {t / c:.1f} | {t / tot:.1%} |")
    return "\n".join(out)

def raw(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
            "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        out.append(f"### `{d['Kernel Name'][:90]}`\n")
        for k in keys:
            if k in d and d[k] != "":
                out.append(f"- {k}: {d[k]}")
        st = [(k.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(d[k])) for k in hdr
              if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k and d.get(k, "") != ""]
        tot = sum(v for _, v in st) or 1
        top = sorted(st, key=lambda kv: -kv[1])[:6]
        out.append("- warp stall samples: " + ", ".join(f"{k} {v / tot:.0%}" for k, v in top))
        out.append("")
    return "\n".join(out)

if __name__ == "__main__":
    if sys.argv[1] == "launches":
        print(launches(sys.argv[2]))
    else:
        print(raw(sys.argv[2]))
