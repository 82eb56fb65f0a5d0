"""Top SASS lines by warp-stall samples from an ncu report: python tools/sass_hot.py rep.ncu-rep [N] [context] [kernel#]"""
import csv, subprocess, sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
which = int(sys.argv[4]) if len(sys.argv) > 4 else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
for n, s0 in enumerate(starts):
    if which is not None and n != which:
        continue
    s1 = starts[n + 1] if n + 1 < len(starts) else len(rows)
    sec = rows[s0:s1]
    hi = next(i for i, r in enumerate(sec) if r and r[0] == "Address")
    hdr, data = sec[hi], [r for r in sec[hi + 1:] if len(r) == len(sec[hi])]
    i_src, i_s, i_ex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    d = [(int(r[i_s] or 0), k, r[i_src].strip(), int(r[i_ex] or 0)) for k, r in enumerate(data)]
    tot = sum(x[0] for x in d) or 1
    print(f"== kernel {n}: {sec[0][1][:100]}\ntotal samples {tot}, {len(d)} SASS instructions")
    for s, k, src, ex in sorted(d, reverse=True)[:top]:
        print(f"{s:7d} {100 * s / tot:5.1f}%  #{k:5d} ex={ex:9d}  {src[:100]}")
        for j in range(max(0, k - ctx), k):
            print(f"{'':24s}#{j:5d} {d[j][2][:100]}")
