"""SASS opcode evidence per kernel from the built library (run in the build container, no GPU needed):
    python tools/sass_summary.py [viforsdes_b200/libvisde.so] > profiles/r2_sass_summary.md
Counts the mnemonics that prove Blackwell-native code paths (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM / STTM =
tcgen05.ld / st, UTMALDG / UTMASTG = cp.async.bulk.tensor, UBLKCP = cp.async.bulk (1-D), FFMA2 = packed dual fp32 FMA."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "viforsdes_b200/libvisde.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "FFMA2", "FFMA", "MUFU", "HMMA", "LDGSTS"]
per = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        per[cur]["_total"] += 1
        for k in KEYS:
            if op == k or (k.startswith("UTC") and op.startswith(k)):
                per[cur][k] += 1


def demangle(n: str) -> str:
    out = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    out = re.sub(r"visde::\(anonymous namespace\)::|visde::", "", out)
    return re.sub(r"\(.*", "", out)[:80]


tot = collections.Counter()
print("# SASS opcode summary of `viforsdes_b200/libvisde.so` (sm_100a)\n")
print("`cuobjdump -sass`, mnemonic counts per kernel (static code, not executed counts). UTC*MMA = `tcgen05.mma`, LDTM / STTM = "
      "`tcgen05.ld / st`, UTMALDG / UTMASTG = TMA tensor copies, UBLKCP = `cp.async.bulk`, SYNCS = mbarrier ops, FFMA2 = packed fp32 FMA; "
      "HMMA (legacy `mma.sync`) must be absent.\n")
print("| kernel | SASS instr | " + " | ".join(KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
rows = []
for fn, c in per.items():
    tot.update(c)
    if any(c[k] for k in KEYS[:8]) or c["FFMA2"] > 50:
        rows.append((demangle(fn), c))
for name, c in sorted(rows, key=lambda r: -r[1]["_total"]):
    print(f"| `{name}` | {c['_total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |")
print(f"\n**Library totals** ({len(per)} kernels, {tot['_total']} SASS instructions): " + ", ".join(f"{k} {tot[k]}" for k in KEYS))
