"""Static SASS statistics of libmvae_b200.so -> profiles/r<NN>_sass_summary.md (runs without a GPU: cuobjdump only).
usage: python scripts/sass_summary.py <round>"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mvae_b200", "libmvae_b200.so")
rnd = sys.argv[1] if len(sys.argv) > 1 else "02"
COLS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "MUFU", "REDG", "RED", "ATOMG",
        "LDG", "STG"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                       text=True).stdout.splitlines()
stats, cur, it = collections.OrderedDict(), None, iter(names)
for ln in sass.splitlines():
    if "Function :" in ln:
        cur = re.sub(r"\(.*", "", next(it)).replace("void ", "")
        stats.setdefault(cur, collections.Counter())
    elif cur and re.match(r"^\s+/\*[0-9a-f]+\*/\s+[A-Z@]", ln):
        op = ln.split("*/", 1)[1].split()
        op = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
        stats[cur]["inst"] += 1
        base = op.split(".")[0].rstrip(";")
        for c in COLS:
            if base == c:
                stats[cur][c] += 1
out = os.path.join(ROOT, "profiles", f"r{rnd}_sass_summary.md")
with open(out, "w") as fh:
    fh.write(f"# SASS summary of `mvae_b200/libmvae_b200.so` (round {int(rnd)}, `cuobjdump -sass`, sm_100a; "
             "`scripts/sass_summary.py`)\n\n")
    fh.write("Static instruction counts per kernel.  tcgen05 / TMEM / TMA show up as `UTCHMMA` (tcgen05.mma), `UTCBAR` "
             "(tcgen05.commit), `LDTM` (tcgen05.ld), `UTMALDG` (cp.async.bulk.tensor load), `UTMASTG` (cp.async.bulk.tensor "
             "store: the GEMM's plane outputs, DESIGN.md §3.2), `UBLKCP` (cp.async.bulk), `SYNCS` (mbarrier ops).  No "
             "`STTM`: nothing is written back to TMEM.\n\n")
    fh.write("| kernel | inst | " + " | ".join(COLS) + " |\n|---|---:|" + "---:|" * len(COLS) + "\n")
    for k in sorted(stats):
        fh.write(f"| `{k}` | {stats[k]['inst']} | " + " | ".join(str(stats[k][c]) for c in COLS) + " |\n")
print(out)
