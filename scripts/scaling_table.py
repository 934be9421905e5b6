"""Markdown table of a scaling run from profiles/r02_scale_<workload>_n<N>_<tag>.json.  usage: scaling_table.py <tag>"""
import glob
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "c"
rows = {}
for f in glob.glob(os.path.join(ROOT, "profiles", f"r02_scale_*_n*_{tag}.json")):
    m = re.search(r"r02_scale_(\w+)_n(\d+)_", f)
    d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    rows[(m.group(1), int(m.group(2)))] = d
print("| workload | N | ms/step | samples/s | efficiency | e2e ms/step | e2e samples/s | e2e efficiency | exchange kernel at the end of the step: total / wait for peers' gradients (µs, slowest rank) |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---|")
for wl in sorted({k[0] for k in rows}):
    base = rows.get((wl, 1))
    for n in sorted(k[1] for k in rows if k[0] == wl):
        d = rows[(wl, n)]
        eff = (base["ms_per_step"] / d["ms_per_step"]) if base else float("nan")
        e2e_eff = (base["e2e"]["ms_per_step"] / d["e2e"]["ms_per_step"]) if base else float("nan")
        ph = (d.get("dp_check") or {}).get("phases_us_max_over_ranks") or {}
        ex = f"{ph.get('late_total', 0):.1f} / {ph.get('late_wait_grads', 0):.1f}" if ph else "—"
        print(f"| {wl} | {n} | {d['ms_per_step']:.4f} | {d['value'] / 1e6:.2f} M | {eff:.3f} | {d['e2e']['ms_per_step']:.4f} | "
              f"{d['e2e']['value'] / 1e6:.2f} M | {e2e_eff:.3f} | {ex} |")
