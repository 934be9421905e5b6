"""Turn the ncu outputs a gpurun call left in gpurun_out/ into the small tracked summaries under profiles/.

  launches.csv (ncu --metrics gpu__time_duration.sum)  -> profiles/r<NN>_launches_<tag>.md   per-kernel share of a step
  prof_*.ncu-rep (ncu --set full)                      -> profiles/r<NN>_ncu_<name>.md        key raw metrics per launch
usage: python scripts/summarize_profiles.py <round> <tag>
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.environ.get("MVAE_PROF_DIR", os.path.join(ROOT, "profiles"))  # on the GPU box: under gpurun_out/
rnd, tag = sys.argv[1], sys.argv[2]
os.makedirs(PROF, exist_ok=True)

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor"]


def launches():
    path = os.path.join(OUT, "launches.csv")
    if not os.path.exists(path):
        return
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    rows = list(csv.DictReader(lines))
    # the bench ran `--steps 2 --warmup 3(min)`: keep the kernels of the LAST step = everything after the last
    # split_kernel of the input batch... simpler and robust: aggregate all launches and report per-kernel share.
    agg = collections.OrderedDict()
    for r in rows:
        name = r["Kernel Name"].split("(")[0].replace("void ", "")
        agg.setdefault(name, []).append(float(r["Metric Value"].replace(",", "")))
    total = sum(sum(v) for v in agg.values())
    with open(os.path.join(PROF, f"r{rnd}_launches_{tag}.md"), "w") as fh:
        fh.write(f"# ncu launch list — `bench.py --steps 2 --warmup 1 --no-graph` ({tag})\n\n")
        fh.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES).\n\n")
        fh.write("| kernel | launches | mean us | total us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            fh.write(f"| `{k}` | {len(v)} | {sum(v)/len(v)/1e3:.2f} | {sum(v)/1e3:.1f} | {100*sum(v)/total:.1f}% |\n")
        fh.write(f"\nall launches: {len(rows)}, total {total/1e3:.1f} us\n")


def full(rep):
    path = os.path.join(OUT, rep)
    if not os.path.exists(path):
        return
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    name = rep.replace(".ncu-rep", "")
    with open(os.path.join(PROF, f"r{rnd}_ncu_{name}_{tag}.md"), "w") as fh:
        fh.write(f"# `ncu --set full --clock-control none --import-source on` — {name} ({tag})\n\n")
        for r in rows[2:]:
            fh.write(f"## {r[hdr.index('Kernel Name')][:90]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n\n")
            fh.write("| metric | value | unit |\n|---|---:|---|\n")
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    fh.write(f"| {m} | {r[i]} | {units[i]} |\n")
            fh.write("\n")


def traffic():
    """profiles/ncu_traffic.json: DRAM bytes per launch of the product-manifold kernels from the `ncu --set full`
    captures of scripts/prof_driver.py pm <signature> (2^22 samples), with a hash of the kernels' sources so that
    bench.py can tell whether a capture still describes the code it runs."""
    import json
    sys.path.insert(0, ROOT)
    import bench
    table = {}
    for rep, sig in (("prof_pm.ncu-rep", "h2,s2,e2"), ("prof_pm_cfg3.ncu-rep", "h6,h6,s6,s6,e6")):
        path = os.path.join(OUT, rep)
        if not os.path.exists(path):
            continue
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

        def val(r, m):
            i = hdr.index(m)
            return float(r[i].replace(",", "")) * scale.get(units[i], 1.0)

        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            kern = "pm_forward_kernel" if "pm_forward" in name else "pm_backward_kernel" if "pm_backward" in name else None
            if kern is None:
                continue
            entry = {"signature": sig, "samples": 1 << 22,
                     "dram_bytes": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
                     "us": val(r, "gpu__time_duration.sum") / 1e3 if units[hdr.index("gpu__time_duration.sum")] in ("ns", "nsecond") else None,
                     "capture": f"profiles/r{rnd}_ncu_{rep.replace('.ncu-rep', '')}_{tag}.md",
                     "source_hash": bench.source_hash(kern)}
            table.setdefault(kern, [])
            table[kern] = [e for e in table[kern] if e["signature"] != sig] + [entry]   # last launch of each wins
    if table:
        with open(os.path.join(PROF, "ncu_traffic.json"), "w") as fh:
            json.dump(table, fh, indent=1)


launches()
traffic()
for rep in sorted(os.listdir(OUT)):
    if rep.endswith(".ncu-rep"):
        full(rep)
print(sorted(os.listdir(PROF)))
