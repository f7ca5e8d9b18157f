#!/usr/bin/env python
"""Summarises an `ncu --set full` report (one step of a workload) into the JSON that bench.py reads for
`roofline.traffic` / `kernels_ncu`:  profiles/r2_ncu_kernels_<workload>.json

    python tools/ncu_summary.py gpurun_out/r2_full_searchp.ncu-rep searchp [--git HASH] [--sha KERNEL_SOURCES_SHA]

One entry per kernel NAME (launches of the same instantiation are summed): time, DRAM bytes (dram__bytes_read.sum +
dram__bytes_write.sum), achieved DRAM bandwidth against the measured copy peak, L2 hit rate, ALU-pipe utilisation,
issue-slot utilisation, achieved occupancy, registers.  The capture is stamped with the git hash and the sha of the
kernel sources it was taken from (bench.kernel_sources_sha): bench.py refuses a capture from other sources."""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}

WANT = {
    "gpu__time_duration.sum": "ms",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_inst",
    "smsp__inst_executed.sum": "warp_instructions",
}


def short_name(full):
    n = full.split("(")[0].strip()
    n = re.sub(r"^void\s+", "", n)
    return n.replace("lgpu::", "")


def main():
    rep, wl = sys.argv[1], sys.argv[2]
    git = sys.argv[sys.argv.index("--git") + 1] if "--git" in sys.argv else subprocess.run(
        ["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    import bench
    sha = sys.argv[sys.argv.index("--sha") + 1] if "--sha" in sys.argv else bench.kernel_sources_sha()
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    col = {h: i for i, h in enumerate(hdr)}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm = float(peaks.get("hbm_gbs", 6533.2))
    agg = {}
    for r in rows[2:]:
        if len(r) <= ki:
            continue
        name = short_name(r[ki])
        a = agg.setdefault(name, {"name": name, "launches": 0, "ms": 0.0, "dram_bytes_read": 0.0, "dram_bytes_write": 0.0,
                                  "_w": []})
        a["launches"] += 1
        vals = {}
        for metric, key in WANT.items():
            if metric not in col:
                continue
            try:
                v = float(r[col[metric]].replace(",", ""))
            except ValueError:
                continue
            v *= UNIT.get(units[col[metric]], 1.0) if key in ("ms", "dram_bytes_read", "dram_bytes_write") else 1.0
            vals[key] = v
        a["ms"] += vals.get("ms", 0.0)
        a["dram_bytes_read"] += vals.get("dram_bytes_read", 0.0)
        a["dram_bytes_write"] += vals.get("dram_bytes_write", 0.0)
        a["_w"].append(vals)
    kernels = []
    for a in agg.values():
        w = a.pop("_w")
        tot = sum(x.get("ms", 0.0) for x in w) or 1.0
        for key in ("l2_hit_pct", "l1_hit_pct", "pipe_alu_pct", "pipe_fma_pct", "issue_active_pct", "achieved_occupancy_pct",
                    "threads_per_inst"):
            a[key] = sum(x.get(key, 0.0) * x.get("ms", 0.0) for x in w) / tot  # time-weighted over the launches
        a["registers"] = max((x.get("registers", 0) for x in w), default=0)
        a["warp_instructions"] = sum(x.get("warp_instructions", 0.0) for x in w)
        a["dram_bytes"] = a["dram_bytes_read"] + a["dram_bytes_write"]
        a["dram_gbs"] = a["dram_bytes"] / (a["ms"] * 1e-3) / 1e9 if a["ms"] else 0.0
        a["hbm_frac"] = a["dram_gbs"] / hbm
        m = re.match(r"swDpxKernel<\(?int\)?(\d+), \(?int\)?(\d+), \(?bool\)?(\w+), \(?bool\)?(\w+)>", a["name"])
        if m:
            a["trace"] = m.group(4) in ("1", "true")
        kernels.append(a)
    kernels.sort(key=lambda k: -k["ms"])
    total = sum(k["ms"] for k in kernels)
    for k in kernels:
        k["share_of_captured_time"] = k["ms"] / total if total else 0.0
    # the two DP passes as single entries in front (what bench.py looks for)
    for trace in (False, True):
        sel = [k for k in kernels if k["name"].startswith("swDpxKernel") and k.get("trace") is trace]
        if sel:
            ms = sum(k["ms"] for k in sel)
            kernels.insert(0, {"name": "swDpxKernel (all classes, DP pass %d)" % (2 if trace else 1), "trace": trace,
                               "launches": sum(k["launches"] for k in sel), "ms": ms,
                               "dram_bytes": sum(k["dram_bytes"] for k in sel),
                               "pipe_alu_pct": sum(k["pipe_alu_pct"] * k["ms"] for k in sel) / ms,
                               "issue_active_pct": sum(k["issue_active_pct"] * k["ms"] for k in sel) / ms,
                               "share_of_captured_time": ms / total if total else 0.0})
    doc = {"workload": wl, "git": git, "kernel_sources_sha": sha, "report": os.path.basename(rep),
           "how": "ncu --set full --clock-control none --import-source on, one serial step (tools/profile_run.py), "
                  "summarised by tools/ncu_summary.py; times are cold-cache and serialised: compare shares, not absolutes",
           "hbm_peak_gbs": hbm, "captured_ms": total, "kernels": kernels}
    path = os.path.join(ROOT, "profiles", f"r2_ncu_kernels_{wl}.json")
    json.dump(doc, open(path, "w"), indent=1)
    print(path)
    for k in kernels[:12]:
        print(f"{k['ms']:8.2f} ms  {k['launches']:3d}x  dram {k['dram_bytes'] / 1e9:7.2f} GB  alu {k.get('pipe_alu_pct', 0):5.1f}%  "
              f"{k['name'][:70]}")


if __name__ == "__main__":
    main()
