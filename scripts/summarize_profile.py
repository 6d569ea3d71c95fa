#!/usr/bin/env python
"""Turn one gpurun_out/<tag>/ directory (scripts/gpu_check.sh output) into the tracked summary
profiles/<tag>.md: the bench line, the ncu launch list aggregated per kernel (share of the step) and the
key counters of the `ncu --set full` capture.   python scripts/summarize_profile.py gpurun_out/r01a [title]"""
import collections
import csv
import json
import os
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg"]


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    return agg


def main():
    d = sys.argv[1].rstrip("/")
    tag = os.path.basename(d)
    title = sys.argv[2] if len(sys.argv) > 2 else tag
    out = [f"# {title}", "", f"Source: `{d}/` (scratch; produced by `scripts/gpu_check.sh {tag}` under gpurun on one B200).", ""]
    bj = os.path.join(d, "bench.json")
    if os.path.exists(bj) and os.path.getsize(bj):
        line = open(bj).read().strip().splitlines()[-1]
        b = json.loads(line)
        out += ["## bench.py line (not under a profiler)", "", "```json", json.dumps(b, indent=1), "```", ""]
    lc = os.path.join(d, "launches.csv")
    if os.path.exists(lc):
        agg = launches(lc)
        tot = sum(v[1] for v in agg.values())
        out += ["## ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline`",
                "(`--metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: shares, not absolutes)", "",
                "| kernel | launches | total ms | share |", "|---|---|---|---|"]
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
            out.append(f"| `{k[:110]}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.1f}% |")
        out.append("")
    rep = os.path.join(d, "prof.ncu-rep")
    if os.path.exists(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        out += ["## `ncu --set full --clock-control none` capture (per launch)", ""]
        seen = set()
        traffic = {}
        stage_of = {"ssg_plane_fwd_kernel": "ssg_plane_fwd", "ssg_plane_bwd_kernel": "ssg_plane_bwd",
                    "ssg_point_fwd": "ssg_point_fwd", "ssg_point_bwd": "ssg_point_bwd"}

        def to_bytes(v, u):
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            return float(v.replace(",", "")) * scale

        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            if name in seen:
                continue
            seen.add(name)
            for frag, stage in stage_of.items():
                if frag in name and "dram__bytes_read.sum" in hdr:
                    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
                    traffic[stage] = to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw])
            out += [f"### `{name}`", "", "| metric | value | unit |", "|---|---|---|"]
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    out.append(f"| {k} | {r[i]} | {units[i]} |")
            out.append("")
        if traffic:
            # bench.py reads this for roofline.traffic (dram read + write bytes per launch of the kernel)
            with open(os.path.join("profiles", "dram_traffic.json"), "w") as f:
                json.dump(traffic, f, indent=1)
            out += ["DRAM traffic per launch (read + write, bytes): " + json.dumps(traffic), ""]
    os.makedirs("profiles", exist_ok=True)
    path = os.path.join("profiles", tag + ".md")
    open(path, "w").write("\n".join(out) + "\n")
    print(path)


if __name__ == "__main__":
    main()
