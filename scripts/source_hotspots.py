#!/usr/bin/env python
"""Per-source-line view of one kernel of an `ncu --set full --import-source on` capture: share of executed instructions, of
stall samples (with the top stall reasons) and shared-memory wavefronts.
    python scripts/source_hotspots.py gpurun_out/r02j/prof_bwd.ncu-rep ssg_plane_bwd [top-N] [samples|inst]"""
import collections, csv, io, subprocess, sys


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    key = sys.argv[4] if len(sys.argv) > 4 else "samples"
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel,
                          "--print-source", "cuda,sass"], capture_output=True, text=True, errors="replace").stdout
    cur, agg, tot_inst, tot_samp = None, {}, 0, 0
    ie = isamp = iw = None
    stall_cols = []
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            ie, isamp, iw = r.index("Instructions Executed"), r.index("# Samples"), r.index("L1 Wavefronts Shared")
            stall_cols = [(i, h) for i, h in enumerate(r) if h.startswith("stall_") and "Not Issued" not in h]
        elif r[0] not in ("", "Function Name") and ie is not None:
            try:
                inst, samp = int(r[ie]), int(r[isamp])
            except ValueError:
                continue
            wf = int(r[iw]) if r[iw] not in ("", "-") else 0
            st = {h[6:]: int(r[i]) for i, h in stall_cols if r[i] not in ("", "-")}
            agg[(cur, int(r[0]))] = (inst, samp, st, r[1].strip()[:70], wf)
            tot_inst += inst
            tot_samp += samp
    print(f"{kernel}: {tot_inst / 1e6:.1f} M warp instructions, {tot_samp} stall samples")
    idx = 1 if key == "samples" else 0
    for (f, l), (inst, samp, st, src, wf) in sorted(agg.items(), key=lambda kv: -kv[1][idx])[:top_n]:
        big = [(k, round(100 * v / samp)) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3]] if samp else []
        print(f"{f[:18]:18s}:{l:4d}  samples {100 * samp / tot_samp:5.1f}%  inst {100 * inst / tot_inst:5.1f}%  "
              f"wavefronts {wf / 1e6:6.1f} M  {big}  {src}")
    tot = collections.Counter()
    for v in agg.values():
        tot.update(v[2])
    s = sum(tot.values())
    print("stall reasons:", {k: round(100 * v / s, 1) for k, v in tot.most_common(10)})


if __name__ == "__main__":
    main()
