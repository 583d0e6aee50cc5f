#!/usr/bin/env python
"""Turn gpurun_out/{launches_<tag>.csv, prof_mlp_tc_<tag>.ncu-rep} into a committed text summary under profiles/.
Usage: python tools/summarize_profile.py <tag> [note] [workload]"""
import collections
import csv
import io
import subprocess
import sys

tag = sys.argv[1]
note = sys.argv[2] if len(sys.argv) > 2 else ""
wl = sys.argv[3] if len(sys.argv) > 3 else "c2"
WL_DESC = {"c2": "c2 = 4 images of 320x240 rays x 64 pairs = 19,660,800 points",
           "c3": "c3 = 8 images of 640x480 rays x 64 pairs = 157,286,400 points (the bench default)"}[wl]
out = io.StringIO()
out.write(f"# ncu summary {tag}\n\n{note}\n\n")
out.write(f"Command profiled: `python bench.py --workload {wl} --steps 2 --warmup 3 --no-e2e --no-cpu-baseline` "
          f"({WL_DESC}; per-launch times are cold-cache/serialised: compare shares).\n\n")

rows = list(csv.reader(open(f"gpurun_out/launches_{tag}.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
iK, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    k = r[iK].split("(")[0][:60]
    v = float(r[iV].replace(",", ""))
    v = {"us": v / 1e3, "ns": v / 1e6, "s": v * 1e3}.get(r[iU], v)
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
ours = {k: v for k, v in agg.items() if k.startswith(("k_", "void k_"))}
tot = sum(v[1] for v in ours.values())
out.write("## launch list (`--metrics gpu__time_duration.sum --clock-control none`), this library's kernels only\n\n")
out.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
for k, (n, ms) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
    out.write(f"| {k} | {n} | {ms:.3f} | {100 * ms / tot:.1f}% |\n")
out.write(f"\n(total {tot:.1f} ms; torch kernels of the synthetic-input generator excluded)\n\n")

raw = subprocess.run(["ncu", "-i", f"gpurun_out/prof_mlp_tc_{tag}.ncu-rep", "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h, u, v = rr[0], rr[1], rr[2]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__pcsamp_sample_count", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_selected",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second"]
out.write("## `ncu --set full` of the decoder kernel (k_mlp_tc<3>), one launch\n\n| metric | value | unit |\n|---|---|---|\n")
for name, unit, val in zip(h, u, v):
    if name in want:
        out.write(f"| {name} | {val} | {unit} |\n")
src = subprocess.run(["ncu", "-i", f"gpurun_out/prof_mlp_tc_{tag}.ncu-rep", "--page", "source", "--csv"],
                     capture_output=True, text=True).stdout
sr = list(csv.reader(io.StringIO(src)))
sh, sd = sr[1], sr[2:]
iS, iSrc = sh.index("# Samples"), sh.index("Source")
tots = sum(int(r[iS]) for r in sd)
out.write(f"\n## hottest SASS instructions by warp-stall samples (total {tots})\n\n| samples | share | instruction |\n|---|---|---|\n")
for i in sorted(sorted(range(len(sd)), key=lambda i: -int(sd[i][iS]))[:12]):
    out.write(f"| {sd[i][iS]} | {100 * int(sd[i][iS]) / tots:.1f}% | `{sd[i][iSrc].strip()[:90]}` |\n")
mn = collections.Counter()
for r in sd:
    op = r[iSrc].strip().split()[0] if r[iSrc].strip() else ""
    if op.startswith("@"):
        op = r[iSrc].strip().split()[1]
    if op.startswith(("UTC", "UBLKCP", "LDTM", "STTM")):
        mn[op.split(".")[0] if not op.startswith(("LDTM", "STTM")) else op] += 1
out.write("\n## Blackwell-native SASS mnemonics present (static count)\n\n" + ", ".join(f"{k} x{v}" for k, v in sorted(mn.items())) + "\n")
# optional: the two row-prep kernels (image-feature gather = ROIAlign per ray; per-ray / per-voxel layer-1 terms)
import os
if os.path.exists(f"gpurun_out/prof_prep_{tag}.ncu-rep"):
    raw = subprocess.run(["ncu", "-i", f"gpurun_out/prof_prep_{tag}.ncu-rep", "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    h = rr[0]
    cols = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "launch__grid_size", "launch__block_size"]
    cols = [c for c in cols if c in h]
    out.write("\n## `ncu --set full` of the row-prep kernels (one launch each)\n\n| " + " | ".join(cols) + " |\n|" + "---|" * len(cols) + "\n")
    out.write("| " + " | ".join(rr[1][h.index(c)] for c in cols) + " |\n")
    for r in rr[2:]:
        if len(r) == len(h):
            out.write("| " + " | ".join(r[h.index(c)][:40] for c in cols) + " |\n")
open(f"profiles/{tag}_k_mlp_tc.md", "w").write(out.getvalue())
print(out.getvalue())
