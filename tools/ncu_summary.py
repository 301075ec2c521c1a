#!/usr/bin/env python
"""Turn the scratch ncu captures of one round into the tracked summaries under profiles/.

    python tools/ncu_summary.py --full gpurun_out/r04a_top.ncu-rep --launches gpurun_out/r04a_launches.csv \
        --breakdown gpurun_out/r04a_breakdown.json --tag r01

--full      `ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernelILi(7|4|2)E" --launch-skip 28 -c 14`
            of `DRVAE_B200_GRAPH=0 python bench.py --steps 1 --warmup 3 --no-cpu-baseline`: the 14 big-GEMM launches of the
            third step in launch order -> profiles/<tag>_ncu_full_top_kernels.csv + profiles/ncu_traffic.json
--launches  `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv` launch list of the same command
            -> profiles/<tag>_ncu_launches.csv (copy) + profiles/<tag>_ncu_launches_summary.md
"""
import argparse
import collections
import csv
import io
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# launch order of the kernels matched by the --full capture inside one DrVAE training step (run_step in plan.cu)
STEP_ORDER = ["dz1.bwd:gemm_dx.head", "dz1.bwd:gemm_dw_adam.head", "dz1.bwd:gemm_dw_adam.h0",
              "z3.bwd:gemm_dx.head", "z3.bwd:gemm_dw_adam.head", "z3.bwd:gemm_dw_adam.h0",
              "dec.fwd:gemm_nt_decloss.head", "dec.bwd:gemm_dx.head", "dec.bwd:gemm_dw_adam.head", "dec.bwd:gemm_dw_adam.h0",
              "T.bwd:gemm_dw_adam.head", "enc.bwd:gemm_dx.head", "enc.bwd:gemm_dw_adam.head", "enc.bwd:gemm_dw_adam.h0"]

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warp_latency_per_inst_issued.ratio"]


def to_bytes(value, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(value) * scale.get(unit, 1.0)


def full_summary(rep, tag, note):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    idx = {n: i for i, n in enumerate(head)}
    if len(data) != len(STEP_ORDER):
        raise SystemExit("expected %d captured kernels, found %d" % (len(STEP_ORDER), len(data)))
    out = os.path.join(ROOT, "profiles", "%s_ncu_full_top_kernels.csv" % tag)
    traffic = {"source_report": "ncu --set full --clock-control none (profiles/%s_ncu_full_top_kernels.csv)" % tag, "kernels": {}}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["# ncu --set full --clock-control none (graphs off so every kernel is a launch), bench.py default workload "
                    "(32 DrVAE models x 150 rows), third step after warm-up; %s; source report %s (scratch)" % (note, rep)])
        cols = [m for m in METRICS if m in idx]
        w.writerow(["step_kernel", "Kernel Name"] + cols)
        w.writerow(["unit", ""] + [units[idx[m]] for m in cols])
        for tagk, row in zip(STEP_ORDER, data):
            w.writerow([tagk, row[idx["Kernel Name"]]] + [row[idx[m]] for m in cols])
            traffic["kernels"][tagk] = {
                "dram_read_bytes": to_bytes(row[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]),
                "dram_write_bytes": to_bytes(row[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]]),
                "duration_us_under_ncu": float(row[idx["gpu__time_duration.sum"]]),
                "source": "profiles/%s_ncu_full_top_kernels.csv" % tag}
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    print("wrote", out, "and profiles/ncu_traffic.json")


def launches_summary(path, breakdown, tag, note):
    dst = os.path.join(ROOT, "profiles", "%s_ncu_launches.csv" % tag)
    if os.path.abspath(path) != os.path.abspath(dst):
        shutil.copyfile(path, dst)
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    idx = {n: i for i, n in enumerate(rows[0])}
    data = rows[1:]
    names = [r[idx["Kernel Name"]] for r in data]
    starts = [i for i, n in enumerate(names) if "set_dyn" in n]
    s, e = starts[-2], starts[-1]  # the last complete step of the capture
    agg = collections.OrderedDict()
    for r in data[s:e]:
        n = r[idx["Kernel Name"]].replace("drvae::", "").replace("void ", "")
        n = n.split("(")[0]
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += float(r[idx["Metric Value"]]) / 1e3
    total = sum(a[1] for a in agg.values())
    bd = json.load(open(breakdown)) if breakdown else None
    with open(os.path.join(ROOT, "profiles", "%s_ncu_launches_summary.md" % tag), "w") as f:
        f.write("# ncu launch list, one steady-state DrVAE ensemble step (32 models x 150 rows), %s\n\n" % note)
        f.write("Command: `DRVAE_B200_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py "
                "--steps 2 --warmup 3 --no-cpu-baseline`\n(raw list: `profiles/%s_ncu_launches.csv`; graphs off so that every kernel is a "
                "separate launch).  Under ncu every launch is serialised\nand cold-cache, so only the SHARES are comparable with "
                "`bench.py`'s live CUDA-event breakdown (`profiles/%s_breakdown.json`).\n\n" % (tag, tag))
        f.write("| kernel | launches / step | us / step (ncu) | share (ncu) |\n|---|---|---|---|\n")
        for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f %% |\n" % (n, c, us, 100.0 * us / total))
        f.write("\nTotal %.1f us serialised under ncu; %d launches per step." % (total, e - s))
        if bd:
            f.write("  bench.py: %.1f us per step (two-stream schedule + graph replay), %.1f us as the sum of serialised event times.\n"
                    % (bd["ms_per_step"] * 1e3, bd["profiled_ms_per_step"] * 1e3))
            big = [k for k in bd["kernels"] if "gemm_dw_adam" in k["kernel"]]
            f.write("\nFused weight-gradient + Adam kernels (`gemm_tc_kernel<7, ...>`): %.1f us of the event-timed %.1f us (%.1f %%).\n"
                    % (sum(k["ms_per_launch"] for k in big) * 1e3, bd["profiled_ms_per_step"] * 1e3,
                       100.0 * sum(k["ms_per_launch"] for k in big) / bd["profiled_ms_per_step"]))
        else:
            f.write("\n")
    print("wrote profiles/%s_ncu_launches_summary.md" % tag)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full")
    ap.add_argument("--launches")
    ap.add_argument("--breakdown")
    ap.add_argument("--tag", default="r01")
    ap.add_argument("--note", default="final kernels of round 1")
    a = ap.parse_args()
    if a.full:
        full_summary(a.full, a.tag, a.note)
    if a.launches:
        launches_summary(a.launches, a.breakdown, a.tag, a.note)


if __name__ == "__main__":
    main()
