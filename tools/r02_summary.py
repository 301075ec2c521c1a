#!/usr/bin/env python
"""Round-2 scratch captures (gpurun_out/) -> tracked summaries under profiles/.

    python tools/r02_summary.py

Inputs (written by tools/r02_profile_final.sh on the GPU box):
  gpurun_out/r02_ncu_launches.csv   `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv` launch list of
                                    `DRVAE_B200_GRAPH=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras`
  gpurun_out/r02_top_raw.csv        `ncu --set full` raw page of the 4th step's decoder-loss GEMM, decoder dX GEMMs and the grouped
                                    dW+Adam kernel
Outputs: profiles/r02_ncu_launches.csv (copy), profiles/r02_ncu_launches_summary.md, profiles/r02_ncu_full_top_kernels.csv,
profiles/ncu_traffic.json (read by bench.py for roofline.traffic).
"""
import collections
import csv
import json
import os
import re
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
           "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]
# capture order of the --set full pass (kernels matching dwadam_kernel|gemm_tc_kernel from the decoder-loss GEMM of step 4 on)
ORDER = ["dec.fwd:gemm_nt_decloss.head", "dec.bwd:gemm_dx.head", "bwd:dw_adam_early", "dec.bwd:gemm_dx.h0", "T.bwd:gemm_dx.head", "enc.bwd:gemm_dx.head", "bwd:dw_adam_rest"]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("drvae::", "")
    return re.sub(r"\(.*$", "", name)


def launches():
    src = os.path.join(OUT, "r02_ncu_launches.csv")
    shutil.copy(src, os.path.join(PROF, "r02_ncu_launches.csv"))
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ix = {k: i for i, k in enumerate(hdr)}
    recs = [(short(r[ix["Kernel Name"]]), float(r[ix["Metric Value"]]) / 1e3) for r in rows[1:] if r[ix["Metric Name"]] == "gpu__time_duration.sum"]
    # one steady-state step: from the last set_dyn_kernel but one to the last dwadam
    starts = [i for i, (k, _) in enumerate(recs) if k.startswith("set_dyn_kernel")]
    ends = [i for i, (k, _) in enumerate(recs) if k.startswith("dwadam_kernel")]
    a = starts[-2] if len(starts) >= 2 and starts[-1] > ends[-1] else starts[-1]
    a = max(s for s in starts if s < ends[-1])
    step = recs[a:ends[-1] + 1]
    agg = collections.OrderedDict()
    for k, us in step:
        n, t = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, t + us)
    total = sum(t for _, t in agg.values())
    lines = ["# ncu launch list, one steady-state DrVAE ensemble step (32 models x 150 rows), final kernels of round 2", "",
             "Command: `DRVAE_B200_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras`",
             "(raw list: `profiles/r02_ncu_launches.csv`; graphs off so that every kernel is a separate launch).  Under ncu every launch is serialised",
             "and cold-cache, so only the SHARES are comparable with `bench.py`'s live CUDA-event breakdown (`breakdown` in `profiles/r02_bench.json`).", "",
             "| kernel | launches / step | us / step (ncu) | share (ncu) |", "|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("| `%s` | %d | %.1f | %.1f %% |" % (k, n, t, 100 * t / total))
    lines += ["", "Total %.1f us serialised under ncu; %d launches per step." % (total, len(step))]
    open(os.path.join(PROF, "r02_ncu_launches_summary.md"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[-3:]))
    return agg, total


def full():
    src = os.path.join(OUT, "r02_top_raw.csv")
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    ix = {k: i for i, k in enumerate(hdr)}
    out_rows = [["step_kernel_tag", "kernel"] + ["%s [%s]" % (m, units[ix[m]]) for m in METRICS if m in ix]]
    traffic = {"source_report": "ncu --set full --clock-control none (profiles/r02_ncu_full_top_kernels.csv)", "kernels": {}}
    for i, r in enumerate(rows[2:]):
        tag = ORDER[i] if i < len(ORDER) else "launch%d" % i
        out_rows.append([tag, short(r[ix["Kernel Name"]])] + [r[ix[m]] for m in METRICS if m in ix])

        def val(m):
            v, u = float(r[ix[m]]), units[ix[m]]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        traffic["kernels"][tag] = {"dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
                                   "duration_us_under_ncu": float(r[ix["gpu__time_duration.sum"]]), "source": "profiles/r02_ncu_full_top_kernels.csv"}
    k = traffic["kernels"]
    if "bwd:dw_adam_early" in k and "bwd:dw_adam_rest" in k:  # bench.py times the two parts as one launch (profiling mode)
        k["bwd:dw_adam_all"] = {"dram_read_bytes": k["bwd:dw_adam_early"]["dram_read_bytes"] + k["bwd:dw_adam_rest"]["dram_read_bytes"],
                                "dram_write_bytes": k["bwd:dw_adam_early"]["dram_write_bytes"] + k["bwd:dw_adam_rest"]["dram_write_bytes"],
                                "duration_us_under_ncu": k["bwd:dw_adam_early"]["duration_us_under_ncu"] + k["bwd:dw_adam_rest"]["duration_us_under_ncu"],
                                "source": "profiles/r02_ncu_full_top_kernels.csv (early + rest parts)"}
    with open(os.path.join(PROF, "r02_ncu_full_top_kernels.csv"), "w", newline="") as f:
        csv.writer(f).writerows(out_rows)
    json.dump(traffic, open(os.path.join(PROF, "ncu_traffic.json"), "w"), indent=1)
    for row in out_rows[1:]:
        print(row[0], row[1][:40], row[2], "us  dram rd", row[6], "wr", row[7])


if __name__ == "__main__":
    launches()
    full()
