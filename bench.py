#!/usr/bin/env python
"""bench.py — DrVAE training throughput (samples/s, forward + ELBO + backward + Adam) on B200.

Contract: `python bench.py --gpus N --steps K --warmup W` (under torchrun for N > 1) prints ONE
JSON line from rank 0.  Workload (BASELINE.json configs[4], weak scaling): an ensemble of
independent DrVAE models in the README architecture (dim-z1 100, dim-z3 100, enc-z1 800,
dec-x 600, enc-z3 200, dec-z1 200, L=2, --train-w-noise, batch 150 each), `--models-per-gpu`
(32) models per GPU, i.e. the 256-model drug x fold ensemble at 8 GPUs; members are independent,
so ranks exchange nothing on the data path.  A "step" trains every member of the shard once.

`--impl reference` times the reference's CPU implementation of the same step on this box's host
cores.  /root/reference is not present on the GPU box, so that arm runs the oracle port
(oracle/drvae_oracle.py: same torch CPU ops, autograd and torch.optim.Adam the reference calls;
pinned to the reference in oracle/make_golden.py) on a bounded sample: one member of the ensemble.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

README = dict(dim_x=978, dim_y=2, dim_z1=100, dim_z3=100, enc_z1=[800], dec_x=[600], enc_z3=[200], dec_z1=[200])
L = 2
BATCH = 150
N_PARAMS = 2321758
METRIC = "DrVAE train samples/sec (fwd+bwd+Adam)"
UNIT = "samples/s"


_STDOUT_FD = None


def quiet_stdout():
    """stdout carries exactly one JSON line: native libraries that print there (NCCL's version banner at communicator
    creation) are pointed at stderr for the duration of the run; emit() restores the real stdout for the result."""
    global _STDOUT_FD
    if _STDOUT_FD is None:
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    if _STDOUT_FD is not None:
        os.dup2(_STDOUT_FD, 1)
    print(json.dumps(line), flush=True)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], bf16=p["bf16_tflops"], bf16_sustained=p["bf16_tflops_sustained"], which="measured")
    except Exception:
        return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, which="fallback")


def ncu_traffic(kernel_tag):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel_tag`, from the committed
    `ncu --set full` capture of this workload (profiles/ncu_traffic.json names the source report)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)["kernels"].get(kernel_tag)
        return None if t is None else t["dram_read_bytes"] + t["dram_write_bytes"]
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, n in enumerate(names):
                if len(r) > 5 + i and r[5 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference step on host cores
# ------------------------------------------------------------------------------------------------
def cpu_step_time(threads, steps, warmup):
    import torch
    from oracle import drvae_oracle as orc
    from drvae_b200.init import init_state_dict
    torch.set_num_threads(threads)
    sd = init_state_dict("drvae", seed=123, **README)
    batch = orc.synthetic_batch(BATCH, README["dim_x"])
    om = orc.OracleModel(sd, orc.default_cfg("drvae", L=L))
    for i in range(warmup):
        om.step(batch, orc.Tape(seed=i))
    t0 = time.perf_counter()
    for i in range(steps):
        om.step(batch, orc.Tape(seed=100 + i))
    return (time.perf_counter() - t0) / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = cores
    dt = cpu_step_time(threads, max(1, args.steps), max(1, args.warmup))
    value = BATCH / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "1 ensemble member (DrVAE README config, batch 150) per step; members are independent, "
                                   "so the ensemble's CPU throughput is this figure"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args):
    return {"workload": "DrVAE README-config ensemble (BASELINE configs[4] shard): %d independent models per GPU x batch %d, "
                        "L=2, train-w-noise, Philox noise drawn in-step" % (args.models_per_gpu, BATCH),
            "models_per_gpu": args.models_per_gpu, "batch_per_model": BATCH, "parallelism": "ensemble-shard x%d (no collective)" % args.gpus,
            "l2_policy": "working set per step (%.1f GB of parameters, Adam state and activations) exceeds the 126 MB L2"
                         % (args.models_per_gpu * N_PARAMS * 16 / 1e9)}


# ------------------------------------------------------------------------------------------------
# algorithmic work per launch class: GEMM FLOPs from the valid dims
# ------------------------------------------------------------------------------------------------
def gemm_dims(N=BATCH):
    Np, Nlab = (N + 1) // 2, N - (N + 2) // 3
    R0, LN, Rd, F = N + Np, L * N, L * (N + 2 * Np), L * (Nlab + 2 * (N - Nlab))
    a = README
    d = {}

    def block(name, rows, kin, hidden, out2, first_dx):
        prev = kin
        for i, h in enumerate(hidden):
            d["%s.fwd:gemm_nt.h%d" % (name, i)] = (rows, h, prev)
            d["%s.bwd:gemm_dw.h%d" % (name, i)] = (h, prev, rows)
            if i > 0 or first_dx:
                d["%s.bwd:gemm_dx.h%d" % (name, i)] = (rows, prev, h)
            prev = h
        d["%s.fwd:gemm_nt.head" % name] = (rows, out2, prev)
        d["%s.bwd:gemm_dw.head" % name] = (out2, prev, rows)
        d["%s.bwd:gemm_dx.head" % name] = (rows, prev, out2)

    block("enc", R0, a["dim_x"], a["enc_z1"], 2 * a["dim_z1"], False)
    block("dec", Rd, a["dim_z1"], a["dec_x"], 2 * a["dim_x"], True)
    d["dec.fwd:gemm_nt_decloss.head"] = d.pop("dec.fwd:gemm_nt.head")
    block("z3", F, a["dim_z1"], a["enc_z3"], 2 * a["dim_z3"], True)
    block("dz1", F, a["dim_z3"], a["dec_z1"], 2 * a["dim_z1"], True)
    d["T.fwd:gemm_nt.head"] = (LN, 2 * a["dim_z1"], a["dim_z1"])
    d["T.bwd:gemm_dw.head"] = (2 * a["dim_z1"], a["dim_z1"], LN)
    d["T.bwd:gemm_dx.head"] = (LN, a["dim_z1"], 2 * a["dim_z1"])
    return d


def run_dp8192(args):
    """BASELINE configs[3]: DrVAE README architecture, ONE model, global minibatch 8192, rows sharded over the
    ranks (drvae_b200.dp): drvae_grad_step per shard with global normalisers -> per-bucket NCCL all-reduce on a side
    stream, overlapped with the rest of backward -> replicated drvae_adam_step.  Strong scaling: the global batch is fixed."""
    import torch
    from drvae_b200 import dp as dpm
    from drvae_b200.init import init_state_dict
    from drvae_b200.plan import Plan
    from oracle import drvae_oracle as orc  # synthetic batch generator only

    GLOBAL_N = 8192
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = dpm.shard_rows(GLOBAL_N, world, rank)
    plan = Plan("drvae", L=L, max_batch=hi - lo, n_models=1, **README)
    plan.load_state_dict(init_state_dict("drvae", seed=123, **README))
    full = orc.synthetic_batch(GLOBAL_N, README["dim_x"], seed=0)
    host = {k: full[k][lo:hi].contiguous().pin_memory() for k in ("x1", "x2", "y", "has_x2", "has_y")}
    devb = {k: v.to(dev) for k, v in host.items()}
    runner = dpm.DataParallel(dpm.PlanBackend(plan))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        runner.step(devb, seed=1, row_offset=lo, host_flags=host)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    l0 = plan.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        losses = runner.step(devb, seed=1, row_offset=lo, host_flags=host)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = plan.launch_count() - l0
    value = GLOBAL_N * args.steps / (ms / 1e3)
    # end to end: pinned host shard -> H2D -> step -> loss D2H, every step
    from drvae_b200.feed import DeviceFeeder
    feeder = DeviceFeeder(dev)

    loss_host = [torch.empty(8).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event() for _ in range(2)]

    def e2e_loop(n):
        # H2D of the shard and D2H of the loss terms every step; the host reads step i's losses after enqueueing step i+1
        feeder.put(host)
        for i in range(n):
            if i + 1 < n:
                feeder.put(host)
            b, slot = feeder.get()
            res = runner.step(b, seed=1, row_offset=lo, host_flags=host)
            feeder.done(slot)
            loss_host[i % 2].copy_(res, non_blocking=True)
            loss_ready[i % 2].record()
            if i > 0:
                loss_ready[(i - 1) % 2].synchronize()
        loss_ready[(n - 1) % 2].synchronize()
        return loss_host[(n - 1) % 2].clone()

    e2e_loop(2)
    barrier()
    t0 = time.perf_counter()
    out = e2e_loop(args.steps)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    dump = os.environ.get("BENCH_BREAKDOWN")
    if dump and rank == 0:
        plan.profile_begin()
        for _ in range(3):
            runner.step(devb, seed=1, row_offset=lo, host_flags=host)
        prof = plan.profile_end()
        tot = sum(v[1] for v in prof.values())
        with open(dump, "w") as f:
            json.dump({"ms_per_step": ms / args.steps, "profiled_ms_per_step": tot / 3,
                       "kernels": sorted(({"kernel": k, "ms_per_launch": v[1] / v[0], "share": v[1] / tot} for k, v in prof.items()),
                                         key=lambda r: -r["share"])}, f, indent=1)
    peaks = measured_peaks()
    flops = 3.106e11  # GEMM FLOPs of one N=8192 DrVAE step (fwd + dX + dW)
    tf = flops / (ms / args.steps * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "DrVAE README config, single model, global batch 8192 (BASELINE configs[3]), rows sharded over %d "
                               "rank(s); per-bucket NCCL all-reduce overlapped with backward; Philox noise keyed by global row" % world,
                   "global_batch": GLOBAL_N, "rows_per_rank": hi - lo, "parallelism": "dp%d" % world,
                   "l2_policy": "inputs and activations of a 8192-row step exceed the 126 MB L2"},
        "clocks": sampler.summary(),
        "e2e": {"value": GLOBAL_N * args.steps / e2e_s, "unit": UNIT,
                "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in host.values()), "d2h_bytes_per_step": out.numel() * 4},
        "gpu_launches": launches,
        "roofline": {"kernel": "whole step (GEMM FLOPs)", "bound": "tensor", "achieved": tf / world, "peak": peaks["bf16_sustained"],
                     "unit": "TFLOP/s", "frac": tf / world / peaks["bf16_sustained"], "traffic": None,
                     "peak_source": "%s (MEASURED_PEAKS.json bf16_tflops_sustained), per GPU" % peaks["which"]},
        "losses": {k: float(out[i]) for i, k in enumerate(("RECL", "KLD", "PERT", "YL", "MMD", "ELBO", "CMPL"))},
    }
    if rank == 0:
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--models-per-gpu", type=int, default=32)
    ap.add_argument("--workload", default="ensemble", choices=["ensemble", "dp8192"],
                    help="ensemble: BASELINE configs[4] shard (default, weak scaling); dp8192: configs[3], one model, "
                         "global batch 8192 row-sharded over the ranks with an NCCL gradient all-reduce (strong scaling)")
    ap.add_argument("--cpu-baseline-steps", type=int, default=60)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "dp8192":
        return run_dp8192(args)

    import torch
    from drvae_b200.init import init_state_dict
    from drvae_b200.plan import Plan, anneal_coef
    from oracle import drvae_oracle as orc  # synthetic batch generator + cpu_baseline leg only

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a GPU (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    M = args.models_per_gpu
    plan = Plan("drvae", L=L, max_batch=BATCH, n_models=M, **README)
    host = {k: [] for k in ("x1", "x2", "y", "has_x2", "has_y")}
    for m in range(M):
        gm = rank * M + m
        plan.load_state_dict(init_state_dict("drvae", seed=1000 + gm, **README), model=m)
        b = orc.synthetic_batch(BATCH, README["dim_x"], seed=gm)
        for k in host:
            host[k].append(b[k])
    host = {k: torch.stack(v).contiguous().pin_memory() for k, v in host.items()}
    devb = {k: v.to(dev) for k, v in host.items()}
    step_no = [0]

    def hp():
        s = step_no[0]
        step_no[0] += 1
        return plan.hparams(step=s, beta_pert=anneal_coef(s, 1, 0))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        plan.train_step(devb, hp(), seed=rank)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    # ---- device-resident timed region ----
    l0 = plan.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        plan.train_step(devb, hp(), seed=rank)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = plan.launch_count() - l0
    samples = world * M * BATCH * args.steps
    value = samples / (ms / 1e3)
    # ---- end-to-end through the public API: pinned host batch -> H2D -> step -> loss D2H, every step ----
    # The user-facing loop (drvae_b200.feed.DeviceFeeder + Plan.train_step): batch i+1 is copied from pinned
    # memory on a copy stream while step i runs; every step ends with a device->host read of its losses.
    from drvae_b200.feed import DeviceFeeder
    feeder = DeviceFeeder(dev)

    loss_host = [torch.empty(M, 8).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event() for _ in range(2)]

    def e2e_loop(n):
        """Every step: H2D of its batch (pinned -> device, on the feeder's copy stream while the previous step runs)
        and a D2H read of its loss terms into pinned memory; the host reads step i's losses right after it has
        enqueued step i+1, so the GPU never waits for the host between steps."""
        feeder.put(host)
        out = None
        for i in range(n):
            if i + 1 < n:
                feeder.put(host)
            b, slot = feeder.get()
            res = plan.train_step(b, hp(), seed=rank)
            feeder.done(slot)
            loss_host[i % 2].copy_(res, non_blocking=True)
            loss_ready[i % 2].record()
            if i > 0:
                loss_ready[(i - 1) % 2].synchronize()
                out = loss_host[(i - 1) % 2].clone()
        loss_ready[(n - 1) % 2].synchronize()
        return loss_host[(n - 1) % 2].clone() if n > 0 else out

    e2e_loop(3)
    barrier()
    t0 = time.perf_counter()
    losses = e2e_loop(args.steps)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    e2e_value = samples / e2e_s
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = losses.numel() * losses.element_size()
    finite = bool(torch.isfinite(losses).all())

    # ---- per-launch breakdown (events around every launch; separate pass) ----
    plan.profile_begin()
    PSTEPS = 3
    for _ in range(PSTEPS):
        plan.train_step(devb, hp(), seed=rank)
    prof = plan.profile_end()
    total_prof = sum(v[1] for v in prof.values())
    peaks = measured_peaks()
    dims = gemm_dims()
    rows = []
    for tag, (n, tms) in prof.items():
        per = tms / n  # ms per launch
        r = {"kernel": tag, "ms_per_launch": per, "share": tms / total_prof}
        base = tag.replace("gemm_dw_adam", "gemm_dw")
        if base != tag and base in dims:
            # weight-gradient GEMM with Adam fused into its epilogue: the kernel streams this layer's
            # optimizer state (read p, m, v; write p, m, v: 24 B/param) -> HBM-bound
            mm, nn, kk = dims[base]
            by = 24.0 * (mm * nn + mm) * M  # weights + biases of the layer
            r.update(bound="hbm", achieved=by / (per * 1e-3) / 1e9, peak=peaks["hbm"], unit="GB/s", algorithmic_bytes=by,
                     gemm_tflops=2.0 * mm * nn * kk * M / (per * 1e-3) / 1e12)
        elif tag in dims:
            mm, nn, kk = dims[tag]
            fl = 2.0 * mm * nn * kk * M
            r.update(bound="tensor", achieved=fl / (per * 1e-3) / 1e12, peak=peaks["bf16_sustained"], unit="TFLOP/s")
        elif tag == "opt:adam":
            by = 24.0 * N_PARAMS * M
            r.update(bound="hbm", achieved=by / (per * 1e-3) / 1e9, peak=peaks["hbm"], unit="GB/s")
        if "achieved" in r:
            r["frac"] = r["achieved"] / r["peak"]
        rows.append(r)
    rows.sort(key=lambda r: -r["share"])
    top = next((r for r in rows if "achieved" in r), None)
    roofline = None
    if top is not None:
        roofline = {"kernel": top["kernel"], "bound": top["bound"], "achieved": top["achieved"], "peak": top["peak"],
                    "unit": top["unit"], "frac": top["frac"], "traffic": ncu_traffic(top["kernel"]),
                    "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full (profiles/ncu_traffic.json); "
                                    "algorithmic bytes per launch = 24 B x parameters of the layer x models",
                    "algorithmic_bytes": top.get("algorithmic_bytes"), "share_of_step": top["share"],
                    "peak_source": "%s (MEASURED_PEAKS.json %s)" % (peaks["which"], "hbm_gbs" if top["bound"] == "hbm" else "bf16_tflops_sustained")}
    step_flops = sum(2.0 * a * b * c for a, b, c in dims.values())
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": workload_config(args),
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "roofline": roofline,
        "step_roofline": {"gemm_tflops": step_flops * M / (ms / args.steps * 1e-3) / 1e12,
                          "gemm_frac_of_bf16_sustained": step_flops * M / (ms / args.steps * 1e-3) / 1e12 / peaks["bf16_sustained"],
                          "adam_state_gbs": 24.0 * N_PARAMS * M / (ms / args.steps * 1e-3) / 1e9,
                          "adam_state_frac_of_hbm": 24.0 * N_PARAMS * M / (ms / args.steps * 1e-3) / 1e9 / peaks["hbm"]},
        "breakdown": [{k: (round(v, 6) if isinstance(v, float) else v) for k, v in r.items()} for r in rows[:12]],
        "losses_finite": finite,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = 4  # the reference's own setting: torch.set_num_threads(4), src/run_drvae.py:38-41
        dt = cpu_step_time(threads, args.cpu_baseline_steps, 3)
        line["cpu_baseline"] = {"value": BATCH / dt, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "%d steps of 1 ensemble member (batch 150) with the oracle port; the reference runs "
                                          "members as separate 4-thread processes" % args.cpu_baseline_steps}
    if rank == 0:
        dump = os.environ.get("BENCH_BREAKDOWN")
        if dump:
            with open(dump, "w") as f:
                json.dump({"ms_per_step": ms / args.steps, "profiled_ms_per_step": total_prof / PSTEPS, "kernels": rows}, f, indent=1)
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
