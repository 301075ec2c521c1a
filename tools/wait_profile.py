#!/usr/bin/env python
"""Where do the roles of the persistent GEMM kernel wait?  (measurement tool, not part of the product path)

Builds the instrumented library variant (`python -m drvae_b200.build --waits`, -DGEMM_PROFILE_WAITS), runs a few
steps of the bench.py ensemble workload kernel by kernel (graphs off) and prints, per (contraction mode, epilogue):
cycles per k-block the producer spends blocked on `empty`, the MMA issuer on `full` / `acc_empty`, the first
epilogue warp on `acc_full`, and the CTA lifetime per tile.  clock64 runs at the SM clock.

    DRVAE_B200_LIB=drvae_b200/lib/libdrvae_b200_waits.so python tools/wait_profile.py [--models 32] [--steps 3]
"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

MODES = ["NT", "DX", "DW"]
EPIS = ["store_f32", "elu_c8", "dact_c8", "grad", "decloss", "decout", "lin_c8", "grad_adam"]
COUNTERS = ["prod_wait_empty", "mma_wait_full", "mma_wait_acc_empty", "epi_wait_acc_full", "epi_loop", "cta_total", "tiles", "kblocks"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--models", type=int, default=32)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    if "DRVAE_B200_LIB" not in os.environ:
        from drvae_b200 import build
        os.environ["DRVAE_B200_LIB"] = build.build(extra_flags=["-DGEMM_PROFILE_WAITS"], variant="waits")
    os.environ["DRVAE_B200_GRAPH"] = "0"
    import torch
    from drvae_b200 import _lib
    from drvae_b200.init import init_state_dict
    from drvae_b200.plan import Plan, anneal_coef
    from oracle import drvae_oracle as orc  # synthetic batch generator only
    import bench

    lib = _lib.load()
    M = args.models
    plan = Plan("drvae", L=bench.L, max_batch=bench.BATCH, n_models=M, **bench.README)
    host = {k: [] for k in ("x1", "x2", "y", "has_x2", "has_y")}
    for m in range(M):
        plan.load_state_dict(init_state_dict("drvae", seed=1000 + m, **bench.README), model=m)
        b = orc.synthetic_batch(bench.BATCH, bench.README["dim_x"], seed=m)
        for k in host:
            host[k].append(b[k])
    devb = {k: torch.stack(v).contiguous().cuda() for k, v in host.items()}
    for s in range(3):
        plan.train_step(devb, plan.hparams(step=s, beta_pert=anneal_coef(s, 1, 0)), seed=0)
    buf = (ctypes.c_ulonglong * (3 * 8 * 8))()
    _lib.check(lib.drvae_debug_wait_stats(buf, 1), "wait_stats reset")
    for s in range(3, 3 + args.steps):
        plan.train_step(devb, plan.hparams(step=s, beta_pert=anneal_coef(s, 1, 0)), seed=0)
    _lib.check(lib.drvae_debug_wait_stats(buf, 0), "wait_stats")
    print("per step (%d models); cycles at the SM clock" % M)
    print("%-4s %-10s %7s %8s | per k-block: %9s %9s | per tile: %10s %10s %10s %10s" %
          ("mode", "epilogue", "tiles", "kblocks", "prod:empty", "mma:full", "mma:accE", "epi:accF", "epi loop", "cta life"))
    for mi, mode in enumerate(MODES):
        for ei, epi in enumerate(EPIS):
            c = [buf[(mi * 8 + ei) * 8 + j] / args.steps for j in range(8)]
            if c[6] == 0:
                continue
            tiles, kb = c[6], max(c[7], 1)
            print("%-4s %-10s %7d %8d | %21.0f %9.0f | %20.0f %10.0f %10.0f %10.0f" %
                  (mode, epi, tiles, kb, c[0] / kb, c[1] / kb, c[2] / tiles, c[3] / tiles, c[4] / tiles, c[5] / tiles))


if __name__ == "__main__":
    main()
