"""GPU diagnostic (not a pytest): one forward+backward of every model family / architecture on
the GPU against the bf16-emulating oracle, printing per-loss and per-tensor gradient errors for
both GEMM mainloops.  Run under gpurun; output goes to stdout."""
import sys
import time
import traceback

import torch

sys.path.insert(0, "tests")
from helpers import ARCH, KINDS, L, NROWS, SEED_MODEL, SEED_TAPE, batch_fields, orc, rel_err, rel_l2  # noqa: E402

from drvae_b200.init import init_state_dict  # noqa: E402
from drvae_b200.noise import eps_block_from_tape  # noqa: E402
from drvae_b200.plan import LOSS_KEYS, Plan, anneal_coef  # noqa: E402


def run_case(case, kind, impl, step, verbose=True):
    arch = ARCH[case]
    N = NROWS[case]
    sd = init_state_dict(kind, seed=SEED_MODEL, **arch)
    batch = orc.synthetic_batch(N, arch["dim_x"])
    cfg = orc.default_cfg(kind, L=L)
    # oracle (bf16-emulating and plain fp32)
    om = orc.OracleModel(sd, cfg)
    om.iters = step
    tape = orc.Tape(seed=SEED_TAPE)
    lo_emu, g_emu = om.grads(batch, tape, emulate_bf16=True)
    lo_f32, g_f32 = om.grads(batch, orc.Tape(recorded=tape.log), emulate_bf16=False)
    plan = Plan(kind, L=L, max_batch=N, n_models=1, **arch)
    plan.set_gemm_impl(impl)
    plan.load_state_dict(sd)
    eps = eps_block_from_tape(plan, tape.log, batch.get("has_x2"), batch.get("has_y"), noisy=True)
    hp = plan.hparams(step=step, beta_pert=anneal_coef(step, 1, 0))
    t0 = time.time()
    losses = plan.grad_step(batch_fields(kind, batch), hp, eps=eps)
    torch.cuda.synchronize()
    dt = time.time() - t0
    row = losses[0].cpu()
    out = {}
    worst_loss = 0.0
    for i, k in enumerate(LOSS_KEYS):
        if k not in lo_emu or k == "MMD":
            continue
        ref = float(lo_emu[k])
        e = abs(float(row[i]) - ref) / (abs(ref) + 1e-12)
        e32 = abs(float(row[i]) - float(lo_f32[k])) / (abs(float(lo_f32[k])) + 1e-12)
        out[k] = (float(row[i]), ref, e, e32)
        worst_loss = max(worst_loss, e)
    gv = plan.tensor_views(plan.grads, 0)
    gerrs = {k: (rel_l2(gv[k], g_emu[k]), rel_l2(gv[k], g_f32[k])) for k in g_emu}
    worst_g = max(v[0] for v in gerrs.values())
    print("== %s/%s impl=%s step=%d: worst loss rel %.2e, worst grad relL2 vs emu %.2e (%.0f ms, %d launches)" %
          (case, kind, impl, step, worst_loss, worst_g, dt * 1e3, plan.launch_count()), flush=True)
    if verbose or worst_loss > 1e-3 or worst_g > 2e-2:
        for k, (v, r, e, e32) in out.items():
            print("     %-5s gpu % .6f emu % .6f rel %.2e | vs fp32 rel %.2e" % (k, v, r, e, e32))
        for k, (e, e32) in gerrs.items():
            flag = "  <<<" if e > 2e-2 else ""
            print("     grad %-45s vs emu %.2e vs fp32 %.2e%s" % (k, e, e32, flag))
    return worst_loss, worst_g


if __name__ == "__main__":
    cases = sys.argv[1].split(",") if len(sys.argv) > 1 else ["tiny", "deep", "readme"]
    impls = sys.argv[2].split(",") if len(sys.argv) > 2 else ["simt", "tc"]
    for case in cases:
        for kind in KINDS:
            for impl in impls:
                for step in (0, 1):
                    try:
                        run_case(case, kind, impl, step, verbose=(step == 1))
                    except Exception:
                        print("!! %s/%s impl=%s step=%d raised:" % (case, kind, impl, step))
                        traceback.print_exc()
                        sys.stdout.flush()
