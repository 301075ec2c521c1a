"""Multi-GPU parity of the data-parallel step (BASELINE configs[3]; round-1 VERDICT item 4 / ADVICE): needs >= 2 GPUs."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(backend, arch, N, steps, world=2, port=29711):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dp_worker.py"), backend, arch, str(N), str(steps)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    res = [json.loads(l[len("DPRESULT "):]) for l in out.stdout.splitlines() if l.startswith("DPRESULT ")]
    same = [l for l in out.stdout.splitlines() if l.startswith("DPSAME ")]
    assert len(res) == 1 and same == ["DPSAME 1"], out.stdout[-2000:]
    return res[0]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("backend,arch,N", [("peer", "tiny", 96), ("peer", "readme", 4096), ("nccl", "tiny", 96)])
def test_two_rank_data_parallel_step_equals_single_gpu_step(backend, arch, N):
    """Rows sharded over 2 ranks (gradients summed over NVLink peer memory inside the optimizer kernel, or NCCL) follow
    the single-GPU trajectory of the unsharded minibatch: losses equal to 1e-5, parameters to rounding, and the ranks'
    parameters are bit-identical to each other.  6 steps: eager, then the captured graph."""
    r = _run(backend, arch, N, 6)
    print(r)
    assert r["worst_rel_loss_err"] < 2e-5 if arch == "tiny" else r["worst_rel_loss_err"] < 1e-4, r
    assert r["param_rel_l2"] < 1e-4, r
    if backend == "peer":
        assert r["graphs"] == 1
