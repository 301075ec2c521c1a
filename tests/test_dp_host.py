"""Host logic of the data-parallel path (drvae_b200/dp.py) on CPU: world_size 2 over gloo, with the
oracle as the compute backend (the CUDA Plan is substituted by an object with the same four
methods).  Checks the property the GPU path relies on: shards that are given the GLOBAL
normalisers produce additive shares, so all-reduced gradients / losses and the replicated Adam
update equal the single-process full-batch step."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import ARCH, orc  # noqa: E402

from drvae_b200 import dp  # noqa: E402
from drvae_b200.init import init_state_dict  # noqa: E402

KIND, CASE, N, L = "drvae", "tiny", 26, 2


class OracleBackend:
    """Same interface as dp.PlanBackend, computed by the CPU oracle."""

    def __init__(self, sd, cfg):
        self.om = orc.OracleModel(sd, cfg)
        self.names = list(self.om.sd.keys())
        self.sizes = [self.om.sd[k].numel() for k in self.names]
        self.flat = torch.zeros(sum(self.sizes))

    def flat_grads(self):
        return self.flat

    def buckets(self):
        # two buckets, deliberately not aligned with tensors
        h = self.flat.numel() // 3
        return [(h, self.flat.numel() - h), (0, h)]

    def grad_step(self, batch, hp_kwargs, counts, step, eps=None, seed=0, row_offset=0):
        self.om.cfg["counts"] = dict(N=counts[0], Np=counts[1], Nlab=counts[2])
        self.om.iters = step
        losses, g = self.om.grads(batch, eps)
        self.flat.copy_(torch.cat([g[k].reshape(-1) for k in self.names]))
        vec = torch.zeros(8)
        for i, k in enumerate(("RECL", "KLD", "PERT", "YL", "MMD", "ELBO", "CMPL")):
            if k in losses:
                vec[i] = float(losses[k])
        return vec, [None] * len(self.buckets())

    def adam_step(self):
        off = 0
        for k, n in zip(self.names, self.sizes):
            self.om.sd[k].grad = self.flat[off:off + n].view_as(self.om.sd[k]).clone()
            off += n
        self.om.opt.step()


class RowTape:
    """ε keyed by the GLOBAL row (what drvae_noise_t.row_offset does on the GPU): every draw of shape
    (n_group, d) takes the rows of a fixed per-draw-kind table, so shards see the noise of the
    unsharded run.  The oracle calls draw() once per (group, draw) in reference order."""

    def __init__(self, global_rows_of_group, seed=5):
        self.groups = global_rows_of_group  # list of index tensors, one per draw call in order
        self.g = torch.Generator().manual_seed(seed)
        self.tables = {}
        self.calls = 0
        self.log = []

    def draw(self, *shape):
        rows = self.groups[self.calls]
        key = (self.calls_key(), shape[1])
        self.calls += 1
        full = self.tables[key]
        out = full[rows]
        assert out.shape == tuple(shape), (out.shape, shape)
        self.log.append(out)
        return out.clone()

    def calls_key(self):
        return self.keys[self.calls]


def make_row_tape(batch, rows_global, arch, seed):
    """Build a RowTape for a (sub)batch whose local row i is global row rows_global[i]."""
    from drvae_b200.noise import group_indices
    Z, Z3, X, Y = arch["dim_z1"], arch["dim_z3"], arch["dim_x"], arch["dim_y"]
    NG = 64
    g = torch.Generator().manual_seed(seed)
    tables = {}
    kinds = ["x1", "x2"] + ["z1_%d" % l for l in range(L)] + ["z2_%d" % l for l in range(L)] + \
            ["z2f_%d" % l for l in range(L)] + ["z3_%d_%d" % (l, j) for l in range(L) for j in range(Y)]
    for k in kinds:
        d = X if k in ("x1", "x2") else (Z3 if k.startswith("z3") else Z)
        tables[(k, d)] = torch.randn(NG, d, generator=g)
    groups, keys = [], []
    for idx, pair, lab in group_indices(KIND, batch["has_x2"], batch["has_y"]):
        if len(idx) == 0:
            continue
        gi = rows_global[idx]
        groups.append(gi), keys.append("x1")
        if pair:
            groups.append(gi), keys.append("x2")
        for l in range(L):
            groups.append(gi), keys.append("z1_%d" % l)
            if pair:
                groups.append(gi), keys.append("z2_%d" % l)
            groups.append(gi), keys.append("z2f_%d" % l)
            for j in range(1 if lab else Y):
                groups.append(gi), keys.append("z3_%d_%d" % (l, j))
    t = RowTape(groups)
    t.keys = keys
    t.tables = tables
    return t


def sub_batch(batch, lo, hi):
    return {k: v[lo:hi].clone() for k, v in batch.items()}


def _worker(rank, world, port, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    arch = ARCH[CASE]
    sd = init_state_dict(KIND, seed=123, **arch)
    batch = orc.synthetic_batch(N, arch["dim_x"])
    lo, hi = dp.shard_rows(N, world, rank)
    be = OracleBackend(sd, orc.default_cfg(KIND, L=L))
    runner = dp.DataParallel(be)
    res = []
    for it in range(2):
        shard = sub_batch(batch, lo, hi)
        tape = make_row_tape(shard, torch.arange(lo, hi), arch, seed=100 + it)
        losses = runner.step(shard, eps=tape)
        res.append(losses.clone())
    if rank == 0:
        # numpy: pickled by value (torch tensors would be shared through file descriptors of a process that exits)
        out_q.put(([r.numpy() for r in res], {k: v.detach().numpy().copy() for k, v in be.om.sd.items()}))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_rows_partition():
    for n, w in ((8192, 8), (150, 4), (7, 3), (3, 4)):
        spans = [dp.shard_rows(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for a, b in zip(spans, spans[1:]):
            assert a[1] == b[0]
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_step_equals_full_batch_step():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res, sd_dp = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference: the full batch through the same oracle
    arch = ARCH[CASE]
    sd = init_state_dict(KIND, seed=123, **arch)
    batch = orc.synthetic_batch(N, arch["dim_x"])
    om = orc.OracleModel(sd, orc.default_cfg(KIND, L=L))
    for it in range(2):
        tape = make_row_tape(batch, torch.arange(N), arch, seed=100 + it)
        want = om.step(batch, tape)
        for i, k in enumerate(("RECL", "KLD", "PERT", "YL", "MMD", "ELBO", "CMPL")):
            assert abs(float(res[it][i]) - float(want[k])) <= 2e-5 * abs(float(want[k])) + 1e-6, (it, k)
    for k, v in om.state_dict().items():
        assert torch.allclose(torch.from_numpy(sd_dp[k]), v, rtol=1e-4, atol=2e-6), k
