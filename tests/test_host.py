"""CPU: host-side logic and the C-ABI surface (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from helpers import ARCH, KINDS, L, NROWS, ROOT, SEED_MODEL, golden, orc

from drvae_b200 import _lib
from drvae_b200.init import init_state_dict
from drvae_b200.noise import group_indices, tape_shapes
from drvae_b200.plan import anneal_coef


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "drvae_b200.h")).read()
    declared = sorted(set(re.findall(r"^(?:const char\*|int|long long)\s+(drvae_[a-z0-9_]+)\s*\(", hdr, re.M)))
    assert declared, "no declarations parsed"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), "libdrvae_b200.so does not export %s" % name
    assert sorted(_lib.EXPORTS) == declared


def test_struct_sizes_match_header_layout():
    # ints/floats only: natural alignment, no padding surprises between C and ctypes
    assert ctypes.sizeof(_lib.Arch) == 4 * (5 + 4 * (1 + _lib.MAX_HIDDEN) + 3)
    assert ctypes.sizeof(_lib.Batch) == 5 * 8 + 8
    assert ctypes.sizeof(_lib.Noise) == 16
    assert ctypes.sizeof(_lib.HParams) == 4 * 17 + 4 * 8
    assert ctypes.sizeof(_lib.EpsLayout) == 7 * 8
    assert ctypes.sizeof(_lib.InferOut) == 10 * 8


def test_plan_create_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from drvae_b200.plan import Plan
    with pytest.raises(RuntimeError):
        Plan("drvae", L=1, max_batch=8, **ARCH["tiny"])


def test_last_error_is_reported():
    lib = _lib.load()
    a = _lib.Arch()
    a.kind = 7
    h = ctypes.c_void_p()
    assert lib.drvae_plan_create(ctypes.byref(a), 1, ctypes.byref(h)) != 0
    assert b"kind" in lib.drvae_last_error()


@pytest.mark.parametrize("case", ("tiny", "deep", "readme"))
@pytest.mark.parametrize("kind", KINDS)
def test_init_reproduces_reference_weights(kind, case):
    g = golden(kind, case)
    sd = init_state_dict(kind, seed=SEED_MODEL, **ARCH[case])
    names = [k[6:] for k in g.files if k.startswith("sdsum/")]
    assert sorted(names) == sorted(sd.keys())
    for k in names:
        a = sd[k].numpy()
        s = g["sdsum/" + k]
        assert abs(a.sum(dtype=np.float64) - s[0]) < 1e-9 and abs(np.abs(a).sum(dtype=np.float64) - s[1]) < 1e-9, k


def test_anneal_coef_matches_reference_rule():
    # DGMMixin.py:77-89 with iter_max=1, offset=0: 0.01 on the first step, 1.0 afterwards
    assert anneal_coef(0, 1, 0) == 0.01
    assert anneal_coef(1, 1, 0) == 1.0
    assert anneal_coef(5, 1000, 0) == pytest.approx(0.015)
    assert anneal_coef(3, 1000, 10) == 0.01


@pytest.mark.parametrize("kind", KINDS)
def test_tape_shapes_follow_the_oracle_draw_order(kind):
    arch, N = ARCH["tiny"], NROWS["tiny"]
    sd = init_state_dict(kind, seed=SEED_MODEL, **arch)
    batch = orc.synthetic_batch(N, arch["dim_x"])
    om = orc.OracleModel(sd, orc.default_cfg(kind, L=L))
    for train in (True, False):
        tape = orc.Tape(seed=1)
        om.loss(batch, tape, train=train)
        want = tape_shapes(kind, batch["has_x2"], batch["has_y"], arch["dim_x"], arch["dim_z1"], arch["dim_z3"],
                           arch["dim_y"], L, noisy=train)
        assert [tuple(t.shape) for t in tape.log] == want


def test_group_indices_partition_rows():
    b = orc.synthetic_batch(31, 5)
    for kind in KINDS:
        idx = torch.cat([g[0] for g in group_indices(kind, b["has_x2"], b["has_y"])])
        assert sorted(idx.tolist()) == list(range(31))
