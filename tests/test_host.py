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
    declared = sorted(set(re.findall(r"^(?:const char\*|const long long\*|int|long long)\s+(drvae_[a-z0-9_]+)\s*\(", hdr, re.M)))
    assert declared, "no declarations parsed"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), "libdrvae_b200.so does not export %s" % name
    assert sorted(_lib.EXPORTS) == declared


def test_struct_sizes_match_header_layout():
    # ints/floats only: natural alignment, no padding surprises between C and ctypes
    assert ctypes.sizeof(_lib.Arch) == 4 * (5 + 4 * (1 + _lib.MAX_HIDDEN) + 3)
    assert ctypes.sizeof(_lib.Batch) == 5 * 8 + 8 + 8 + 8  # 5 pointers, N (+pad), row_index, dataset_rows (+pad)
    assert ctypes.sizeof(_lib.Noise) == 24  # eps pointer, seed, row_offset
    assert ctypes.sizeof(_lib.HParams) == 4 * 17 + 4 * 8 + 4 + 8  # ... + padding + global_counts_dev pointer
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


def test_cli_flags_and_defaults_match_the_reference_drivers():
    """run_drvae.py:248-283 / run_pvae.py:238-266 / run_vfae.py:243-277."""
    from drvae_b200.cli import build_parser
    base = ["--modelid", "auto", "--datafile", "synthetic"]
    a = build_parser("drvae").parse_args(base)
    assert (a.batch_size, a.L, a.rseed, a.yloss_rate, a.noise_var, a.dim_z1, a.dim_z3) == (200, 1, 12345, 50., 0.01, 50, 50)
    assert (a.enc_z1, a.dec_x, a.enc_z3, a.dec_z1, a.enc_z2Fz1, a.class_y) == ([200, 200], [200, 200], [200], [200], [], [])
    assert a.semi_supervised and not a.stopearly and not a.train_w_noise and a.drug == "26" and a.data_mode == "strictC2C"
    # the README command line (workspace/example-cmd.sh)
    a = build_parser("drvae").parse_args(base + "--dim-z1 100 --dim-z3 100 --enc-z1 800 --dec-x 600 --enc-z3 200 --dec-z1 200 "
                                         "--L 2 --batch-size 150 --train-w-noise --stopearly --yloss-rate 1".split())
    assert (a.dim_z1, a.enc_z1, a.dec_x, a.L, a.batch_size, a.train_w_noise, a.stopearly, a.yloss_rate) == (100, [800], [600], 2, 150, True, True, 1.0)
    p = build_parser("pvae").parse_args(base)
    assert p.kl_z2_rate == 1. and not p.with_pairdata_test and not hasattr(p, "yloss_rate")
    v = build_parser("vfae").parse_args(base)
    assert (v.dim_z2, v.enc_z2, v.alldata) == (50, [200], False) and not hasattr(v, "pair_data_only")


def test_dataset_wrappers_follow_the_reference_layout():
    """wrap_in_DrVAEDataset (DrVAE.py:907-963): singletons first with a zero x2 and has_x2 = 0, then pairs."""
    import numpy as np
    from drvae_b200.training import wrap_in_DrVAEDataset, wrap_in_VFAEDataset
    sing = dict(x1=np.ones((3, 5), np.float32), y=np.array([0, 1, 0]), has_y=np.array([1, 0, 1]), s=np.zeros(3), cid=np.arange(3))
    pair = dict(x1=2 * np.ones((2, 5), np.float32), x2=3 * np.ones((2, 5), np.float32), y=np.array([1, 1]), has_y=np.array([1, 1]),
                s=np.zeros(2), cid=np.arange(3, 5))
    ds, d = wrap_in_DrVAEDataset(sing, pair)
    assert len(ds) == 5 and ds.has_x2.tolist() == [0, 0, 0, 1, 1] and ds.has_x2.dtype == torch.int32
    assert float(ds.x2[:3].abs().sum()) == 0 and float(ds.x2[3:].mean()) == 3
    x1, x2, s, y, hx, hy = ds[4]
    assert float(x1[0]) == 2 and int(hx) == 1 and int(hy) == 1
    ds2, _ = wrap_in_DrVAEDataset(sing, pair, remove_unlabeled=True)
    assert len(ds2) == 4
    ds3, _ = wrap_in_DrVAEDataset(sing, pair, concat="pair_only")
    assert len(ds3) == 2 and ds3.has_x2.tolist() == [1, 1]
    dv, _ = wrap_in_VFAEDataset(sing, pair, concat="both")
    assert len(dv) == 5 and len(dv[0]) == 4


def test_reconstruction_metrics_match_scipy_and_sklearn():
    """eval_x_reconstruction (DGMMixin.py:128-158) computed as batched reductions."""
    import numpy as np
    import scipy.stats
    import sklearn.metrics
    from drvae_b200.training import eval_x_reconstruction
    g = torch.Generator().manual_seed(0)
    x = torch.randn(17, 40, generator=g)
    r = x + 0.5 * torch.randn(17, 40, generator=g)
    sg = 0.3 + torch.rand(17, 40, generator=g)
    m = eval_x_reconstruction(x, r, sg)
    xn, rn = x.double().numpy(), r.double().numpy()
    assert abs(m["rmse"] - np.sqrt(((xn - rn) ** 2).mean())) < 1e-12
    assert abs(m["r2"] - sklearn.metrics.r2_score(xn, rn, multioutput="variance_weighted")) < 1e-10
    assert abs(m["pearr"] - np.mean([scipy.stats.pearsonr(xn[i], rn[i])[0] for i in range(17)])) < 1e-10
    sgn = sg.double().numpy()
    ll = (-0.5 * (np.log(2 * np.pi) + np.log(sgn ** 2) + (xn - rn) ** 2 / sgn ** 2)).sum(1).mean()
    assert abs(m["ll"] - ll) < 1e-9


def test_balanced_sampler_weights_and_epoch_indices():
    """compute_balanced_weights = utils.compute_balanced_weights(labels, unlabeled_data_ratio=None) (src/utils.py:292-327):
    every class weighs 1 / its size; device_epoch_indices = WeightedRandomSampler(weights, len(weights)) cut into
    DataLoader minibatches with drop_last when the dataset holds at least one batch (run_drvae.py:148-166)."""
    import numpy as np
    import torch
    from drvae_b200.training import compute_balanced_weights, device_epoch_indices
    labels = np.array([3, 3, 3, 7, 7, 9, 3, 9, 9, 9])
    w = compute_balanced_weights(labels)
    assert w.dtype == torch.float64
    want = {3: 1 / 4, 7: 1 / 2, 9: 1 / 4}
    assert np.allclose(w.numpy(), [want[int(l)] for l in labels])
    # classes are drawn with equal probability: class totals of the weights are equal
    for c in (3, 7, 9):
        assert abs(float(w[torch.from_numpy(labels == c)].sum()) - 1.0) < 1e-12
    g = torch.Generator().manual_seed(0)
    idx = device_epoch_indices(w, 4, g)
    assert idx.dtype == torch.int32 and tuple(idx.shape) == (2, 4) and int(idx.min()) >= 0 and int(idx.max()) < 10
    short = device_epoch_indices(w, 16, g)  # fewer rows than one batch: a single short batch, like DataLoader(drop_last=False)
    assert tuple(short.shape) == (1, 10)
    big = device_epoch_indices(compute_balanced_weights(np.arange(3000) % 3 == 0), 150, g)
    frac = float((big.long() % 3 == 0).float().mean())
    assert tuple(big.shape) == (20, 150) and 0.45 < frac < 0.55  # the 1/3 minority class fills half of the draws
