"""Shared fixtures for the parity tests: architectures, synthetic batches, oracle runs."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import drvae_oracle as orc  # noqa: E402  (tests are allowed to use the oracle)

GOLDEN = os.path.join(ROOT, "tests", "golden")
ARCH = dict(
    readme=dict(dim_x=978, dim_y=2, dim_z1=100, dim_z3=100, enc_z1=[800], dec_x=[600], enc_z3=[200], dec_z1=[200]),
    tiny=dict(dim_x=40, dim_y=2, dim_z1=12, dim_z3=10, enc_z1=[24], dec_x=[28], enc_z3=[20], dec_z1=[18]),
    deep=dict(dim_x=40, dim_y=2, dim_z1=12, dim_z3=10, enc_z1=[24, 20], dec_x=[16, 28], enc_z3=[20, 12], dec_z1=[18, 14]),
)
ARCH["tiny_wn"] = ARCH["tiny"]  # same dims, built from layers.WeightNormLinear (golden recorded with wn=True)
NROWS = dict(readme=150, tiny=24, deep=24, tiny_wn=24)
KINDS = ("drvae", "pvae", "vfae")
SEED_MODEL, SEED_TAPE, L = 123, 777, 2


def golden(kind, case):
    return np.load(os.path.join(GOLDEN, "%s_%s.npz" % (kind, case)))


def batch_fields(kind, batch):
    """Subset of the synthetic batch a model kind consumes (as keyword arguments of Plan calls)."""
    b = dict(x1=batch["x1"])
    if kind in ("drvae", "pvae"):
        b.update(x2=batch["x2"], has_x2=batch["has_x2"])
    if kind in ("drvae", "vfae"):
        b.update(y=batch["y"], has_y=batch["has_y"])
    return b


def oracle_batch(kind, batch):
    return batch


def rel_err(a, b):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def rel_l2(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()
