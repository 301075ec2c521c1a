"""Run the UNMODIFIED reference (/root/reference/src) in this container.  TEST INFRASTRUCTURE ONLY.

/root/reference exists only in the build container (never on the GPU box), so this module is
used by oracle/make_golden.py to (a) pin oracle/drvae_oracle.py against the reference itself and
(b) write the golden fixtures under tests/golden/.  Nothing is copied from the reference; its
modules are imported from where they lie.

Compatibility shim — the reference targets PyTorch 0.3.1:
  1. `h5py` is absent: a stub module is registered (only utils.load/save_to_HDF use it);
  2. `blocks.one_hot` relies on `.data.unsqueeze_` reshaping the caller's tensor
     (src/blocks.py:84) which no longer happens: replaced by an equivalent scatter;
  3. `PVAE.__init__` reads `self.prior_y`, which is never set (src/PVAE.py:77): class attribute;
  4. the run_*.py drivers are never imported (version gate, `git log` at import);
  5. `model.add_noise` is set by `fit()` (src/DrVAE.py:769): set explicitly here.
ε injection: `torch.Tensor.normal_` is patched for the duration of a step to draw from a seeded
CPU generator and to record every draw (the "ε tape").
"""
import contextlib
import io
import os
import sys
import types
import warnings

import torch

REF_SRC = "/root/reference/src"


def available():
    return os.path.isdir(REF_SRC)


_mods = None


def load_reference():
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("reference sources not present at %s" % REF_SRC)
    if "h5py" not in sys.modules:
        sys.modules["h5py"] = types.ModuleType("h5py")
    sys.path.insert(0, REF_SRC)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import blocks as blk  # noqa
        import DrVAE as drvae_mod  # noqa
        import PVAE as pvae_mod  # noqa
        import VFAE as vfae_mod  # noqa

    def one_hot(y, max_dim):
        if y is not None and len(y) > 0:
            out = torch.zeros(y.size(0), max_dim)
            out.scatter_(1, y.detach().long().view(-1, 1), 1)
            return out
        return None

    blk.one_hot = one_hot
    pvae_mod.PVAE.prior_y = None
    _mods = dict(blk=blk, DrVAE=drvae_mod.DrVAE, PVAE=pvae_mod.PVAE, VFAE=vfae_mod.VFAE)
    return _mods


# constructor arguments of the run_*.py drivers (run_drvae.py:173-185, run_pvae.py:169-179,
# run_vfae.py:169-180) with the architecture left as parameters
def build_reference_model(kind, arch, seed=123, L=2, noise=0.01, yloss_rate=1.0, batch_size=150, weight_norm=False):
    """weight_norm=True builds the blocks from layers.WeightNormLinear.  The reference constructors
    hard-code `self.wn = False` (DrVAE.py:80) before `_build_blocks`, so the flag is flipped by a
    wrapper around `_build_blocks` for the duration of the construction; no reference code changes."""
    m = load_reference()
    cls = m[{"drvae": "DrVAE", "pvae": "PVAE", "vfae": "VFAE"}[kind]]
    orig_build = cls._build_blocks
    if weight_norm:
        def _build_wn(self):
            self.wn = True
            return orig_build(self)
        cls._build_blocks = _build_wn
    try:
        return _build_reference_model(m, kind, arch, seed, L, noise, yloss_rate, batch_size)
    finally:
        cls._build_blocks = orig_build


def _build_reference_model(m, kind, arch, seed, L, noise, yloss_rate, batch_size):
    common = dict(type_rec="diag_gaussian", epochs=1, batch_size=batch_size, nonlinearity="elu",
                  learning_rate=0.0005, optim_alg="adam", L=L, weight_decay=0.05, dropout_rate=0.,
                  input_x_dropout=0., add_noise_var=noise, use_MMD=False, kernel_MMD="rbf_fourier",
                  mmd_rate=1., use_s=False, random_seed=seed, log_txt=None)
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if kind == "drvae":
            model = m["DrVAE"](dim_x=arch["dim_x"], dim_s=arch.get("dim_s", 1), dim_y=arch["dim_y"],
                               dim_h_en_z1=arch["enc_z1"], dim_h_de_z1=arch["dec_z1"], dim_h_en_z2Fz1=[],
                               dim_h_en_z3=arch["enc_z3"], dim_h_de_x=arch["dec_x"], dim_h_clf=[],
                               dim_z1=arch["dim_z1"], dim_z3=arch["dim_z3"], clf_z1z2=True, type_y="discrete",
                               prior_y="uniform", clf_1sig=False, yloss_rate=yloss_rate, anneal_yloss_offset=1,
                               kl_qz2pz2_rate=1., pertloss_rate=0.05, anneal_perturb_rate_itermax=1,
                               anneal_perturb_rate_offset=0, **common)
        elif kind == "pvae":
            model = m["PVAE"](dim_x=arch["dim_x"], dim_s=arch.get("dim_s", 1), dim_y=arch["dim_y"],
                              dim_h_en_z1=arch["enc_z1"], dim_h_en_z2Fz1=[], dim_h_de_x=arch["dec_x"],
                              dim_z1=arch["dim_z1"], kl_qz2pz2_rate=1., pertloss_rate=0.05,
                              anneal_perturb_rate_itermax=1, anneal_perturb_rate_offset=0, **common)
        elif kind == "vfae":
            model = m["VFAE"](dim_x=arch["dim_x"], dim_s=arch.get("dim_s", 1), dim_y=arch["dim_y"],
                              dim_h_en_z1=arch["enc_z1"], dim_h_de_z1=arch["dec_z1"], dim_h_en_z2=arch["enc_z2"],
                              dim_h_de_x=arch["dec_x"], dim_h_clf=[], dim_z1=arch["dim_z1"], dim_z2=arch["dim_z2"],
                              type_y="discrete", prior_y="uniform", semi_supervised=True, clf_1sig=False,
                              yloss_rate=yloss_rate, anneal_yloss_offset=1, **common)
        else:
            raise ValueError(kind)
    model.add_noise = noise > 0
    return model


class TapeRecorder:
    """Patches torch.Tensor.normal_ so that every draw comes from a seeded CPU generator and is
    recorded in order."""

    def __init__(self, seed):
        self.gen = torch.Generator().manual_seed(seed)
        self.draws = []

    def __enter__(self):
        self._orig = torch.Tensor.normal_
        rec = self

        def normal_(t, mean=0.0, std=1.0, generator=None):
            e = torch.randn(*t.shape, generator=rec.gen)
            rec.draws.append(e)
            with torch.no_grad():
                t.copy_(e * std + mean)
            return t

        torch.Tensor.normal_ = normal_
        return self

    def __exit__(self, *exc):
        torch.Tensor.normal_ = self._orig


def batch_kwargs(kind, batch):
    """The keyword batch each model's loss_function takes (DrVAE.py:545, PVAE.py:411, VFAE.py:403)."""
    b = {k: v.clone() for k, v in batch.items()}
    if kind == "drvae":
        return dict(x1=b["x1"], x2=b["x2"], s=b["s"], y=b["y"], has_x2=b["has_x2"], has_y=b["has_y"])
    if kind == "pvae":
        return dict(x1=b["x1"], x2=b["x2"], s=b["s"], has_x2=b["has_x2"])
    return dict(x1=b["x1"], s=b["s"], y=b["y"], has_y=b["has_y"])


def reference_step(model, kind, batch, tape_seed, train=True):
    """One run_on_batch of the reference; returns (losses, tape draws)."""
    with TapeRecorder(tape_seed) as rec, warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        losses = model.run_on_batch(train_mode=train, **batch_kwargs(kind, batch))
    out = {}
    for k, v in losses.items():
        out[k] = float(v.detach().reshape(-1)[0]) if torch.is_tensor(v) else float(v)
    return out, rec.draws


def reference_grads(model):
    return {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p))
            for k, p in model.named_parameters()}


def reference_forward(model, x1):
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        with torch.no_grad():
            return model.forward(x1=x1.clone(), s=[])
