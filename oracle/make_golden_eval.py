"""Golden vectors of the evaluation metrics from the REFERENCE ITSELF (build container only; needs /root/reference and
scipy / sklearn):  python -m oracle.make_golden_eval  ->  tests/golden/eval_metrics.npz

The unmodified DeepGenerativeModelMixin.eval_x_reconstruction (src/DGMMixin.py:128-158) of a reference DrVAE instance
is called on seeded inputs for rmse / r2 / pearr.  Its log-likelihood line ends in `.data.numpy()[0]`, which indexes a
0-dim array under torch >= 0.4, so `ll` is taken from the very expression of that line —
`self.decoder_x.logp_perx(x, x_rec, x_rec_logvar).mean()` — evaluated on the same reference module.  Cases: all rows,
and the rows selected by a mask (the reference indexes the paired rows before the call, src/DrVAE.py:693-700).
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402

TINY = dict(dim_x=40, dim_y=2, dim_z1=12, dim_z3=10, dim_z2=10, enc_z1=[24], dec_x=[28], enc_z3=[20], enc_z2=[20], dec_z1=[18])


def main():
    model = rh.build_reference_model("drvae", TINY, seed=123, L=2)
    out = {}
    for name, (N, X, seed) in {"small": (37, 50, 1), "l1000": (150, 978, 2)}.items():
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(N, X, generator=g)
        rec = x + 0.5 * torch.randn(N, X, generator=g)
        sg = torch.rand(N, X, generator=g) * 0.9 + 0.1
        mask = (torch.arange(N) % 3 != 1).int()
        out[name + "/shape_seed"] = np.array([N, X, seed])
        if name == "small":  # inputs stored; the large case is regenerated from its seed by the test (same torch build)
            out[name + "/x"], out[name + "/rec"], out[name + "/sg"], out[name + "/mask"] = x.numpy(), rec.numpy(), sg.numpy(), mask.numpy()
        for tag, idx in (("all", torch.arange(N)), ("masked", torch.nonzero(mask).view(-1))):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                pd = model.eval_x_reconstruction(x[idx], rec[idx], None)
                ll = float(model.decoder_x.logp_perx(x[idx], rec[idx], sg[idx]).mean())
            out["%s/%s" % (name, tag)] = np.array([pd["rmse"], pd["r2"], pd["pearr"], ll], dtype=np.float64)
            print(name, tag, out["%s/%s" % (name, tag)])
    path = os.path.join(ROOT, "tests", "golden", "eval_metrics.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
