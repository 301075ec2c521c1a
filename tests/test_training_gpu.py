"""GPU: the training-loop surface around the step (reference fit / evaluate_performance / run_*.py drivers)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ("drvae", "pvae", "vfae"))
def test_cli_trains_snapshots_reloads_and_reports(kind, tmp_path):
    from drvae_b200 import cli
    argv = [kind, "--modelid", "auto", "--datafile", "synthetic:400", "--outdir", str(tmp_path), "--epochs", "3", "--batch-size", "40",
            "--L", "2", "--dim-z1", "12", "--enc-z1", "24", "--dec-x", "28", "--train-w-noise", "--stopearly", "--rseed", "7"]
    if kind == "drvae":
        argv += ["--dim-z3", "10", "--enc-z3", "20", "--dec-z1", "18", "--yloss-rate", "1"]
    if kind == "vfae":
        argv += ["--dim-z2", "10", "--enc-z2", "20", "--dec-z1", "18", "--yloss-rate", "1"]
    res = cli.main(argv)
    for split in ("train", "valid", "test"):
        r = res[split]
        assert np.isfinite(r["x1_rmse"]) and np.isfinite(r["x1_pearr"]) and r["losses"] is not None
        assert np.isfinite(r["losses"]["CMPL"])
        if kind != "pvae":
            assert 0.0 <= r["y_acc"] <= 1.0
    files = os.listdir(os.path.join(str(tmp_path), "results"))
    assert len(files) == 1
    with open(os.path.join(str(tmp_path), "results", files[0])) as f:
        assert "26" in json.load(f)
    assert len(os.listdir(os.path.join(str(tmp_path), "models"))) == 1


def test_fit_reduces_the_training_loss_and_learns_the_labels():
    """A few epochs on separable synthetic data: CMPL falls and the classifier beats chance."""
    from drvae_b200 import DrVAE, wrap_in_DrVAEDataset
    from drvae_b200.cli import synthetic_data, split
    sing, pair = synthetic_data(600, dim_x=60, seed=1)
    tr_s, va_s, _ = split(sing, 1)
    tr_p, va_p, _ = split(pair, 1)
    train_ds, _ = wrap_in_DrVAEDataset(tr_s, tr_p)
    valid_ds, _ = wrap_in_DrVAEDataset(va_s, va_p)
    model = DrVAE(dim_x=60, dim_s=1, dim_y=2, dim_h_en_z1=[32], dim_h_de_z1=[16], dim_h_en_z2Fz1=[], dim_h_en_z3=[16], dim_h_de_x=[32],
                  dim_h_clf=[], dim_z1=8, dim_z3=6, type_rec="diag_gaussian", epochs=30, batch_size=60, nonlinearity="elu",
                  learning_rate=5e-3, L=1, weight_decay=0.0, add_noise_var=0.01, yloss_rate=50., use_MMD=False, pertloss_rate=0.05,
                  random_seed=3)
    before, _ = model.evaluate_performance_on_dataset(valid_ds)
    loader = torch.utils.data.DataLoader(train_ds, batch_size=60, shuffle=True, drop_last=True)
    vloader = torch.utils.data.DataLoader(valid_ds, batch_size=60)
    model.fit(loader, vloader, add_noise=True)
    after, _ = model.evaluate_performance_on_dataset(valid_ds)
    assert float(after["losses"]["RECL"]) > float(before["losses"]["RECL"]), (float(before["losses"]["RECL"]), float(after["losses"]["RECL"]))
    assert after["y_auroc"] > 0.65 and after["y_auroc"] > before["y_auroc"], (before["y_auroc"], after["y_auroc"])
    assert model.finished_training_iters == 30 * len(loader)


@pytest.mark.gpu
def test_ensemble_checkpoints_interchange_with_the_reference_format(tmp_path):
    """Batched save / load of an ensemble's state_dicts (reference keys, order, shapes: one .pth per member, loadable by
    DGMMixin.load_params_from_file), and a resume file (parameters + Adam moments + step) that continues bit-exactly."""
    import numpy as np
    from helpers import ARCH, L, batch_fields, golden, orc
    from drvae_b200 import checkpoint as ckpt
    from drvae_b200.init import init_state_dict
    from drvae_b200.plan import Plan, anneal_coef
    kind, arch, N, E = "drvae", ARCH["tiny"], 24, 5
    sds = [init_state_dict(kind, seed=50 + m, **arch) for m in range(E)]
    batches = [batch_fields(kind, orc.synthetic_batch(N, arch["dim_x"], seed=m)) for m in range(E)]
    big = {k: torch.stack([b[k] for b in batches]).contiguous().cuda() for k in batches[0]}

    def fresh():
        p = Plan(kind, L=L, max_batch=N, n_models=E, **arch)
        ckpt.load_ensemble(p, sds)
        return p

    a = fresh()
    for it in range(3):
        a.train_step(big, a.hparams(step=it, beta_pert=anneal_coef(it, 1, 0)), seed=2)
    paths = ckpt.save_ensemble(a, str(tmp_path / "member{:02d}.pth"))
    ref_keys = [k[3:] for k in golden(kind, "tiny").files if k.startswith("sd/")]
    for m, path in enumerate(paths):
        sd = torch.load(path)
        assert list(sd.keys()) == ref_keys  # the reference's state_dict keys in the reference's order
        for k, v in sd.items():
            assert v.is_contiguous() and tuple(v.shape) == tuple(sds[m][k].shape)
            assert torch.equal(v, a.tensor_views(a.params, m)[k].cpu())
    b = Plan(kind, L=L, max_batch=N, n_models=E, **arch)
    ckpt.load_ensemble(b, paths)
    assert torch.equal(a.params, b.params)
    # resume: 3 steps + save + 2 steps  ==  load + 2 steps
    ckpt.save_resume(a, str(tmp_path / "resume.pt"), step=3)
    c = Plan(kind, L=L, max_batch=N, n_models=E, **arch)
    step = ckpt.load_resume(c, str(tmp_path / "resume.pt"))
    assert step == 3
    for it in range(3, 5):
        la = a.train_step(big, a.hparams(step=it, beta_pert=1.0), seed=2).cpu().clone()
        lc = c.train_step(big, c.hparams(step=it, beta_pert=1.0), seed=2).cpu().clone()
        assert torch.equal(la, lc)
    assert torch.equal(a.params, c.params) and torch.equal(a.adam_v, c.adam_v)
    with pytest.raises(ValueError):
        ckpt.load_ensemble(b, paths[:-1])


@pytest.mark.gpu
def test_reconstruction_metrics_kernel_matches_the_reference():
    """drvae_eval_x_reconstruction (fp64 device reductions) against the outputs of the reference's own
    eval_x_reconstruction / logp_perx (tests/golden/eval_metrics.npz, oracle/make_golden_eval.py): all rows and a row
    mask; stored inputs and inputs regenerated from the fixture's seed at the L1000 shape."""
    import numpy as np
    from helpers import GOLDEN as GOLDEN_DIR
    from drvae_b200.training import eval_x_reconstruction
    g = np.load(os.path.join(GOLDEN_DIR, "eval_metrics.npz"))
    for name in ("small", "l1000"):
        N, X, seed = (int(v) for v in g[name + "/shape_seed"])
        gen = torch.Generator().manual_seed(seed)
        x = torch.randn(N, X, generator=gen)
        rec = x + 0.5 * torch.randn(N, X, generator=gen)
        sg = torch.rand(N, X, generator=gen) * 0.9 + 0.1
        mask = (torch.arange(N) % 3 != 1).int()
        if name == "small":
            assert np.array_equal(x.numpy(), g["small/x"]) and np.array_equal(sg.numpy(), g["small/sg"])
        for tag, mk in (("all", None), ("masked", mask)):
            want = g["%s/%s" % (name, tag)]
            got = eval_x_reconstruction(x.cuda(), rec.cuda(), sg.cuda(), mask=mk.cuda() if mk is not None else None)
            # rmse / r2 / pearr: the reference works in float64 on the same float32 data (1e-9); its log-likelihood is
            # evaluated in float32 (blocks.py:233-234), the kernel's in float64 (1e-6 = the reference's own rounding)
            tol = dict(rmse=1e-9, r2=1e-9, pearr=1e-9, ll=1e-6)
            for k, w in zip(("rmse", "r2", "pearr", "ll"), want):
                assert abs(got[k] - w) <= tol[k] * max(1.0, abs(w)), (name, tag, k, got[k], w)
            host = eval_x_reconstruction(x, rec, sg, mask=mk)  # the torch formulas used by the CPU tests
            for k, w in zip(("rmse", "r2", "pearr", "ll"), want):
                assert abs(host[k] - w) <= tol[k] * max(1.0, abs(w)), (name, tag, k, host[k], w)
    no_sigma = eval_x_reconstruction(x.cuda(), rec.cuda(), None)
    assert no_sigma["ll"] != no_sigma["ll"] and abs(no_sigma["rmse"] - g["l1000/all"][0]) < 1e-9


@pytest.mark.gpu
def test_device_resident_fit_trains_without_host_batches():
    """fit_resident: dataset on the GPU, one device multinomial draw per epoch (the reference's balanced
    WeightedRandomSampler), steps reading the dataset through row indices and replaying ONE CUDA graph.  The loss falls,
    the classifier learns, the step counter follows the reference's bookkeeping, and the balanced weights equal
    utils.compute_balanced_weights."""
    from drvae_b200 import DrVAE, wrap_in_DrVAEDataset
    from drvae_b200.cli import synthetic_data, split
    from drvae_b200.training import compute_balanced_weights, device_epoch_indices, fit_resident
    sing, pair = synthetic_data(600, dim_x=60, seed=1)
    tr_s, va_s, _ = split(sing, 1)
    tr_p, va_p, _ = split(pair, 1)
    train_ds, _ = wrap_in_DrVAEDataset(tr_s, tr_p)
    valid_ds, _ = wrap_in_DrVAEDataset(va_s, va_p)
    model = DrVAE(dim_x=60, dim_s=1, dim_y=2, dim_h_en_z1=[32], dim_h_de_z1=[16], dim_h_en_z2Fz1=[], dim_h_en_z3=[16], dim_h_de_x=[32],
                  dim_h_clf=[], dim_z1=8, dim_z3=6, type_rec="diag_gaussian", epochs=30, batch_size=60, nonlinearity="elu",
                  learning_rate=5e-3, L=1, weight_decay=0.0, add_noise_var=0.01, yloss_rate=50., use_MMD=False, pertloss_rate=0.05,
                  random_seed=3)
    n = len(train_ds)
    labels = np.arange(n) % 7  # stand-in for the cell line ids the reference balances by (run_drvae.py:148-151)
    w = compute_balanced_weights(labels)
    counts = np.bincount(labels)
    assert np.allclose(w.numpy(), 1.0 / counts[labels])
    idx = device_epoch_indices(w.cuda(), 60, torch.Generator(device="cuda").manual_seed(0))
    assert idx.dtype == torch.int32 and tuple(idx.shape) == (n // 60, 60) and int(idx.min()) >= 0 and int(idx.max()) < n
    before, _ = model.evaluate_performance_on_dataset(valid_ds)
    replays0 = model.plan.graph_replays()
    hist = fit_resident(model, train_ds, batch_size=60, epochs=30, weights=w, add_noise=True, seed=5)
    after, _ = model.evaluate_performance_on_dataset(valid_ds)
    assert len(hist) == 30 and all(np.isfinite(h["CMPL"]) for h in hist)
    assert hist[-1]["CMPL"] < hist[0]["CMPL"], (hist[0]["CMPL"], hist[-1]["CMPL"])
    assert float(after["losses"]["RECL"]) > float(before["losses"]["RECL"])
    assert after["y_auroc"] > 0.65 and after["y_auroc"] > before["y_auroc"], (before["y_auroc"], after["y_auroc"])
    assert model.finished_training_iters == 30 * (n // 60)
    assert model.plan.graph_replays() - replays0 >= 30 * (n // 60) - 4  # every step after the capture is a replay
