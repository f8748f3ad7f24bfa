"""GPU side of SURVEY §8f N4: a Trainer's state survives save -> restore (parameters, Adam moments, step, rng) so that the
resumed run continues the same trajectory; a reference-style run directory drives the evaluator; the host input pipeline
delivers the reference's batch dicts."""
import numpy as np
import pytest
import torch

from tests.util import conditioned_params, spec_of

pytestmark = pytest.mark.gpu


def _trainer(precision="fp32"):
    from posterior_matching_b200 import PosteriorMatchingVAE, Trainer, pm_vae_config
    cfg = pm_vae_config("gas")
    m = PosteriorMatchingVAE.from_config(cfg.model, precision=precision)
    m.load_params(conditioned_params(spec_of("gas")))
    tr = Trainer(cfg, seed=2, precision=precision, model=m)
    tr.step = 5000
    return cfg, tr


def test_save_restore_resumes_the_same_trajectory(tmp_path):
    from posterior_matching_b200 import checkpoint as ck
    xs = [torch.randn(256, 8, device="cuda", generator=torch.Generator(device="cuda").manual_seed(i)) for i in range(6)]
    cfg, a = _trainer()
    for x in xs[:3]:
        a.train_step(x)
    path = str(tmp_path / "run" / "train_state.pkl")
    ck.save_train_state(path, a)
    ck.save_model_config(str(tmp_path / "run"), cfg.model)
    calls = a.mask_generator._calls
    for x in xs[3:]:
        a.train_step(x)
    want = a.metrics()
    _, b = _trainer()
    ck.restore_trainer(b, path)
    b.mask_generator._calls = calls            # the mask stream position belongs to the input pipeline, not to TrainState
    assert b.step == 5003
    for x in xs[3:]:
        b.train_step(x)
    got = b.metrics()
    for k in ("reconstruction_ll", "kl", "matching_ll"):
        assert abs(got[k] - want[k]) <= 2e-5 * max(1.0, abs(want[k])), (k, got[k], want[k])
    # eval_pm_vae_uci.py:76-80: model from model_config.json, parameters from train_state.pkl
    from posterior_matching_b200 import PosteriorMatchingVAE
    m = PosteriorMatchingVAE.from_config(ck.load_model_config(str(tmp_path / "run")), precision="fp32")
    ts = ck.load_params_into(m, path)
    assert int(ts.step) == 5003
    assert torch.equal(m.params["decoder_dist"]["log_scale"].cpu(), torch.as_tensor(ts.params["decoder_dist"]["log_scale"]))


def test_array_dataset_batches():
    from posterior_matching_b200.data import load_datasets
    rng = np.random.default_rng(0)
    arrays = {"train": rng.standard_normal((1000, 8)).astype(np.float32), "val": rng.standard_normal((300, 8)).astype(np.float32)}
    cfg = {"train_split": "train", "validation_split": "val", "train_batch_size": 128, "val_batch_size": 64,
           "training_noise": 0.0, "mask_generator": "BernoulliMaskGenerator", "buffer_size": 100}
    train, val = load_datasets(arrays, cfg, seed=3)
    seen = []
    for batch in train:
        assert set(batch) == {"features", "mask"} and batch["features"].shape == (128, 8) and batch["mask"].shape == (128, 8)
        assert batch["features"].is_cuda and set(batch["mask"].unique().tolist()) <= {0.0, 1.0}
        seen.append(batch["features"].cpu().numpy().copy())
    assert len(seen) == 1000 // 128                       # drop_remainder
    got = np.concatenate(seen)
    # every delivered row is a row of the training array, none twice (shuffle buffer = sampling without replacement)
    keys = {r.tobytes() for r in arrays["train"]}
    assert all(r.tobytes() in keys for r in got) and len({r.tobytes() for r in got}) == got.shape[0]
    assert not np.array_equal(got, arrays["train"][:got.shape[0]])      # shuffled
    vb = [b["features"].cpu().numpy() for b in val]
    assert np.array_equal(np.concatenate(vb), arrays["val"][:256])       # validation order kept, remainder dropped
