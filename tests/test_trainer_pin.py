"""Pins the oracle's restatement of the trainer glue (SURVEY 8 a10 loss, a12 token post-processing, macro-F1) against the
UNMODIFIED reference code: `ASR.compute_objectives`, `unpad`, `idx2string`, `caculate_f1` of pretrain.py / finetune.py are imported
live in the build container (speechbrain & co. stubbed: the functions under test only use torch / numpy / scikit-learn) and fed
the same tensors as the oracle.  /root/reference does not travel, so these tests skip elsewhere; the values they pin are the ones
every GPU parity test compares the CUDA path against."""
import types

import numpy as np
import pytest
import torch

from helpers import make_ground_truth
from oracle import a2s_oracle as O
from oracle import metrics_oracle as MO
from refimport import have_reference_tree as have_reference

pytestmark = pytest.mark.skipif(not have_reference(), reason="/root/reference is only mounted in the build container")


def _brain(mod, stage_train=True):
    """An `ASR` instance without speechbrain's __init__: the attributes compute_objectives touches, losses as in pretrain.yaml:49-54."""
    asr = object.__new__(mod.ASR)
    asr.hparams = types.SimpleNamespace(loss_time_sig=torch.nn.NLLLoss(), loss_key=torch.nn.NLLLoss(),
                                        loss_score=torch.nn.NLLLoss(ignore_index=147))
    for n in ("time_losses", "key_losses", "upper_losses", "lower_losses"):
        setattr(asr, n, [])
    for n in ("upper_pred", "upper_target", "lower_pred", "lower_target", "key_pred", "key_target", "time_sig_pred", "time_sig_target"):
        setattr(asr, n, {})
    return asr


def _batch(B=3, bars=2, Lu=14, Ll=9, seed=4):
    gt = make_ground_truth(B, bars, Lu, Ll, seed=seed, lo_up=(3, 13), lo_lo=(2, 9))
    g = torch.Generator().manual_seed(seed)
    preds = [torch.log_softmax(torch.randn(B, bars, n, generator=g), -1) for n in (7, 14)]
    preds += [torch.log_softmax(torch.randn(B, bars, L, 173, generator=g), -1) for L in (Lu, Ll)]
    batch = (torch.zeros(B, 1, 4, 4), gt[0], gt[1], gt[2], gt[3], gt[4], gt[5], [f"song{b}" for b in range(B)], torch.arange(B))
    return preds, gt, batch


@pytest.mark.parametrize("script", ["pretrain", "finetune"])
def test_loss_matches_reference_compute_objectives(script):
    from refimport import import_reference_trainer
    mod = import_reference_trainer(script)
    sb = __import__("speechbrain")
    preds, gt, batch = _batch()
    asr = _brain(mod)
    ref = mod.ASR.compute_objectives(asr, tuple(p.clone().requires_grad_(True) for p in preds), batch, sb.Stage.TRAIN)
    mine = O.training_loss(preds, gt)
    assert torch.equal(ref.detach(), mine)
    # the four parts the reference logs are the four NLL terms
    parts = [float(asr.time_losses[0]), float(asr.key_losses[0]), float(asr.upper_losses[0]), float(asr.lower_losses[0])]
    assert abs(sum(parts) - float(mine)) < 1e-5


@pytest.mark.parametrize("script", ["pretrain", "finetune"])
def test_validation_records_match_oracle_tokens(script):
    from refimport import import_reference_trainer
    mod = import_reference_trainer(script)
    sb = __import__("speechbrain")
    preds, gt, batch = _batch(seed=6)
    # make some predicted rows end early: put <eos> as the argmax at a few positions
    preds[2][0, 0, 3, 146] = 5.0
    preds[3][1, 1, 0, 146] = 5.0
    asr = _brain(mod)
    mod.ASR.compute_objectives(asr, tuple(preds), batch, sb.Stage.VALID)
    want = O.greedy_tokens(preds)
    for b in range(3):
        rid = "~".join([str(b), f"song{b}"]) if script == "pretrain" else f"song{b}"       # pretrain.py:99 / finetune.py:89
        assert asr.upper_pred[rid] == want["upper"][b] and asr.lower_pred[rid] == want["lower"][b]
        assert asr.key_pred[rid] == want["key"][b] and asr.time_sig_pred[rid] == want["time_sig"][b]
        assert asr.upper_target[rid] == [O.unpad(r) for r in gt[2][b].tolist()]
    # idx2string and unpad themselves
    assert mod.idx2string([0, 5, 144]) == MO.idx2string([0, 5, 144], __import__("models").labels.labels_map_inv)
    row = torch.tensor([4, 9, 146, 3, 146])
    assert mod.unpad(row).tolist() == O.unpad(row.tolist()) == [4, 9]
    assert mod.unpad(torch.tensor([1, 2, 3])).tolist() == [1, 2, 3]


def test_macro_f1_matches_reference_caculate_f1():
    from refimport import import_reference_trainer
    from piano_a2s_b200.metrics import macro_f1
    import warnings
    mod = import_reference_trainer("pretrain")
    rng = np.random.RandomState(2)
    pred = {f"c{i}": rng.randint(0, 14, 5).tolist() for i in range(12)}
    target = {f"c{i}": rng.randint(0, 14, 5).tolist() for i in range(12)}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mean, per = mod.caculate_f1(pred, target)
    mine = {k: macro_f1(target[k], pred[k]) for k in pred}
    assert all(abs(mine[k] - per[k]) < 1e-12 for k in pred) and abs(np.mean(list(mine.values())) - mean) < 1e-12


def test_optimizer_step_matches_torch_clip_and_adadelta():
    """a11: the reference steps `torch.optim.Adadelta(lr=1.0, rho=0.95, eps=1e-8)` (pretrain.yaml:44-47) after speechbrain's
    `check_gradients` has clipped the global gradient norm [max_grad_norm = 5.0, speechbrain's default, recalled].  The oracle's
    `adadelta_step` against exactly those torch calls, over several steps and with gradients both above and below the clip."""
    g = torch.Generator().manual_seed(3)
    shapes = [(7, 5), (11,), (3, 4, 2)]
    ref_params = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in shapes]
    opt = torch.optim.Adadelta(ref_params, lr=1.0, rho=0.95, eps=1e-8)
    mine = {str(i): p.detach().clone() for i, p in enumerate(ref_params)}
    state = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in mine.items()}
    for step, scale in enumerate((10.0, 0.01, 3.0, 1.0)):                  # norms far above, far below and around 5.0
        grads = [scale * torch.randn(s, generator=g) for s in shapes]
        for p, gr in zip(ref_params, grads):
            p.grad = gr.clone()
        total_ref = torch.nn.utils.clip_grad_norm_(ref_params, 5.0)
        opt.step()
        total = O.adadelta_step(mine, {str(i): gr.clone() for i, gr in enumerate(grads)}, state)
        assert abs(float(total) - float(total_ref)) <= 1e-6 * float(total_ref)
        for i, p in enumerate(ref_params):
            assert torch.allclose(mine[str(i)], p.detach(), rtol=1e-6, atol=1e-7), (step, i)


@pytest.mark.parametrize("script", ["pretrain", "finetune"])
def test_reference_call_site_binds_to_the_drop_in_forward(script):
    """The keyword arguments the reference's `ASR.compute_forward` passes (pretrain.py:41-53, finetune.py:40-52) bind to the
    drop-in `models.ScoreTranscription.forward` in both stages, with the values the reference uses."""
    import inspect
    import models
    from refimport import import_reference_trainer
    mod = import_reference_trainer(script)
    sb = __import__("speechbrain")
    calls = []

    def transcription(**kw):
        calls.append(kw)
        return tuple(torch.zeros(1) for _ in range(4))
    asr = object.__new__(mod.ASR)
    asr.modules = types.SimpleNamespace(transcription=transcription)
    asr.teacher_forcing_ratio, asr.device = 0.7, "cuda:0"                     # pretrain.py keeps the (decaying) ratio on the Brain,
    asr.hparams = types.SimpleNamespace(teacher_forcing_ratio=0.7)            # finetune.py reads it from the hparams
    _, gt, batch = _batch()
    mod.ASR.compute_forward(asr, batch, sb.Stage.TRAIN)
    mod.ASR.compute_forward(asr, batch, sb.Stage.VALID)
    sig = inspect.signature(models.ScoreTranscription.forward)
    for kw in calls:
        bound = sig.bind(None, **kw)                                          # raises TypeError on an unknown / missing argument
        assert set(kw) == {"spectrogram", "inference", "ground_truth", "teacher_forcing_ratio", "device"}
        assert bound.arguments["device"] == "cuda:0"
    train, valid = calls
    assert train["inference"] is False and train["teacher_forcing_ratio"] == 0.7 and len(train["ground_truth"]) == 6
    assert all(a is b for a, b in zip(train["ground_truth"], gt)) or all(torch.equal(a, b) for a, b in zip(train["ground_truth"], gt))
    assert valid["inference"] is True and valid["ground_truth"] is None and valid["teacher_forcing_ratio"] == 0.


def test_teacher_forcing_schedule_matches_reference_on_stage_start():
    """pretrain.py:149-153, executed live on a stub Brain for a few epochs and both stages."""
    from refimport import import_reference_trainer
    from piano_a2s_b200.train import teacher_forcing_schedule
    mod = import_reference_trainer("pretrain")
    sb = __import__("speechbrain")
    asr = _brain(mod)
    asr.hparams.teacher_forcing_ratio, asr.hparams.teacher_forcing_decay = 0.7, 0.99
    mod.load = lambda path: []                                                # on_stage_start also reads a metadata json (pretrain.py:148)
    for epoch in (0, 1, 7, 29):
        mod.ASR.on_stage_start(asr, sb.Stage.TRAIN, epoch)
        assert asr.teacher_forcing_ratio == teacher_forcing_schedule(0.7, 0.99, epoch)
    asr.hparams.output_folder = "/tmp/pa2s_pin_out"
    mod.mkdirs = lambda *a, **k: None
    mod.ASR.on_stage_start(asr, sb.Stage.VALID, 3)
    assert asr.teacher_forcing_ratio == teacher_forcing_schedule(0.7, 0.99, 3, training=False) == 0.
