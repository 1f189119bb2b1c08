"""The fused optimizer step (train.FusedOptimizer: clip_grad_norm_ + SGD-momentum / Adam over flat buffers,
SURVEY.md section 8f N1) against torch's own optimizers on the reference's training loop
(train.py:359-366, 457-471)."""
import copy

import numpy as np
import pytest
import torch

from gpu_util import assert_close, build_model

pytestmark = pytest.mark.gpu

CFG = dict(grid=10, C=8, L1=64, L2=32, L3=8, NC=10, model_input=32)


def _data(B=96, seed=5):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 3, 32, 32, generator=g).cuda(), torch.randint(0, 10, (B,), generator=g).cuda()


@pytest.mark.parametrize("kind,kw", [
    ("sgd", dict(lr=0.01, momentum=0.9, weight_decay=2e-4, max_grad_norm=1.0)),   # config/train_nnue_default.py
    ("sgd", dict(lr=0.05, momentum=0.0, weight_decay=0.0, max_grad_norm=0.0)),
    ("adam", dict(lr=1e-3, weight_decay=5e-4, max_grad_norm=0.0)),                # config/train_nnue_test.py
    ("adam", dict(lr=1e-3, weight_decay=5e-4, max_grad_norm=0.05)),
], ids=["sgd_default", "sgd_plain", "adam_test", "adam_clipped"])
def test_fused_step_matches_torch_optimizer(kind, kw):
    from nnue_vision_b200 import train
    torch.manual_seed(3)
    model = build_model(CFG)
    ref_model = copy.deepcopy(model)
    images, labels = _data()
    dp = train.DataParallelStep(model)
    opt = train.FusedOptimizer(dp, kind, **kw)
    if kind == "sgd":
        ref_opt = torch.optim.SGD(ref_model.parameters(), lr=kw["lr"], momentum=kw["momentum"], weight_decay=kw["weight_decay"])
    else:
        ref_opt = torch.optim.Adam(ref_model.parameters(), lr=kw["lr"], weight_decay=kw["weight_decay"])
    for it in range(4):
        # ours: flat buffers, fused kernels
        loss = dp.step(images, labels)
        opt.step()
        # reference loop: same gradient kernels, torch clip + torch optimizer
        ref_opt.zero_grad()
        ref_loss = ref_model.loss(images, labels)
        ref_loss.backward()
        if kw["max_grad_norm"] > 0:
            total = torch.nn.utils.clip_grad_norm_(ref_model.parameters(), kw["max_grad_norm"])
            assert_close(opt.grad_norm(), total, f"step {it}: total gradient norm", rtol=1e-5)
        ref_opt.step()
        assert_close(loss, ref_loss, f"step {it}: loss", rtol=2e-5)
        ref_named = dict(ref_model.named_parameters())
        for name, p in model.named_parameters():
            assert_close(p, ref_named[name], f"step {it}: {name}", rtol=2e-5)
    assert float(dict(model.named_parameters())["nnue2score"]) == 600.0


def test_flat_parameters_keep_the_state_dict_contract():
    from nnue_vision_b200 import train
    torch.manual_seed(4)
    model = build_model(CFG)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    dp = train.DataParallelStep(model)
    opt = train.FusedOptimizer(dp, "sgd", lr=0.1)
    after = model.state_dict()
    assert list(after) == list(before)
    for k in before:
        assert torch.equal(after[k], before[k]) and after[k].shape == before[k].shape
    named = dict(model.named_parameters())
    off = 0
    for n in opt.params.names:  # parameters alias the flat buffer, in the reference's registration order
        assert named[n].data_ptr() == opt.params.flat.data_ptr() + 4 * off
        off += named[n].numel()
    assert off == opt.n and dp.buf.numel() == (off + 1 + 3) // 4 * 4  # gradients + loss slot, padded to whole float4s
    # the model still serialises and still loads reference checkpoints
    model.load_state_dict(before)
    images, labels = _data(32)
    dp.step(images, labels)
    opt.step()
    assert not torch.equal(named["input.weight"], before["input.weight"].cuda())
    assert np.isfinite(opt.params.flat.cpu().numpy()).all()
