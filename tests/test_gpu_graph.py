"""-m gpu: the CUDA-graph replay of the training step gives bit-identical results to the eager launch
sequence, follows in-place parameter updates, and re-captures when the inputs move."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_util import build_model


CFG = dict(grid=10, C=8, L1=64, L2=32, L3=8, NC=10, model_input=32)


def test_graph_replay_matches_eager_and_tracks_parameter_updates():
    from nnue_vision_b200 import train
    torch.manual_seed(3)
    model_g = build_model(CFG)
    model_e = copy.deepcopy(model_g)
    dp_g = train.DataParallelStep(model_g, cuda_graphs=True)
    dp_e = train.DataParallelStep(model_e, cuda_graphs=False)
    g = torch.Generator().manual_seed(5)
    sets = [(torch.randn(300, 3, 32, 32, generator=g).cuda(), torch.randint(0, 10, (300,), generator=g).cuda())
            for _ in range(2)]
    for it in range(8):
        images, labels = sets[it % 2]
        lg = dp_g.step(images, labels).clone()
        le = dp_e.step(images, labels).clone()
        assert torch.equal(lg, le), f"iteration {it}: loss differs"
        assert torch.equal(dp_g.buf.flat, dp_e.buf.flat), f"iteration {it}: gradients differ"
        with torch.no_grad():  # an optimizer-like in-place update: the captured graph must see the new values
            for pg, pe in zip(model_g.parameters(), model_e.parameters()):
                if pg.grad is not None:
                    pg.add_(pg.grad, alpha=-0.05)
                    pe.add_(pe.grad, alpha=-0.05)
    captured = [e for e in dp_g._graphs.values() if e[1] is not None]
    assert len(captured) == 2 and all(e[0] >= 2 for e in captured)
    # a new input buffer is a cache miss: first eager, then its own graph
    images2, labels2 = sets[0][0].clone(), sets[0][1].clone()
    a = dp_g.step(images2, labels2).clone()
    b = dp_g.step(images2, labels2).clone()
    c = dp_e.step(images2, labels2).clone()
    assert torch.equal(a, b) and torch.equal(a, c)
    assert len(dp_g._graphs) == 3


def test_module_level_loss_replays_a_graph_and_owns_its_gradients():
    """model.loss (train.compute_loss): the second sighting of the same buffers is captured, later ones replayed; the
    result equals the eager launch sequence bit for bit, the upstream factor is applied, and a node's gradients
    survive another step of the same model before its backward() runs."""
    from nnue_vision_b200 import nnue as nn_mod
    torch.manual_seed(3)
    model_g = build_model(CFG)
    model_e = copy.deepcopy(model_g)
    g = torch.Generator().manual_seed(5)
    xa, ya = torch.randn(300, 3, 32, 32, generator=g).cuda(), torch.randint(0, 10, (300,), generator=g).cuda()
    xb, yb = torch.randn(300, 3, 32, 32, generator=g).cuda(), torch.randint(0, 10, (300,), generator=g).cuda()

    def grads(model, x, y, scale, graphs):
        old = nn_mod.CUDA_GRAPHS
        nn_mod.CUDA_GRAPHS = graphs
        try:
            model.zero_grad(set_to_none=True)
            loss = model.loss(x, y)
            (loss * scale).backward()
        finally:
            nn_mod.CUDA_GRAPHS = old
        return loss.detach().clone(), {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}

    for it in range(4):
        lg, gg = grads(model_g, xa, ya, 1.0 + it, True)
        le, ge = grads(model_e, xa, ya, 1.0 + it, False)
        assert torch.equal(lg, le)
        for k in ge:
            assert torch.equal(gg[k], ge[k]), (it, k)
    runner = nn_mod._GRAPH_RUNNERS[model_g]
    assert any(e[1] is not None for e in runner._graphs.values()), "no graph was captured"
    assert all(p.grad is None or p.grad.data_ptr() != runner.buf.flat.data_ptr() for p in model_g.parameters())
    # two forwards before any backward: each node keeps its own gradients
    model_g.zero_grad(set_to_none=True)
    la = model_g.loss(xa, ya)
    lb = model_g.loss(xb, yb)
    la.backward()
    ga = {k: p.grad.clone() for k, p in model_g.named_parameters() if p.grad is not None}
    _, ea = grads(model_e, xa, ya, 1.0, False)
    for k in ea:
        assert torch.equal(ga[k], ea[k]), k
    assert copy.deepcopy(model_g) is not None  # the runner lives outside the module


def test_graph_cache_is_bounded():
    from nnue_vision_b200 import train
    torch.manual_seed(4)
    model = build_model(CFG)
    dp = train.DataParallelStep(model, cuda_graphs=True, max_graphs=2)
    labels = torch.randint(0, 10, (64,)).cuda()
    keep = []
    for _ in range(5):
        images = torch.randn(64, 3, 32, 32).cuda()
        keep.append(images)
        dp.step(images, labels)
        dp.step(images, labels)
    assert len(dp._graphs) == 2
