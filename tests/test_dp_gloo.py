"""Host logic of the batch-sharded data-parallel step (train.DataParallelStep) on CPU: world_size 2,
gloo backend, 127.0.0.1 rendezvous.  The CUDA local step is replaced by the oracle's closed-form step
(the injectable `local_step` exists for exactly this), so what is tested is the part that is ours on
the host: the flat gradient buffer layout, the 1/(global batch) scaling, the single all-reduce and
that `.grad` of every parameter is a view of the reduced buffer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import ROOT  # noqa: F401  (puts the repo root on sys.path)

CFG = dict(grid=8, C=4, L1=64, L2=4, L3=8, NC=10, input=32)
B_GLOBAL = 24


def _model():
    from nnue_vision_b200 import nnue
    torch.manual_seed(42)
    return nnue.NNUE(nnue.GridFeatureSet(CFG["grid"], CFG["C"]), CFG["L1"], CFG["L2"], CFG["L3"],
                     num_classes=CFG["NC"], input_size=CFG["input"])


def _batch():
    g = torch.Generator().manual_seed(7)
    return torch.randn(B_GLOBAL, 3, 32, 32, generator=g), torch.randint(0, CFG["NC"], (B_GLOBAL,), generator=g)


def _oracle_local_step(model):
    """CPU stand-in for the CUDA hot path with the same contract: fill buf.views with the gradient of
    sum_b CE_b * inv_count over THIS shard and buf.loss with that sum."""
    from oracle import float_oracle as fo

    def run(images, labels, inv_count, buf):
        state = {k: v.detach() for k, v in model.state_dict().items()}
        out = fo.step(state, images, labels, model.conv.stride[0], dtype=torch.float64)
        scale = images.shape[0] * inv_count  # the oracle returns the mean over its own batch
        for name, view in zip(buf.names, buf.views):
            view.copy_((out["grads"][name] * scale).to(torch.float32))
        buf.loss.fill_(float(out["loss"]) * scale)

    return run


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shard_sizes, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nnue_vision_b200 import train
        model = _model()
        images, labels = _batch()
        lo = sum(shard_sizes[:rank])
        hi = lo + shard_sizes[rank]
        dp = train.DataParallelStep(model, local_step=_oracle_local_step(model), device="cpu")
        assert dp.world == world
        loss = dp.step(images[lo:hi], labels[lo:hi], global_batch=B_GLOBAL)
        named = dict(model.named_parameters())
        for n, v in zip(dp.buf.names, dp.buf.views):  # .grad aliases the flat buffer
            assert named[n].grad.data_ptr() == v.data_ptr()
        assert named["nnue2score"].grad is None
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), loss=float(loss), flat=dp.buf.flat.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shard_sizes", [(12, 12), (17, 7)], ids=["equal", "ragged"])
def test_two_rank_step_equals_single_process_step(tmp_path, shard_sizes):
    from nnue_vision_b200 import train
    from oracle import float_oracle as fo
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), shard_sizes, str(tmp_path)), nprocs=world, join=True)
    # single-process answer on the concatenated batch
    model = _model()
    images, labels = _batch()
    ref = fo.step({k: v.detach() for k, v in model.state_dict().items()}, images, labels, model.conv.stride[0],
                  dtype=torch.float64)
    buf = train.FlatGradBuffer(dict(model.named_parameters()), "cpu")
    expect = torch.cat([ref["grads"][n].reshape(-1) for n in buf.names] + [ref["loss"].reshape(1)]).numpy()
    got = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert np.array_equal(got[0]["flat"], got[1]["flat"]), "all ranks must hold the same reduced buffer"
    scale = np.abs(expect).max()
    # (the buffer is padded to whole float4s behind the loss slot; the padding stays zero)
    np.testing.assert_allclose(got[0]["flat"][: len(expect)], expect, rtol=1e-5, atol=1e-5 * scale)
    assert len(got[0]["flat"]) % 4 == 0 and not got[0]["flat"][len(expect):].any()
    assert abs(got[0]["loss"] - float(ref["loss"])) < 1e-5 * abs(float(ref["loss"]))


def test_flat_buffer_layout_follows_reference_parameter_order():
    from nnue_vision_b200 import train
    model = _model()
    buf = train.FlatGradBuffer(dict(model.named_parameters()), "cpu")
    assert buf.names == list(train.GRAD_PARAM_NAMES)
    n_real = sum(p.numel() for n, p in model.named_parameters() if n != "nnue2score") + 1  # gradients + the loss slot
    assert buf.numel() == (n_real + 3) // 4 * 4  # padded to whole float4s
    off = 0
    for n, v in zip(buf.names, buf.views):
        assert v.data_ptr() == buf.flat.data_ptr() + 4 * off and v.shape == dict(model.named_parameters())[n].shape
        off += v.numel()
    assert buf.loss.data_ptr() == buf.flat.data_ptr() + 4 * off


def test_single_process_step_issues_no_collective():
    from nnue_vision_b200 import train
    model = _model()
    images, labels = _batch()
    dp = train.DataParallelStep(model, local_step=_oracle_local_step(model), device="cpu")
    assert dp.world == 1
    loss = dp.step(images, labels)
    from oracle import float_oracle as fo
    ref = fo.step({k: v.detach() for k, v in model.state_dict().items()}, images, labels, model.conv.stride[0],
                  dtype=torch.float64)
    assert abs(float(loss) - float(ref["loss"])) < 1e-6


def test_step_reattaches_gradients_after_zero_grad():
    """torch's optimizer.zero_grad() drops .grad (set_to_none=True is the default): the next step() must make the
    flat-buffer views the gradients again, or a torch optimizer would silently see nothing."""
    from nnue_vision_b200 import train
    model = _model()
    images, labels = _batch()
    dp = train.DataParallelStep(model, local_step=_oracle_local_step(model), device="cpu")
    dp.step(images, labels)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    opt.zero_grad()
    named = dict(model.named_parameters())
    assert all(named[n].grad is None for n in dp.buf.names)
    before = named["input.weight"].detach().clone()
    dp.step(images, labels)
    for n, v in zip(dp.buf.names, dp.buf.views):
        assert named[n].grad is not None and named[n].grad.data_ptr() == v.data_ptr()
    opt.step()
    assert not torch.equal(named["input.weight"], before)


def _slice_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nnue_vision_b200 import train
        model = _model()
        buf = train.FlatGradBuffer(dict(model.named_parameters()), "cpu")
        g = torch.Generator().manual_seed(100 + rank)
        vals = torch.randn(buf.numel(), generator=g)
        whole = vals.clone()
        dist.all_reduce(whole)
        buf.flat.copy_(vals)
        x = train.CollectiveExchange(buf, None)
        x.early()   # [split, end): feature transformer, head, loss -- final before the conv gradient
        x.late()    # [0, split): thresholds and conv weights
        np.savez(os.path.join(out_dir, f"slices{rank}.npz"), got=buf.flat.numpy(), whole=whole.numpy(), split=buf.split)
    finally:
        dist.destroy_process_group()


def test_slice_exchange_equals_whole_buffer_exchange(tmp_path):
    """The step exchanges the flat buffer as two slices (early / late, train.OneShotExchange / CollectiveExchange); on
    two gloo ranks the slices together give exactly the whole-buffer all-reduce, and the split falls on a float4
    boundary behind the conv / threshold gradients."""
    world = 2
    mp.spawn(_slice_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"slices{k}.npz") for k in range(world)]
    assert np.array_equal(r[0]["got"], r[0]["whole"]) and np.array_equal(r[1]["got"], r[0]["got"])
    split = int(r[0]["split"])
    assert split % 4 == 0 and CFG["C"] * 28 <= split < CFG["C"] * 28 + 4
