"""-m gpu, needs >= 2 GPUs (skipped otherwise): the data-parallel step over NCCL and over the library's one-shot
all-reduce on NVLink peer memory give the gradients of the full batch, identically on every rank."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(grid=10, C=8, L1=64, L2=32, L3=8, NC=10, model_input=32)


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import sys
        from pathlib import Path
        sys.path.insert(0, str(Path(__file__).resolve().parent))
        from gpu_util import build_model
        from nnue_vision_b200 import train
        torch.manual_seed(11)
        model = build_model(CFG, device=f"cuda:{rank}")
        g = torch.Generator().manual_seed(21)
        per = 320
        images = torch.randn(world * per, 3, 32, 32, generator=g)
        labels = torch.randint(0, 10, (world * per,), generator=g)
        mine = slice(rank * per, (rank + 1) * per)
        x, y = images[mine].cuda(), labels[mine].cuda()
        res = {}
        for kind in ("oneshot", "nccl"):
            dp = train.DataParallelStep(model, allreduce=kind)
            losses = []
            for _ in range(4):  # eager, capture, replays: the exchange follows each of them
                losses.append(float(dp.step(x, y)))
            torch.cuda.synchronize()
            res[kind] = (dp.allreduce, losses, dp.buf.flat.clone())
        if rank == 0:  # single-GPU reference on the whole batch
            ref = train.DataParallelStep(model, cuda_graphs=False, single=True)  # one rank, no exchange
            ref_loss = float(ref.step(images.cuda(), labels.cuda()))
            torch.cuda.synchronize()
            res["ref"] = ("none", [ref_loss], ref.buf.flat.clone())
        flat = res["oneshot"][2]
        others = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(others, flat)
        same_everywhere = all(torch.equal(o, flat) for o in others)
        if rank == 0:
            torch.save({k: (v[0], v[1], v[2].cpu()) for k, v in res.items()} | {"same": same_everywhere}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_one_shot_allreduce_matches_nccl_and_full_batch(tmp_path, world):
    """Both exchanges (issued inside the step, slice by slice, the one-shot form captured in the step's CUDA graph) at
    every world size the scaling run uses; bench.py repeats the one-shot-vs-NCCL check wherever it runs with N > 1
    (`exchange_check`), because the driver's GPU test tier has a single GPU."""
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(world, 29641 + world, out), nprocs=world, join=True)
    res = torch.load(out)
    kind, losses, flat = res["oneshot"]
    assert res["same"], "ranks disagree"
    assert all(l == losses[0] for l in losses), "the step is deterministic: every repeat gives the same loss"
    nk, nlosses, nflat = res["nccl"]
    assert nk == "collective"
    if kind == "oneshot_p2p" and world == 2:  # a + b in rank order: identical to the two-rank NCCL sum
        assert torch.equal(flat, nflat)
    ntol = 1e-6 * nflat.abs() + 1e-6 * nflat.abs().max()
    assert ((flat - nflat).abs() <= ntol).all(), "one-shot and NCCL disagree beyond summation order"
    _, (ref_loss,), rflat = res["ref"]
    tol = 1e-5 * rflat.abs() + 1e-5 * rflat.abs().max()
    assert ((flat - rflat).abs() <= tol).all()
    assert abs(losses[0] - ref_loss) <= 1e-5 * abs(ref_loss)
    print("exchange used:", kind)
