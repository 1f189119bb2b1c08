"""Where the gradient exchange's cost inside the config-D step comes from: the same N-rank step timed with no exchange,
with only the early slice (feature transformer + head, on the side stream behind the table gradient), with only the late
slice (conv weights + thresholds, at the end of the step), with both as whole exchanges, and with the early slice pushed
early and collected at the end (the default).  DIAGNOSTIC: only the `both_*` lines are training steps.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/exchange_parts.py
"""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import bench
from nnue_vision_b200 import train


class Parts:
    def __init__(self, x, early, late):
        self.x, self.e, self.l = x, early, late

    def early(self, stream=None):
        if self.e:
            self.x.early(stream)

    def late(self, stream=None):
        if self.l:
            self.x.late(stream)

    def full(self, stream=None):
        self.early(stream)
        self.late(stream)


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    torch.distributed.init_process_group("nccl", device_id=device)
    w = dict(bench.WORKLOADS["default_cifar_b16384"])
    B = w["batch"]
    model = bench.build_model(w, device)
    for p in model.parameters():
        torch.distributed.broadcast(p.data, src=0)
    dp = train.DataParallelStep(model)
    real = dp._xchg
    sets = [bench.synthetic_batch(w, B, seed=1000 * rank + i, device=device) for i in range(3)]
    res = {}
    # `split`: the early slice is pushed behind the table gradient and collected at the end of the step with the late
    # slice's exchange (the default); the others run each slice as a whole exchange where it is issued
    for name, (e, l, split) in {"both_split": (1, 1, True), "none": (0, 0, False), "both_whole": (1, 1, False),
                                "early_only_whole": (1, 0, False), "late_only_whole": (0, 1, False),
                                "both_split_again": (1, 1, True)}.items():
        real.split_phases = split
        dp._xchg = Parts(real, e, l)
        dp._graphs.clear()
        for _ in range(2):
            for imgs, labs in sets:
                dp.step(imgs, labs, global_batch=B * world)
        for i in range(6):
            dp.step(*sets[i % 3], global_batch=B * world)
        wins = []
        for _ in range(5):
            bench.barrier(world)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(30):
                dp.step(*sets[i % 3], global_batch=B * world)
            e1.record()
            bench.barrier(world)
            wins.append(bench.max_over_ranks(e0.elapsed_time(e1) / 30, world, device))
        res[name] = sorted(wins)[2]
        if rank == 0:
            print(f"{name:12s} {1e3 * res[name]:8.2f} us per step", flush=True)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "exchange": dp.allreduce, "us_per_step": {k: 1e3 * v for k, v in res.items()}}))
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
