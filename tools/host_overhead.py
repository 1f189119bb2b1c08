"""Host enqueue time per training step vs device time (is the step launch-bound on the host?)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from nnue_vision_b200 import train

w = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "default_cifar_b16384"])
dev = torch.device("cuda", 0)
model = bench.build_model(w, dev)
dp = train.DataParallelStep(model)
images, labels = bench.synthetic_batch(w, w["batch"], seed=1, device=dev)
for _ in range(5):
    dp.step(images, labels)
torch.cuda.synchronize()
n = 200
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(n):
    dp.step(images, labels)
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e6 * (t1 - t0) / n:.1f} us/step, device {1e3 * e0.elapsed_time(e1) / n:.1f} us/step, wall {1e6 * (t2 - t0) / n:.1f} us/step")
