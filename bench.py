#!/usr/bin/env python3
"""bench.py -- NNUE train samples/s (fwd+bwd) on B200, with roofline, CPU baseline and e2e legs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME]   # our arm (N=1 default)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W               # N > 1: one rank per GPU, NCCL
    python bench.py --impl reference [...]                                   # the reference's CPU path

A "step" is one pass of the hot path (grid-feature extraction -> feature transformer ->
pairwise + head -> mean CE -> full backward -> gradient all-reduce when N > 1) over one batch of
synthetic input.  The default workload is config `train_nnue_default.py` at batch 16384 per GPU
(BASELINE configs[1], weak scaling); `--workload` selects the reference's other configurations
(T, D1k, I-s, I).  Prints ONE JSON line on rank 0.

Timing: W >= 3 warm-up steps, then `--windows` (default 5) timed windows of EXACTLY K steps each, every
window bracketed by barrier + synchronize and timed with CUDA events on the launching stream (max over
ranks); `value` / `ms_per_step` are the MEDIAN window, all windows are listed in `windows_ms`.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
REF_DIR = ROOT / "baseline" / "_ref"   # the unmodified reference's nnue.py (git-ignored; __graft_entry__.build copies it)

# reference configurations; `ref_batch` = samples per step of the CPU reference arm (its cost is
# linear in the batch: per-sample Python loops, and a dense [F, L1] gradient allocated per sample in the backward)
WORKLOADS = {
    # config/train_nnue_default.py:16-35 (batch overridden to the BASELINE value)
    "default_cifar_b16384": dict(grid=10, C=8, L1=64, L2=32, L3=8, NC=10, input=32, image=32, batch=16384, ref_batch=256),
    # config/train_nnue_test.py:9-23
    "test_cifar_b16": dict(grid=8, C=4, L1=64, L2=4, L3=8, NC=10, input=32, image=32, batch=16, ref_batch=16),
    # config/train_nnue.py:16-36 (the reference's "real" config)
    "real_cifar_l1_1024_b16384": dict(grid=10, C=8, L1=1024, L2=128, L3=32, NC=10, input=32, image=32, batch=16384, ref_batch=64),
    # engine test shape (engine/tests/test_nnue_engine.cpp:12-16) on ImageNet-shaped input, 1000 classes
    "imagenet_small_b2048": dict(grid=16, C=32, L1=256, L2=16, L3=32, NC=1000, input=224, image=224, batch=2048, ref_batch=16),
    "imagenet_small_b16384": dict(grid=16, C=32, L1=256, L2=16, L3=32, NC=1000, input=224, image=224, batch=16384, ref_batch=16),
    # largest grid the integer engine supports (64 channels, nnue_engine.h:243; serialize.py:745): F = 65536, table 268 MB
    "imagenet_large_b4096": dict(grid=32, C=64, L1=1024, L2=128, L3=32, NC=1000, input=224, image=224, batch=4096, ref_batch=2),
}
METRIC = "nnue_train_samples_per_sec_fwd_bwd"
UNIT = "samples/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--windows", type=int, default=5, help="timed windows of --steps steps each; the median is reported")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="default_cifar_b16384", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: the workload's batch is the GLOBAL batch, split evenly over the GPUs (default: weak)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "oneshot", "nccl", "none"],
                    help="gradient exchange at N > 1: the library's one-shot all-reduce over NVLink peer memory (auto) or NCCL; "
                         "`none` is a DIAGNOSTIC (every rank steps on its shard without exchanging gradients: not a training "
                         "step, it separates the exchange's cost from what running N GPUs at once costs each of them)")
    ap.add_argument("--no-int", action="store_true", help="skip the integer-inference side measurement")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (extra scaling lines only)")
    ap.add_argument("--no-module-api", action="store_true", help="skip the compute_loss(model, batch); loss.backward() leg")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE",
                    help="library tuning knob (nnue_set_option), e.g. --opt input_bwd_variant=1")
    return ap.parse_args()


def _peaks_file():
    try:
        return json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        return {}


def tensor_peak():
    d = _peaks_file()
    if "bf16_tflops_sustained" in d:
        return float(d["bf16_tflops_sustained"]), "measured sustained bf16 (MEASURED_PEAKS.json)"
    return 1400.0, "fallback (B200_PROFILING.md: sustained bf16)"


def peaks():
    d = _peaks_file()
    if "hbm_gbs" in d:
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 50 ms.  It is started early (nvidia-smi needs a
    moment to come up) and `window()` marks the loaded region -- warm-up, timed steps, end-to-end steps --
    whose samples `summary()` reports."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t0, self.t1 = index, [], None, None, None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def begin(self):
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0 = self.t0 if self.t0 is not None else 0.0
        t1 = (self.t1 if self.t1 is not None else float("inf")) + 0.06  # a sample reports the preceding interval
        for ts, r in self.rows:
            if not (t0 <= ts <= t1):
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "window": "warm-up + timed steps + end-to-end steps"}


def build_model(w, device, seed=42):
    from nnue_vision_b200 import nnue
    torch.manual_seed(seed)  # the reference's own initialisers, same RNG stream (tests/test_serialize.py)
    model = nnue.NNUE(nnue.GridFeatureSet(w["grid"], w["C"]), w["L1"], w["L2"], w["L3"], num_classes=w["NC"],
                      input_size=w["input"])
    return model.to(device)


def synthetic_batch(w, B, seed, device=None, pin=False):
    """N(0,1) images [B,3,H,W] and uniform labels.  CPU tensors come from a CPU generator; for a CUDA
    device (or a pinned host copy of an ImageNet-sized batch) the values are drawn on the device."""
    shape = (B, 3, w["image"], w["image"])
    on_gpu = torch.cuda.is_available() and (pin or (device is not None and torch.device(device).type == "cuda"))
    if on_gpu:
        dev = torch.device(device) if device is not None and torch.device(device).type == "cuda" else torch.device("cuda", torch.cuda.current_device())
        g = torch.Generator(device=dev).manual_seed(seed)
        images = torch.randn(shape, generator=g, device=dev)
        labels = torch.randint(0, w["NC"], (B,), generator=g, device=dev)
        if not pin:
            return images, labels
        himg = torch.empty(shape, dtype=torch.float32, pin_memory=True)
        hlab = torch.empty((B,), dtype=torch.long, pin_memory=True)
        himg.copy_(images); hlab.copy_(labels)
        torch.cuda.synchronize()
        return himg, hlab
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(shape, generator=g)
    labels = torch.randint(0, w["NC"], (B,), generator=g)
    return images, labels


def max_over_ranks(ms, world, device):
    if world == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


RANK_MS = None


def per_rank(v, world, device):
    if world == 1:
        return [v]
    t = torch.zeros(world, dtype=torch.float64, device=device)
    t[torch.distributed.get_rank()] = v
    torch.distributed.all_reduce(t)
    return [float(x) for x in t.tolist()]


def barrier(world):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


# ---- the reference's CPU implementation of the path ------------------------------------------------------------------
def _reference_module():
    """The UNMODIFIED reference nnue.py from baseline/_ref (copied there, git-ignored, by __graft_entry__.build() in the
    build container; it travels to the GPU box with the snapshot).  None when absent."""
    if not (REF_DIR / "nnue.py").exists():
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("reference_nnue", REF_DIR / "nnue.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class ReferenceStepper:
    """One training step of the reference on the host cores: `F.cross_entropy(model(images), targets.long())` +
    `loss.backward()` (train.py:250-254, 352-361) on the stock `nnue.NNUE` (kind "reference"); when baseline/_ref is
    absent, the oracle's restatement of the same per-sample loops (kind "port")."""

    def __init__(self, w, threads):
        import torch.nn.functional as F
        torch.set_num_threads(threads)
        self.B = min(w["ref_batch"], w["batch"])
        self.images, self.labels = synthetic_batch(w, self.B, seed=7, device="cpu")
        ref = _reference_module()
        if ref is not None:
            self.kind = "reference"
            torch.manual_seed(42)
            self.model = ref.NNUE(ref.GridFeatureSet(w["grid"], w["C"]), w["L1"], w["L2"], w["L3"], num_classes=w["NC"],
                                  input_size=w["input"])
            self.model.train()
            self.what = "stock nnue.NNUE forward + F.cross_entropy + backward (baseline/_ref/nnue.py, unmodified)"

            def step():
                self.model.zero_grad(set_to_none=True)
                loss = F.cross_entropy(self.model(self.images), self.labels.long())
                loss.backward()
                return float(loss.detach())
        else:
            from oracle import float_oracle as fo
            self.kind = "port"
            state = cpu_reference_state(w)
            stride = fo.python_stride(w["input"], w["grid"])
            self.what = "oracle.float_oracle.reference_style_step (baseline/_ref absent)"

            def step():
                return fo.reference_style_step(state, self.images, self.labels, stride)
        self.step = step


def cpu_baseline_training(w, seconds, threads=None):
    from oracle.int_oracle import host_threads
    threads = threads or host_threads()
    rs = ReferenceStepper(w, threads)
    rs.step()  # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        rs.step()
        n += 1
        dt = time.perf_counter() - t0
        if dt > seconds or n >= 200:
            break
    return {"value": n * rs.B / dt, "unit": UNIT, "cores": threads, "kind": rs.kind,
            "sample": f"{n} fwd+bwd passes of batch {rs.B} of the workload: {rs.what}, {dt:.1f} s"}


def cpu_reference_state(w):
    """Reference-initialised parameters as plain CPU tensors (no CUDA needed)."""
    from nnue_vision_b200 import nnue
    torch.manual_seed(42)
    m = nnue.NNUE(nnue.GridFeatureSet(w["grid"], w["C"]), w["L1"], w["L2"], w["L3"], num_classes=w["NC"],
                  input_size=w["input"])
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the training path on the host cores, every host
    thread, each step a bounded sample (`ref_batch` samples) of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = dict(WORKLOADS[args.workload])
    if args.batch:
        w["batch"] = args.batch
    from oracle.int_oracle import host_threads
    threads = host_threads()
    rs = ReferenceStepper(w, threads)
    for _ in range(max(1, min(args.warmup, 3))):
        rs.step()
    steps = max(1, min(args.steps, 40))
    t0 = time.perf_counter()
    for _ in range(steps):
        rs.step()
    dt = time.perf_counter() - t0
    value = steps * rs.B / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "batch_per_step": rs.B,
                   "note": "CPU only; each step is a bounded sample of the workload; the reference's cost is linear in the batch"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": rs.kind,
                         "sample": f"{steps} fwd+bwd passes of batch {rs.B}: {rs.what}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def stage_breakdown(dp, images, labels, reps=5):
    """Average device time of each C-ABI stage, from CUDA events recorded on the launching stream."""
    totals = {}
    for _ in range(3):  # the eager, marked path has its own allocations to warm (the timed loop replays a graph)
        dp.step(images, labels, marks=[])
    torch.cuda.synchronize()
    for _ in range(reps):
        marks = []
        dp.step(images, labels, marks=marks)
        torch.cuda.synchronize()
        for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
            totals[n1] = totals.get(n1, 0.0) + e0.elapsed_time(e1)
    return {k: v / reps for k, v in totals.items()}


def int_inference_leg(w, device, seconds):
    """Side measurement: batched bit-exact integer inference (nnue_q_infer) vs the reference's C++ engine
    (oracle/_ref, kind "reference") or the C oracle (kind "port") on the host cores."""
    import tempfile
    import numpy as np
    from nnue_vision_b200 import engine, serialize
    from oracle import int_oracle
    model = build_model(w, "cpu")
    B = w["batch"]
    out = {}
    with tempfile.TemporaryDirectory() as td:
        path = Path(td) / "m.nnue"
        serialize.serialize_model(model, path)
        ev = engine.NNUEEvaluator(path)
        imgs = torch.randn(B, w["image"], w["image"], 3, generator=torch.Generator().manual_seed(3))
        d_imgs = imgs.to(device)
        # the GPU has been idle while the CPU legs ran: spin it back up to its clocks before timing
        t_warm = time.perf_counter()
        while time.perf_counter() - t_warm < 0.5:
            ev.evaluate_logits(d_imgs)
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 200
        e0.record()
        for _ in range(reps):
            logits, dens = ev.evaluate_logits(d_imgs)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out["device"] = {"value": B / (ms * 1e-3), "unit": "samples/s", "batch": B, "ms": ms}
        # BASELINE config "batch sweep 1..65536": device-resident throughput per batch size (images [Bs,32,32,3] views of
        # one buffer; <= 592 samples: a CTA per sample, < 2048: a warp per sample, >= 2048: conv + bitmask -> tcgen05
        # accumulate (+ the layer stack in its epilogue from 8192 up)).  Two numbers per size: the Python call loop
        # (evaluate_logits with reused result tensors; small batches are bound by ~19 us of host time per call) and the
        # same calls replayed as one CUDA graph = device time per call
        sweep, sweep_dev, call_us, dev_us = {}, {}, {}, {}
        big = torch.randn(65536, w["image"], w["image"], 3, generator=torch.Generator().manual_seed(4)).to(device) \
            if w["image"] <= 64 else None
        if big is not None:
            for Bs in (1, 16, 256, 4096, 16384, 65536):
                x = big[:Bs]
                res = (torch.empty(Bs, ev.num_classes, device=device), torch.empty(Bs, device=device))
                for _ in range(3):
                    ev.evaluate_logits(x, out=res)
                n_rep = 50
                e0.record()
                for _ in range(n_rep):
                    ev.evaluate_logits(x, out=res)
                e1.record()
                torch.cuda.synchronize()
                call_us[str(Bs)] = 1e3 * e0.elapsed_time(e1) / n_rep
                sweep[str(Bs)] = Bs / (e0.elapsed_time(e1) / n_rep * 1e-3)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in range(n_rep):
                        ev.evaluate_logits(x, out=res)
                g.replay()
                torch.cuda.synchronize()
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                dev_us[str(Bs)] = 1e3 * e0.elapsed_time(e1) / n_rep
                sweep_dev[str(Bs)] = Bs / (e0.elapsed_time(e1) / n_rep * 1e-3)
                del g
            out["device_sweep_samples_per_s"] = sweep
            out["device_sweep_graph_replay_samples_per_s"] = sweep_dev
            out["us_per_call"] = {"python_loop": call_us, "graph_replay": dev_us}
            # the conv + bitmask kernel is the dominant launch of the large-batch form: its image-stream roofline
            # (algorithmic bytes = the 23 of 32 image rows the stride-4 conv touches; ncu: dram bytes equal them)
            if w["image"] == 32 and w["grid"] in (9, 10, 11):
                algo = 65536 * 23 * 32 * 3 * 4
                out["roofline_b65536"] = {"bound": "hbm", "kernel": "q_conv_bits32_kernel", "algorithmic_bytes": algo,
                                          "whole_call_us": dev_us["65536"], "whole_call_gbs": algo / (dev_us["65536"] * 1e-6) / 1e9,
                                          "note": "whole call = conv + bitmask, tcgen05 accumulate + layer stack; the conv kernel "
                                                  "alone: profiles/r2_ncu_int_path_summary.txt"}
            del big
        # end to end through host buffers (pinned), copies inside the C call
        pinned = imgs.pin_memory().numpy()
        ev.evaluate_logits_host(pinned)
        t0 = time.perf_counter()
        for _ in range(5):
            hl, hd = ev.evaluate_logits_host(pinned)
        dt = (time.perf_counter() - t0) / 5
        out["e2e_host"] = {"value": B / dt, "unit": "samples/s", "h2d_bytes": int(pinned.nbytes),
                           "d2h_bytes": int(hl.nbytes + hd.nbytes)}
        # CPU: the reference engine on all host threads (bounded sample), checked bit-exact first
        n = min(B, 4096)
        threads = int_oracle.host_threads()
        kind = "reference" if int_oracle.RefEngine.available() else "port"
        cpu = int_oracle.RefEngine(path) if kind == "reference" else int_oracle.IntOracle(path)
        sample = imgs[:n].numpy()
        cl, cd = cpu.eval_batch(sample, threads=threads)
        out["bit_exact_vs_cpu"] = bool(np.array_equal(cl, hl[:n]) and np.array_equal(cd, hd[:n]))
        reps, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds / 3:
            cpu.eval_batch(sample, threads=threads)
            reps += 1
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": reps * n / dt, "unit": "samples/s", "cores": threads, "kind": kind,
                               "sample": f"{reps} x {n} images through evaluate_logits, one evaluator per thread"}
        t0 = time.perf_counter()
        cpu.eval_batch(sample[:512], threads=1)
        out["cpu_single_thread"] = {"value": 512 / (time.perf_counter() - t0), "unit": "samples/s", "cores": 1}
        try:
            out["incremental"] = incremental_leg(ev, cpu if kind == "reference" else None, device)
        except Exception as e:
            out["incremental"] = {"error": repr(e)}
    return out


def incremental_leg(ev, ref_engine, device, changed=8):
    """N3: the engine's incremental accumulator interface (evaluate_incremental: update_features + clipped ReLU + the
    single-score stack, nnue_engine.cpp:739-821) over S independent streams with device-resident CSR feature lists;
    every evaluation adds and removes `changed` features per stream.  Next to it the reference engine's own
    evaluate_incremental on one host thread (benchmark_engine.cpp:68-124 measures the same call)."""
    import numpy as np
    res = {"changed_features_per_eval": changed}
    rng = np.random.default_rng(5)
    F = ev.num_features
    for S in (1, 4096):
        base = [np.sort(rng.choice(F, size=F // 4, replace=False)).astype(np.int32) for _ in range(min(S, 64))]
        lists = [base[i % len(base)] for i in range(S)]
        off = np.zeros(S + 1, np.int32)
        np.cumsum([len(l) for l in lists], out=off[1:])
        d_off, d_idx = torch.from_numpy(off).to(device), torch.from_numpy(np.concatenate(lists)).to(device)
        ev.refresh_accumulator_csr(d_off, d_idx)
        # two alternating update sets: +A -B, then +B -A (the accumulators return to their start every second eval)
        a = torch.from_numpy(rng.integers(0, F, size=(S, changed)).astype(np.int32)).to(device).reshape(-1)
        b = torch.from_numpy(rng.integers(0, F, size=(S, changed)).astype(np.int32)).to(device).reshape(-1)
        uoff = torch.arange(0, (S + 1) * changed, changed, dtype=torch.int32, device=device)
        score = torch.empty(S, dtype=torch.float32, device=device)
        def one(i):
            if i % 2 == 0:
                ev.update_features_csr(uoff, a, uoff, b)
            else:
                ev.update_features_csr(uoff, b, uoff, a)
            ev.score_accumulators(out=score)
        for i in range(10):
            one(i)
        torch.cuda.synchronize()
        reps = 200
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(reps):
            one(i)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        # the same 200 evaluations replayed as one CUDA graph (no host launch cost)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(reps):
                one(i)
        g.replay()
        torch.cuda.synchronize()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        g.replay()
        e3.record()
        torch.cuda.synchronize()
        res[f"streams_{S}"] = {"evals_per_s": S * reps / wall, "us_per_eval_call": 1e6 * wall / reps,
                               "device_us_per_eval_call": 1e3 * e0.elapsed_time(e1) / reps,
                               "graph_replay_evals_per_s": S * reps / (e2.elapsed_time(e3) * 1e-3),
                               "graph_replay_us_per_eval_call": 1e3 * e2.elapsed_time(e3) / reps}
    if ref_engine is not None:  # the reference engine, one host thread: alternate two feature lists differing in `changed` features
        base = np.sort(rng.choice(F, size=F // 4, replace=False)).astype(np.int32)
        other = base.copy()
        other[:changed] = (other[:changed] + 1) % F
        ref_engine.mark_dirty()
        ref_engine.eval_incremental(base)
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 1.0:
            for _ in range(500):
                ref_engine.eval_incremental(other)
                ref_engine.eval_incremental(base)
            n += 1000
        dt = time.perf_counter() - t0
        res["cpu_reference_one_thread"] = {"evals_per_s": n / dt, "us_per_eval": 1e6 * dt / n, "kind": "reference",
                                           "note": "oracle/_ref evaluate_incremental through ctypes (the ~1 us ctypes call is included)"}
    return res


def touched_image_bytes(w, shape, B):
    """Bytes of the image rows a 3x3 / pad 1 / stride s conv actually reads: every row for s <= 3, otherwise three rows
    per raster row (whole rows: DRAM moves 32-byte sectors and the three taps of a cell lie s columns apart)."""
    H, s = w["image"], int(shape.stride)
    rows = set()
    for oy in range(int(shape.Gh)):
        for k in (-1, 0, 1):
            y = oy * s + k
            if 0 <= y < H:
                rows.add(y)
    return B * 3 * len(rows) * w["image"] * 4


def module_api_leg(model, sets, steps, global_batch):
    """The drop-in loop of INTEGRATION 2.1(a): loss = compute_loss(model, batch); loss.backward() -- gradients
    accumulated by autograd into .grad (nnue._NNUELoss)."""
    from nnue_vision_b200 import train
    for p in model.parameters():
        p.grad = None
    for i in range(3 * len(sets)):  # every input buffer is seen three times: eager, graph capture, replay
        loss = train.compute_loss(model, sets[i % len(sets)])
        loss.backward()
        for p in model.parameters():
            p.grad = None
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        loss = train.compute_loss(model, sets[i % len(sets)])
        loss.backward()
        for p in model.parameters():
            p.grad = None
    e1.record()
    torch.cuda.synchronize()
    host_ms = (time.perf_counter() - t0) * 1e3 / steps
    ms = e0.elapsed_time(e1) / steps
    B = sets[0][0].shape[0]
    return {"api": "module", "value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "host_ms_per_step": host_ms,
            "note": "loss = train.compute_loss(model, batch); loss.backward() on one GPU, device-resident inputs: the step replays a "
                    "CUDA graph into a private flat buffer, one copy hands the gradients to the autograd node, one kernel applies "
                    "the upstream factor; autograd accumulates into .grad"}


def exchange_check(dp, world, device):
    """Untimed: the exchange the timed loop uses against one NCCL all-reduce of the same per-rank values, and the
    same result on every rank (the driver's GPU test tier has one GPU; this runs wherever bench.py runs with N > 1)."""
    if world == 1:
        return None
    n = dp.buf.numel()
    saved = dp.buf.flat.clone()
    g = torch.Generator(device=device).manual_seed(99 + torch.distributed.get_rank())
    vals = torch.randn(n, generator=g, device=device)
    ref = vals.clone()
    torch.distributed.all_reduce(ref)
    dp.buf.flat.copy_(vals)
    dp._exchange()
    torch.cuda.synchronize()
    got = dp.buf.flat.clone()
    dp.buf.flat.copy_(saved)
    err = float((got - ref).abs().max())
    scale = float(ref.abs().max())
    # bit-identical on every rank: compare against rank 0's copy
    r0 = got.clone()
    torch.distributed.broadcast(r0, src=0)
    same = torch.tensor([1 if torch.equal(r0, got) else 0], device=device)
    torch.distributed.all_reduce(same, op=torch.distributed.ReduceOp.MIN)
    return {"exchange": dp.allreduce, "max_abs_diff_vs_nccl": err, "max_abs_ref": scale,
            "within_1e-6_rel": bool(err <= 1e-6 * scale), "identical_on_all_ranks": bool(int(same.item()))}


def run_b200(args):
    from nnue_vision_b200 import _lib, train
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the NNUE hot path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    clocks = ClockSampler(local_rank).__enter__()  # started now: nvidia-smi takes a moment to come up
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=device)
    w = dict(WORKLOADS[args.workload])
    if args.batch:
        w["batch"] = args.batch
    if args.strong:
        w["batch"] = max(1, w["batch"] // world)
    B = w["batch"]
    for kv in args.opt:
        key, val = kv.split("=")
        _lib.set_option(key, int(val))
    model = build_model(w, device)
    dp = train.DataParallelStep(model, single=True) if args.exchange == "none" else train.DataParallelStep(model, allreduce=args.exchange)
    if world > 1:  # identical replicas
        for p in model.parameters():
            torch.distributed.broadcast(p.data, src=0)

    # input sets cycled so that every step reads images that are not in the 126 MB L2: three sets at CIFAR shape
    # (3 x B x 12 KB = 600 MB), one set when a single set is already far larger than L2 (ImageNet shapes: GBs)
    img_bytes = B * 3 * w["image"] * w["image"] * 4
    n_sets = 3 if img_bytes < (1 << 30) else 1
    sets = [synthetic_batch(w, B, seed=1000 * rank + i, device=device) for i in range(n_sets)]
    global_batch = B * world

    # DataParallelStep replays a CUDA graph per input buffer from the second time it sees the buffer on: prime every
    # set twice (untimed, before the warm-up steps) so that no capture falls into the warm-up or the timed region
    for _ in range(2):
        for imgs, labs in sets:
            dp.step(imgs, labs, global_batch=global_batch)
    xcheck = exchange_check(dp, world, device) if args.exchange != "none" else None
    clocks.begin()
    for i in range(max(args.warmup, 3)):
        dp.step(*sets[i % n_sets], global_batch=global_batch)
    windows, launches, own = [], 0, []
    for wdw in range(max(1, args.windows)):
        barrier(world)
        _lib.lib().nnue_launch_count(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            loss = dp.step(*sets[i % n_sets], global_batch=global_batch)
        e1.record()
        barrier(world)
        launches = int(_lib.lib().nnue_launch_count(0))
        own.append(e0.elapsed_time(e1))
        windows.append(max_over_ranks(e0.elapsed_time(e1), world, device))
    ms_total = sorted(windows)[len(windows) // 2]
    global RANK_MS
    RANK_MS = per_rank(sorted(own)[len(own) // 2] / args.steps, world, device)  # every rank's own median step time
    value = args.steps * global_batch / (ms_total * 1e-3)
    final_loss = float(loss)

    e2e = None
    if not args.no_e2e:
        e2e = e2e_leg(args, w, B, world, rank, device, dp, sets, global_batch, img_bytes)
    clocks.end()
    clocks.__exit__()
    return finish_line(args, w, B, world, rank, device, dp, model, sets, global_batch, img_bytes, n_sets, value, ms_total, windows,
                       launches, final_loss, xcheck, clocks, e2e)


def e2e_leg(args, w, B, world, rank, device, dp, sets, global_batch, img_bytes):
    # ---- e2e: the user-facing call with HOST buffers; H2D of the step's inputs and D2H of the loss timed
    n_host = 2 if img_bytes < (1 << 30) else 1
    host_sets = [synthetic_batch(w, B, seed=2000 * rank + i, pin=True) for i in range(n_host)]
    dev_img = [torch.empty_like(sets[0][0]) for _ in range(2)]
    dev_lab = [torch.empty_like(sets[0][1]) for _ in range(2)]
    copy_stream = torch.cuda.Stream()

    def e2e_steps(n):
        # double-buffered: the copy of step i+1 overlaps the kernels of step i; every step still
        # moves its own inputs host->device and reads its loss back
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        done = [torch.cuda.Event(), torch.cuda.Event()]
        with torch.cuda.stream(copy_stream):
            dev_img[0].copy_(host_sets[0][0], non_blocking=True)
            dev_lab[0].copy_(host_sets[0][1], non_blocking=True)
            ready[0].record()
        losses = []
        for i in range(n):
            cur, nxt = i % 2, (i + 1) % 2
            if i + 1 < n:
                with torch.cuda.stream(copy_stream):
                    if i >= 1:
                        copy_stream.wait_event(done[nxt])
                    dev_img[nxt].copy_(host_sets[nxt % n_host][0], non_blocking=True)
                    dev_lab[nxt].copy_(host_sets[nxt % n_host][1], non_blocking=True)
                    ready[nxt].record()
            torch.cuda.current_stream().wait_event(ready[cur])
            l = dp.step(dev_img[cur], dev_lab[cur], global_batch=global_batch)
            done[cur].record()
            losses.append(float(l))  # D2H read of the step's result
        return losses

    e2e_n = args.steps if img_bytes < (1 << 30) else max(3, min(args.steps, 6))
    e2e_steps(6 if img_bytes < (1 << 30) else 4)  # (both staging buffers seen twice: their graphs exist before the timed region)
    barrier(world)
    t0 = time.perf_counter()
    e2e_steps(e2e_n)
    barrier(world)
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3, world, device)
    e2e_value = e2e_n * global_batch / (e2e_ms * 1e-3)
    h2d = int(host_sets[0][0].numel() * 4 + host_sets[0][1].numel() * 8) * world
    return {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * world,
            "ms_per_step": e2e_ms / e2e_n, "steps": e2e_n,
            "note": "DataParallelStep.step on pinned host batches, H2D double-buffered against compute, loss read back every step"}


def finish_line(args, w, B, world, rank, device, dp, model, sets, global_batch, img_bytes, n_sets, value, ms_total, windows,
                launches, final_loss, xcheck, clocks, e2e):
    from nnue_vision_b200 import _lib

    # ---- the module API (compute_loss + backward, eager)
    module_api = None
    if world == 1 and not args.no_module_api:
        module_api = module_api_leg(model, sets, min(args.steps, 20), global_batch)
        dp.buf.attach()

    # ---- per-stage device times and the roofline of the dominant kernel
    stages = stage_breakdown(dp, *sets[0])
    shape, bits = model.extract_bits(sets[0][0])
    nnz_total = int(sum(int(torch.bitwise_and(bits >> k, 1).sum()) for k in range(32)))
    del bits
    L1, F = w["L1"], shape.F
    img_algo = touched_image_bytes(w, shape, B)
    row_bytes = B * shape.PP * 4  # one fp32 value per (sample, padded position): g_bin, stored activations
    sp = ctypes.byref(shape)
    L = _lib.lib()
    dense_in = bool(L.nnue_input_bwd_is_dense(sp))
    umma = bool(L.nnue_ft_uses_umma(sp))
    mma = bool(L.nnue_ft_uses_mma(sp))
    # algorithmic bytes per launch
    algo = {
        "extract_fwd": img_algo + B * shape.NW * 4 + row_bytes,  # images -> bitmask + stored activations
        "ft_fwd": nnz_total * L1 * 4 + B * L1 * 4 + nnz_total * 4,        # row reads + out + indices
        "ft_bwd_dw": nnz_total * L1 * 4 + F * L1 * 4,                      # g_ft row per active pair + dW
        "ft_bwd_gbin": nnz_total * L1 * 4 + B * L1 * 4 + row_bytes,       # table row per active pair + g_ft + g_bin
        "ft_bwd": 2 * nnz_total * L1 * 4 + F * L1 * 4 + B * L1 * 4 + row_bytes,
        "conv_bwd": img_algo + 2 * row_bytes,                             # images + g_bin + stored activations
        # general input gradient: table rows, g_bin written and read back, stored activations read, touched image rows read
        "input_bwd": nnz_total * L1 * 4 + B * L1 * 4 + img_algo + 3 * row_bytes,
        "head_train": 2 * B * L1 * 4,
    }
    # tensor-core work actually issued by the bf16-split contractions (2 * M * N * K * number of term products)
    mma_flops = {"ft_fwd": 2.0 * B * shape.PP * L1 * 3, "ft_bwd_dw": 2.0 * B * shape.PP * L1 * 3,
                 "ft_bwd_gbin": 2.0 * B * shape.PP * L1 * 6, "input_bwd": 2.0 * B * shape.PP * L1 * 6}
    if L.nnue_head_uses_umma(sp):  # layer 1 of wide stacks: forward, input gradient, weight gradient, six term pairs each
        mma_flops["head_train"] = 3 * 2.0 * B * L1 * w["L2"] * 6
    kernel_of = {"extract_fwd": "extract_fwd_fixed_kernel" if w["C"] in (4, 8, 16, 32) else "extract_fwd_kernel",
                 "ft_fwd": "ft_fwd_mma_kernel", "head_train": "head_train_kernel" if L.nnue_head_is_fused(sp) else "ugemm_kernel",
                 "ft_bwd_dw": "ft_bwd_dw_mma_kernel", "ft_bwd_gbin": "ft_bwd_gbin_mma_kernel", "conv_bwd": "conv_bwd_kernel",
                 "input_bwd": "extract_bwd_kernel"}
    if umma:  # tcgen05 / TMEM contractions (ft_umma.cu)
        kernel_of.update({"ft_fwd": "ft_bitgemm_umma_kernel<0>", "ft_bwd_dw": "ft_bitgemm_umma_kernel<1>",
                          "ft_bwd_gbin": "ft_gbin_umma_kernel", "input_bwd": "ft_gbin_umma_kernel"})
    elif not mma:
        kernel_of.update({"ft_fwd": "ft_gather_fwd_kernel", "ft_bwd_dw": "ft_bwd_dw_kernel", "input_bwd": "ft_gather_dval_kernel"})
    # dram bytes per launch from the committed ncu --set full captures, keyed by workload then kernel; null when this
    # workload has no capture (a capture of another workload says nothing about this one)
    traffic = {}
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get(args.workload, {})
        except Exception:
            traffic = {}
    peak, peak_src = peaks()
    tpeak, tpeak_src = tensor_peak()
    roofs = {}
    for k, nbytes in algo.items():
        if k in stages and stages[k] > 0:
            kern = kernel_of.get(k)
            tr = traffic.get(kern.split("<")[0]) if kern else None
            ach = nbytes / (stages[k] * 1e-3) / 1e9
            hbm = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": tr,
                   "ms": stages[k], "algorithmic_bytes": nbytes, "kernel": kern, "peak_source": peak_src}
            roofs[k] = hbm
            if k in mma_flops and (mma or k == "head_train"):
                tf = mma_flops[k] / (stages[k] * 1e-3) / 1e12
                tens = {"bound": "tensor", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak, "traffic": tr,
                        "ms": stages[k], "algorithmic_flops": mma_flops[k], "kernel": kern, "peak_source": tpeak_src,
                        "note": ("bf16 tcgen05.mma (UMMA, TMEM accumulators)" if umma or k == "head_train" else "bf16 mma.sync") +
                                ": flops issued = 2 M N K x 3 (6) exact split-term products per fp32 product"}
                # a contraction against a table that does not stream from HBM per sample is bound by the tensor pipe
                roofs[k] = dict(tens, hbm_view={kk: hbm[kk] for kk in ("achieved", "frac", "algorithmic_bytes")})
    dominant = max(roofs, key=lambda k: roofs[k]["ms"]) if roofs else None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "windows_ms": windows, "windows_note": f"{len(windows)} windows of {args.steps} steps each; value and ms_per_step are the median window",
        "config": {"workload": args.workload, "per_gpu_batch": B, "global_batch": global_batch,
                   "arch": {k: w[k] for k in ("grid", "C", "L1", "L2", "L3", "NC", "input", "image")},
                   "parallelism": f"dp{world}", "exchange": dp.allreduce, "nnz_per_sample": nnz_total / B,
                   "ft_form": "dense (tcgen05 bit-GEMM)" if umma else ("dense (mma.sync)" if mma else "gather"),
                   "l2_policy": "inputs larger than L2: %d image set(s) x %.0f MB cycled" % (n_sets, img_bytes / 1e6),
                   "loss_after_timed_steps": final_loss},
        "clocks": clocks.summary(),
        "e2e": e2e,
        "gpu_launches": launches,
        "stages_ms": stages,
    }
    if xcheck is not None:
        line["exchange_check"] = xcheck
    if world > 1 and RANK_MS is not None:
        line["ms_per_step_by_rank"] = RANK_MS  # each rank's own device time per step (median window): the spread is the skew
    if args.exchange == "none" and world > 1:
        line["diagnostic"] = "no gradient exchange (--exchange none): NOT a training step; compare ms_per_step with the exchanging run"
    if module_api is not None:
        line["module_api"] = module_api
    if dominant:
        line["roofline"] = dict(roofs[dominant], stage=dominant,
                                note="stage = one C-ABI call = the named kernel + its small helper kernels; algorithmic bytes per "
                                     "SURVEY 8d (image bytes = rows the conv touches); traffic = ncu dram bytes per launch of the "
                                     "named kernel from the committed capture of THIS workload (profiles/traffic.json), else null")
        line["roofline_all"] = roofs
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_training(w, args.cpu_seconds)
        if not args.no_int and w["image"] <= 64:
            try:
                line["int_inference"] = int_inference_leg(w, device, args.cpu_seconds)
            except Exception as e:  # the side measurement must never lose the headline line
                line["int_inference"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
