"""Evaluation callers of the hot path, with the reference's signatures and metric names
(/root/reference/evaluate.py:23-87, 90-400).

`evaluate_compiled_model` is where the two implementations differ most: the reference serialises the
model, then spawns `engine/build/nnue_inference model.nnue img.bin H W` once PER SAMPLE and parses its
CSV line (evaluate.py:153-200).  Here the same `.nnue` file is loaded once into the batched integer
kernel (`engine.NNUEEvaluator`, bit-exact against that executable) and every batch of the loader is one
call.  Images are handed over exactly as the reference hands them: the CHW float bytes of each sample,
read by the engine as HWC (evaluate.py:154-168)."""
import tempfile
import time
from pathlib import Path
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from .engine import NNUEEvaluator
from .serialize import serialize_model


def compute_metrics(outputs: torch.Tensor, targets: torch.Tensor) -> Dict[str, float]:
    """acc / f1 / precision / recall (weighted, zero_division=0) as evaluate.py:23-59 computes them."""
    from sklearn.metrics import accuracy_score, f1_score, precision_score, recall_score
    outputs_np = outputs.detach().cpu().numpy()
    targets_np = targets.detach().cpu().numpy()
    if outputs_np.ndim == 1:
        outputs_np = outputs_np.reshape(-1, 1)
    targets_np = targets_np.reshape(-1)
    if outputs_np.shape[1] == 1:
        predictions = (outputs_np[:, 0] > 0.5).astype(int)
        truth = (targets_np > 0.5).astype(int)
    else:
        predictions = outputs_np.argmax(axis=1)
        truth = targets_np.astype(int)
    return {
        "acc": accuracy_score(truth, predictions),
        "f1": f1_score(truth, predictions, average="weighted", zero_division=0),
        "precision": precision_score(truth, predictions, average="weighted", zero_division=0),
        "recall": recall_score(truth, predictions, average="weighted", zero_division=0),
    }


def evaluate_model(model, loader, loss_fn=None, device: Optional[torch.device] = None) -> Tuple[float, Dict[str, float]]:
    """Float-path evaluation (evaluate.py:62-87): mean of the per-batch cross-entropies, then the metrics."""
    device = device if device is not None else next(model.parameters()).device
    total_loss, outputs, targets = 0.0, [], []
    n_batches = 0
    with torch.no_grad():
        for images, labels in loader:
            images, labels = images.to(device), labels.to(device)
            logits = model(images)
            total_loss += F.cross_entropy(logits, labels.long()).item()
            outputs.append(logits.cpu())
            targets.append(labels.cpu())
            n_batches += 1
    if not n_batches:
        raise RuntimeError("No outputs generated during model evaluation")
    return total_loss / n_batches, compute_metrics(torch.cat(outputs), torch.cat(targets))


def evaluate_compiled_model(model, loader, model_type: str = "nnue", device: Optional[torch.device] = None) -> Dict[str, float]:
    """Quantized-path evaluation (evaluate.py:90-400): serialise -> integer engine -> the same metrics dict,
    plus `ms_per_sample` (device time of the batched calls, copies included, divided by the samples)
    and `latent_density` (mean fraction of active features, nnue_inference.cpp:50-54)."""
    if model_type != "nnue":
        raise ValueError(f"Unknown model type: {model_type}")  # EtinyNet is outside this path
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    with tempfile.TemporaryDirectory() as td:
        path = Path(td) / "model.nnue"
        was_training = model.training
        serialize_model(model, path)  # like the reference: eval() + in-place weight clip (serialize.py:500-528)
        model.train(was_training)
        engine = NNUEEvaluator(path)
    outputs, targets, densities = [], [], []
    seconds, samples = 0.0, 0
    for images, labels in loader:
        if images.dim() != 4 or images.shape[1] != 3:
            raise ValueError(f"expected images [B,3,H,W], got {tuple(images.shape)}")
        B, _, H, W = images.shape
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        # the engine sees each sample's CHW bytes as an H x W x 3 buffer (no permute)
        buf = images.to(device=device, dtype=torch.float32).contiguous().view(B, H, W, 3)
        logits, density = engine.evaluate_logits(buf)
        logits, density = logits.cpu(), density.cpu()
        seconds += time.perf_counter() - t0
        samples += B
        num_classes = int(labels.max().item()) + 1 if labels.numel() else 1
        if num_classes > 2 and logits.shape[1] == 1:
            raise RuntimeError(f"Compiled NNUE produced shape {tuple(logits.shape)} for {num_classes}-class labels.")
        outputs.append(logits)
        targets.append(labels.cpu())
        densities.append(density)
    if not outputs:
        raise RuntimeError("No outputs generated during compiled model evaluation")
    metrics = compute_metrics(torch.cat(outputs), torch.cat(targets))
    metrics["ms_per_sample"] = seconds / samples * 1000.0 if samples else 0.0
    metrics["latent_density"] = float(torch.cat(densities).mean()) if densities else 0.0
    return metrics
