"""Batched quantized inference -- the GPU stand-in for the reference's C++ `NNUEEvaluator`
(engine/include/nnue_engine.h:541-639) and its `nnue_inference` CLI (engine/nnue_inference.cpp),
bit-exact against them.  One fused kernel evaluates a whole batch; the reference evaluates one
image per process (evaluate.py:153-173)."""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check, dptr, stream_ptr


class NNUEEvaluator:
    """load_model / evaluate_logits with the reference's meaning, over batches.

    Like the reference class it keeps per-instance scratch and is not thread-safe: use one
    evaluator per host thread (nnue_engine.h:559-570)."""

    def __init__(self, path=None):
        self._h = ctypes.c_void_p()
        self.num_features = self.l1_size = self.l2_size = self.l3_size = 0
        self.num_classes = self.num_channels_per_square = self.grid_size = self.num_layer_stacks = 0
        self.visual_threshold = 0.0
        if path is not None and not self.load_model(path):
            raise _lib.NnueError(f"cannot load {path}")

    def __del__(self):
        self.close()

    def close(self):
        if getattr(self, "_h", None) and self._h.value and _lib is not None and _lib._lib is not None:
            _lib.lib().nnue_q_free(self._h)  # (module globals are already gone at interpreter shutdown)
            self._h = ctypes.c_void_p()

    def load_model(self, path) -> bool:
        """NNUEEvaluator::load_model (nnue_engine.cpp:544-657): False on a missing or malformed file."""
        self.close()
        h = ctypes.c_void_p()
        rc = _lib.lib().nnue_q_load(str(path).encode(), ctypes.byref(h))
        if rc in (-4, -5):  # NNUE_ERR_IO / NNUE_ERR_FORMAT: the engine reports these as `false`
            return False
        check(rc)
        self._h = h
        dims = (ctypes.c_int32 * 8)()
        thr = ctypes.c_float()
        check(_lib.lib().nnue_q_dims(self._h, dims, ctypes.byref(thr)))
        (self.num_features, self.l1_size, self.l2_size, self.l3_size, self.num_classes,
         self.num_channels_per_square, self.grid_size, self.num_layer_stacks) = (int(v) for v in dims)
        self.visual_threshold = float(thr.value)
        return True

    def _require(self):
        if not self._h.value:
            raise _lib.NnueError("no model loaded")

    def evaluate_logits(self, images: torch.Tensor, layer_stack_index: int = 0):
        """images: CUDA float32 [B, H, W, 3] -- the raw buffer the engine would be handed, read as HWC
        (callers holding CHW tensors pass `chw.contiguous().view(B, H, W, 3)`, the byte
        reinterpretation evaluate.py:154-168 performs).  Returns (logits [B, NC], density [B])."""
        self._require()
        if images.dim() != 4 or images.shape[-1] != 3:
            raise ValueError(f"expected images [B,H,W,3], got {tuple(images.shape)}")
        B, H, W, _ = images.shape
        logits = torch.empty((B, self.num_classes), dtype=torch.float32, device=images.device)
        density = torch.empty((B,), dtype=torch.float32, device=images.device)
        check(_lib.lib().nnue_q_infer(self._h, dptr(images, torch.float32), B, H, W, int(layer_stack_index),
                                      dptr(logits), dptr(density), stream_ptr()))
        return logits, density

    def evaluate_logits_host(self, images: np.ndarray, layer_stack_index: int = 0):
        """Same through host memory: numpy float32 [B, H, W, 3] in, numpy out; copies and the
        stream synchronise happen inside the C call."""
        self._require()
        images = np.ascontiguousarray(images, np.float32)
        if images.ndim != 4 or images.shape[-1] != 3:
            raise ValueError(f"expected images [B,H,W,3], got {images.shape}")
        B, H, W, _ = images.shape
        logits = np.empty((B, self.num_classes), np.float32)
        density = np.empty((B,), np.float32)
        check(_lib.lib().nnue_q_infer_host(self._h, images.ctypes.data_as(ctypes.c_void_p), B, H, W,
                                           int(layer_stack_index), logits.ctypes.data_as(ctypes.c_void_p),
                                           density.ctypes.data_as(ctypes.c_void_p)))
        return logits, density
