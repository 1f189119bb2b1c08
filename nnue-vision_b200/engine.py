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
        # incremental-accumulator state (nnue_engine.h:583-594), over S independent streams
        self._acc = self._backup = None
        self._last_features = None
        self._dirty, self._incremental = True, True
        self._ws_bytes = {}  # scratch size of the large-batch form per batch size
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
        # the tables are uploaded to the CURRENT device; evaluate_* checks its inputs against it
        self.device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
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
        self._acc = self._backup = self._last_features = None
        self._dirty = True
        self._ws_bytes = {}
        return True

    def _require(self, t=None):
        if not self._h.value:
            raise _lib.NnueError("no model loaded")
        if t is not None and torch.is_tensor(t) and t.is_cuda and t.device != self.device:
            raise _lib.NnueError(f"the model was loaded on {self.device} but the input lives on {t.device}: load one "
                                 "evaluator per device (inference runs per GPU, no peer traffic)")

    def evaluate_logits(self, images: torch.Tensor, layer_stack_index: int = 0, out=None):
        """images: CUDA float32 [B, H, W, 3] -- the raw buffer the engine would be handed, read as HWC
        (callers holding CHW tensors pass `chw.contiguous().view(B, H, W, 3)`, the byte
        reinterpretation evaluate.py:154-168 performs).  Returns (logits [B, NC], density [B]).
        `out=(logits, density)` reuses the caller's result tensors (a latency-bound loop over single images
        then allocates nothing per call)."""
        self._require(images)
        if images.dim() != 4 or images.shape[-1] != 3:
            raise ValueError(f"expected images [B,H,W,3], got {tuple(images.shape)}")
        B, H, W, _ = images.shape
        dev = images.device
        guard = None
        if dev.index != torch.cuda.current_device():
            guard = _lib.on_device_of(images)
            guard.__enter__()
        try:
            if out is None:
                logits = torch.empty((B, self.num_classes), dtype=torch.float32, device=dev)
                density = torch.empty((B,), dtype=torch.float32, device=dev)
            else:
                logits, density = out
                if tuple(logits.shape) != (B, self.num_classes) or tuple(density.shape) != (B,) or logits.device != dev \
                        or density.device != dev:
                    raise ValueError("out=(logits [B, NC], density [B]) on the images' device")
            # scratch of the large-batch (tensor-core) form comes from the caller: nothing is allocated or mutated
            # inside the C call, so several streams may share one evaluator's tables
            ws_bytes = self._ws_bytes.get(B)
            if ws_bytes is None:
                ws_bytes = self._ws_bytes[B] = int(_lib.lib().nnue_q_workspace_bytes(self._h, B))
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev) if ws_bytes else None
            check(_lib.lib().nnue_q_infer_ws(self._h, dptr(images, torch.float32), B, H, W, int(layer_stack_index),
                                             dptr(logits, torch.float32), dptr(density, torch.float32), dptr(ws), ws_bytes,
                                             stream_ptr()))
        finally:
            if guard is not None:
                guard.__exit__(None, None, None)
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

    # ---- incremental accumulators (NNUEEvaluator::refresh_accumulator / update_features / evaluate_incremental /
    #      save_ / restore_accumulator, nnue_engine.cpp:739-821), batched over S independent streams ---------------
    # A flat list of feature indices addresses ONE stream (the reference's interface, a Python float comes back);
    # a list of lists addresses S streams at once (a CUDA tensor [S] comes back).
    @staticmethod
    def _streams(features):
        single = len(features) == 0 or not isinstance(features[0], (list, tuple, np.ndarray))
        lists = [list(map(int, features))] if single else [list(map(int, f)) for f in features]
        return single, lists

    @staticmethod
    def _csr(lists, device):
        off = np.zeros(len(lists) + 1, np.int32)
        np.cumsum([len(l) for l in lists], out=off[1:])
        idx = np.fromiter((f for l in lists for f in l), np.int32, count=int(off[-1])) if off[-1] else np.zeros(1, np.int32)
        return torch.from_numpy(off).to(device), torch.from_numpy(idx).to(device)

    def _device(self):
        return self.device  # the device the tables were loaded on

    def refresh_accumulator(self, features):
        """acc = (int16)bias + sum of the listed rows, for every stream (nnue_engine.cpp:804-815)."""
        self._require()
        _, lists = self._streams(features)
        dev = self._device()
        S = len(lists)
        if self._acc is None or self._acc.shape[0] != S or self._acc.device != dev:
            self._acc = torch.empty((S, self.l1_size), dtype=torch.int16, device=dev)
        off, idx = self._csr(lists, dev)
        with torch.cuda.device(dev):
            check(_lib.lib().nnue_q_acc_apply(self._h, S, 1, dptr(off), dptr(idx), None, None, dptr(self._acc), stream_ptr()))

    def update_features(self, added, removed):
        """acc -= rows(removed); acc += rows(added) with int16 wrap-around (nnue_engine.cpp:818-821)."""
        self._require()
        if self._acc is None:
            raise _lib.NnueError("update_features before refresh_accumulator")
        _, la = self._streams(added)
        _, lr = self._streams(removed)
        S = self._acc.shape[0]
        if len(la) != S or len(lr) != S:
            raise ValueError(f"expected feature lists for {S} streams")
        dev = self._acc.device
        ao, ai = self._csr(la, dev)
        ro, ri = self._csr(lr, dev)
        with torch.cuda.device(dev):
            check(_lib.lib().nnue_q_acc_apply(self._h, S, 0, dptr(ao), dptr(ai), dptr(ro), dptr(ri), dptr(self._acc), stream_ptr()))

    def evaluate_incremental(self, current_features, layer_stack_index: int = 0):
        """The engine's chess-style entry point (nnue_engine.cpp:739-787): refresh when dirty / disabled, otherwise
        apply the difference to the previous call's features; then clipped ReLU + the single-score layer stack."""
        self._require()
        single, cur = self._streams(current_features)
        stale = (not self._incremental or self._dirty or self._acc is None or self._last_features is None
                 or len(self._last_features) != len(cur))
        if stale:
            self.refresh_accumulator(cur)
            self._last_features, self._dirty = cur, False
        else:
            added, removed = [], []
            for last, now in zip(self._last_features, cur):
                sl, sn = set(last), set(now)
                removed.append([f for f in last if f not in sn])  # std::find per element: multiplicities are kept
                added.append([f for f in now if f not in sl])
            if any(added) or any(removed):
                self.update_features(added, removed)
                self._last_features = cur
        S = self._acc.shape[0]
        score = torch.empty((S,), dtype=torch.float32, device=self._acc.device)
        with torch.cuda.device(self._acc.device):
            check(_lib.lib().nnue_q_acc_score(self._h, S, dptr(self._acc), int(layer_stack_index), dptr(score), stream_ptr()))
        return float(score[0]) if single else score

    # ---- the same three steps on DEVICE-RESIDENT CSR lists (int32 offsets [S + 1], int32 indices [nnz]): no Python list
    #      handling, no host->device copy per call -- the form a video pipeline that extracts features on the GPU uses ----
    def refresh_accumulator_csr(self, off: torch.Tensor, idx: torch.Tensor):
        self._require(off)
        S = off.numel() - 1
        if self._acc is None or self._acc.shape[0] != S or self._acc.device != off.device:
            self._acc = torch.empty((S, self.l1_size), dtype=torch.int16, device=off.device)
        with torch.cuda.device(off.device):
            check(_lib.lib().nnue_q_acc_apply(self._h, S, 1, dptr(off, torch.int32), dptr(idx, torch.int32), None, None,
                                              dptr(self._acc), stream_ptr()))

    def update_features_csr(self, add_off, add_idx, rem_off, rem_idx):
        self._require(add_off)
        if self._acc is None or self._acc.shape[0] != add_off.numel() - 1:
            raise _lib.NnueError("update_features_csr before refresh_accumulator_csr (or a different stream count)")
        with torch.cuda.device(self._acc.device):
            check(_lib.lib().nnue_q_acc_apply(self._h, self._acc.shape[0], 0, dptr(add_off, torch.int32), dptr(add_idx, torch.int32),
                                              dptr(rem_off, torch.int32), dptr(rem_idx, torch.int32), dptr(self._acc), stream_ptr()))

    def score_accumulators(self, layer_stack_index: int = 0, out: torch.Tensor = None):
        """clipped ReLU + the single-score layer stack over the current accumulators -> CUDA tensor [S]."""
        self._require()
        S = self._acc.shape[0]
        score = out if out is not None else torch.empty((S,), dtype=torch.float32, device=self._acc.device)
        with torch.cuda.device(self._acc.device):
            check(_lib.lib().nnue_q_acc_score(self._h, S, dptr(self._acc), int(layer_stack_index), dptr(score), stream_ptr()))
        return score

    def save_accumulator(self):
        if self._acc is not None:
            self._backup = self._acc.clone()

    def restore_accumulator(self):
        if self._backup is not None and self._acc is not None and self._backup.shape == self._acc.shape:
            self._acc.copy_(self._backup)

    def enable_incremental(self, enable: bool = True):
        self._incremental = bool(enable)

    def mark_dirty(self):
        self._dirty = True

    def get_accumulator(self):
        """The int16 accumulators [S][L1] (device tensor; for inspection and tests)."""
        return self._acc
