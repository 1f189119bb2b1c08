"""oracle/int_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes bindings for the two CPU checkers of the quantized integer path:

* ``IntOracle``  -- oracle/nnue_int_oracle.c, the plain-C restatement (kind "port").
* ``RefEngine``  -- oracle/_ref/libnnue_ref.so, the reference's own C++ engine
  compiled from /root/reference/engine/src (kind "reference"); present only
  where `make -C oracle` ran with the reference checkout available (the built
  .so travels to the GPU box with the snapshot).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this module.  Parity status: pinned (tests/test_oracle_int.py).
"""
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_ORACLE_SO = _HERE / "_build" / "liboracle_int.so"
_REF_SO = _HERE / "_ref" / "libnnue_ref.so"


def build(quiet=True):
    """Compile the C restatement (and the reference engine when its sources exist)."""
    subprocess.run(["make", "-C", str(_HERE)], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class IntOracle:
    """The C restatement of serialize.py + engine/nnue_inference.cpp semantics."""

    def __init__(self, nnue_path):
        if not _ORACLE_SO.exists():
            build()
        lib = ctypes.CDLL(str(_ORACLE_SO))
        lib.oracle_load.restype = ctypes.c_void_p
        lib.oracle_load.argtypes = [ctypes.c_char_p]
        lib.oracle_free.argtypes = [ctypes.c_void_p]
        lib.oracle_dims.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.oracle_threshold.restype = ctypes.c_float
        lib.oracle_threshold.argtypes = [ctypes.c_void_p]
        lib.oracle_stride.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.oracle_eval.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.oracle_eval_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        self._lib = lib
        self._h = lib.oracle_load(str(nnue_path).encode())
        if not self._h:
            raise ValueError(f"oracle: cannot parse {nnue_path}")
        d = np.zeros(8, np.int32)
        lib.oracle_dims(self._h, _ptr(d))
        self.F, self.L1, self.L2, self.L3, self.NC, self.OC, self.G, self.n_buckets = (int(x) for x in d)
        self.threshold = float(lib.oracle_threshold(self._h))

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.oracle_free(self._h)
            self._h = None

    def stride(self, H):
        return int(self._lib.oracle_stride(self._h, H))

    def eval_one(self, img_hwc):
        """One image -> (logits, n_active, conv_buf int8[F], acc int16[L1])."""
        img = np.ascontiguousarray(img_hwc, np.float32)
        H, W = img.shape[0], img.shape[1]
        logits = np.zeros(self.NC, np.float32)
        conv = np.zeros(self.F, np.int8)
        acc = np.zeros(self.L1, np.int16)
        n = self._lib.oracle_eval(self._h, _ptr(img), H, W, _ptr(logits), _ptr(conv), _ptr(acc))
        if n < 0:
            raise ValueError("oracle: conv raster exceeds the feature buffer")
        return logits, n, conv, acc

    def eval_batch(self, imgs_bhwc, threads=1):
        """imgs: float32 [B,H,W,3] (the raw buffer the engine sees) -> (logits [B,NC], density [B])."""
        imgs = np.ascontiguousarray(imgs_bhwc, np.float32)
        B, H, W = imgs.shape[0], imgs.shape[1], imgs.shape[2]
        logits = np.zeros((B, self.NC), np.float32)
        dens = np.zeros(B, np.float32)
        threads = max(1, min(threads, B))
        bounds = np.linspace(0, B, threads + 1).astype(int)

        def run(i):
            return self._lib.oracle_eval_batch(self._h, _ptr(imgs), int(bounds[i]), int(bounds[i + 1]), H, W,
                                               _ptr(logits), _ptr(dens))

        if threads == 1:
            rcs = [run(0)]
        else:
            with ThreadPoolExecutor(threads) as ex:  # ctypes drops the GIL during the call
                rcs = list(ex.map(run, range(threads)))
        if any(rcs):
            raise ValueError("oracle: conv raster exceeds the feature buffer")
        return logits, dens


class RefEngine:
    """The reference C++ engine itself (engine/src/nnue_engine.cpp), batched in-process."""

    @staticmethod
    def available():
        return _REF_SO.exists()

    def __init__(self, nnue_path):
        lib = ctypes.CDLL(str(_REF_SO))
        lib.ref_load.restype = ctypes.c_void_p
        lib.ref_load.argtypes = [ctypes.c_char_p]
        lib.ref_free.argtypes = [ctypes.c_void_p]
        lib.ref_dims.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.ref_num_classes.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        lib.ref_eval_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        self._lib = lib
        self._h = lib.ref_load(str(nnue_path).encode())
        if not self._h:
            raise ValueError(f"reference engine: cannot load {nnue_path}")
        d = np.zeros(8, np.int32)
        lib.ref_dims(self._h, _ptr(d))
        self.F, self.L1, self.L2, self.L3, _, self.OC, self.G, self.n_buckets = (int(x) for x in d)
        self._nc = None

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.ref_free(self._h)
            self._h = None

    def eval_batch(self, imgs_bhwc, threads=1):
        imgs = np.ascontiguousarray(imgs_bhwc, np.float32)
        B, H, W = imgs.shape[0], imgs.shape[1], imgs.shape[2]
        if self._nc is None:
            self._nc = int(self._lib.ref_num_classes(self._h, H, W))
        logits = np.zeros((B, self._nc), np.float32)
        dens = np.zeros(B, np.float32)
        rc = self._lib.ref_eval_batch(self._h, _ptr(imgs), B, H, W, _ptr(logits), _ptr(dens), int(threads))
        if rc < 0:
            raise RuntimeError(f"reference engine failed rc={rc}")
        return logits, dens


    # ---- incremental surface of the reference engine (ref_driver.cpp: evaluate_incremental / mark_dirty) ----
    def eval_incremental(self, features):
        f = np.ascontiguousarray(features, np.int32)
        self._lib.ref_eval_incremental.restype = ctypes.c_float
        self._lib.ref_eval_incremental.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        return float(self._lib.ref_eval_incremental(self._h, _ptr(f) if len(f) else None, int(len(f))))

    def mark_dirty(self):
        self._lib.ref_mark_dirty.argtypes = [ctypes.c_void_p]
        self._lib.ref_mark_dirty(self._h)


def legacy_score(q, features, bucket=0):
    """numpy restatement of NNUEEvaluator::evaluate_incremental after a full refresh: accumulator
    (nnue_engine.cpp:804-815, simd_scalar.cpp:97-104), clipped ReLU (:775-779) and the single-score
    LayerStack::forward (:382-478).  `q` is the dict of nnue_vision_b200.serialize.read_nnue (file parsing only)."""
    md, ft, st = q["metadata"], q["feature_transformer"], q["layer_stacks"][bucket]
    F, L1, L2 = md["num_features"], md["L1"], md["L2"]
    acc = ft["bias"].astype(np.int64)
    acc = acc.astype(np.int16).astype(np.int64)
    for f in features:
        if 0 <= f < F:
            acc = acc + ft["weight"][f].astype(np.int64)
    acc = ((acc + 32768) % 65536 - 32768)                      # int16 wrap-around of the running sums
    x = np.clip(acc, 0, int(np.int16(md["quantized_one"])))     # clipped ReLU

    def dense(w, b, v, scale):                                  # simd_scalar.cpp:116-136 (float divide, truncate)
        a = b.astype(np.int64) + w.astype(np.int64) @ v
        return np.clip(np.trunc(a.astype(np.float32) / np.float32(scale)).astype(np.int64), 0, 127)

    def dense_avx2(w, b, v, scale):
        # simd_avx2.cpp:114-152, the form LayerStack::forward takes on every AVX2 host (nnue_engine.cpp:393-397,
        # 453-457): the accumulator vector starts as set1_epi32(bias) and ALL EIGHT lanes are summed, so the bias
        # counts eight times; the quotient is an integer division
        a = 8 * b.astype(np.int64) + w.astype(np.int64) @ v
        q_ = np.abs(a) // int(scale) * np.sign(a)                # C++ division truncates toward zero
        return np.clip(q_, 0, 127)

    comb = dense_avx2(st["l1_weight"], st["l1_bias"][: L2 + 1], x, st["l1_scale"])
    fact = dense(st["l1_fact_weight"][L2: L2 + 1], st["l1_fact_bias"][L2: L2 + 1], x, st["l1_fact_scale"])
    l1c = np.float32(comb[L2]) / np.float32(st["l1_scale"])
    l1f = np.float32(fact[0]) / np.float32(st["l1_fact_scale"])
    c = comb[:L2]
    expd = np.concatenate([np.clip((c * c * 127) // 128, 0, 127), c])
    l2o = dense_avx2(st["l2_weight"], st["l2_bias"], expd, st["l2_scale"])
    a = int(st["output_bias"][0]) + int(st["output_weight"][0].astype(np.int64) @ l2o)
    l3c = np.float32(a) / np.float32(st["output_scale"])
    return float(np.float32(np.float32(l3c + l1f) + l1c))


def host_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
