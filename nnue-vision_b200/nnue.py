"""Drop-in for the NNUE model of the reference's `nnue.py`, backed by sm_100a CUDA kernels.

Same names, constructor arguments, attributes, parameter registration order and `state_dict`
keys as /root/reference/nnue.py:447-738 (`NNUE`, `FeatureTransformer`, `SimpleClassifier`,
`GridFeatureSet`, `LossParams`, `StraightThroughBinary`), so reference checkpoints load both
ways and the reference's `serialize.serialize_model` accepts these modules unchanged.

What differs is how the hot path runs: `NNUE.forward` is ONE autograd node whose forward and
backward are hand-written kernels called through the C ABI of libnnue_b200.so
(include/nnue_b200.h) -- conv+threshold -> bitmask, feature-transformer gather-accumulate,
pairwise + dense head; backward = head gradients, sorted segment reduction for the table
gradient, value gradient back into the conv / threshold.  There is no CPU or eager-PyTorch
fallback: CPU tensors raise.
"""
import ctypes
import os
import weakref
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, dptr, stream_ptr

DEFAULT_L1 = 1024
DEFAULT_L2 = 128
DEFAULT_L3 = 32


@dataclass
class LossParams:
    """Kept for constructor compatibility (nnue.py:63-72); unused by the vision head."""
    in_offset: float = 270
    out_offset: float = 270
    in_scaling: float = 340
    out_scaling: float = 380
    start_lambda: float = 1.0
    end_lambda: float = 1.0
    pow_exp: float = 2.5
    qp_asymmetry: float = 0.0


@dataclass
class GridFeatureSet:
    """nnue.py:81-90."""
    grid_size: int = 10
    num_features_per_square: int = 8

    @property
    def num_features(self) -> int:
        return self.grid_size * self.grid_size * self.num_features_per_square


class StraightThroughBinary(torch.autograd.Function):
    """Hard threshold forward, identity / sigmoid-surrogate backward (nnue.py:15-59).

    Stand-alone form for callers that use it directly; inside `NNUE.forward` the same maths is
    fused into the extraction kernels."""

    SHARPNESS = 10.0

    @staticmethod
    def forward(ctx, input, threshold=0.0):
        threshold = torch.as_tensor(threshold, dtype=input.dtype, device=input.device)
        ctx.save_for_backward(input, threshold)
        return (input > threshold).float()

    @staticmethod
    def backward(ctx, grad_output):
        input, threshold = ctx.saved_tensors
        grad_threshold = None
        if ctx.needs_input_grad[1]:
            k = StraightThroughBinary.SHARPNESS
            sig = torch.sigmoid(k * (input - threshold))
            g = -(grad_output * (k * sig * (1 - sig)))
            dims = tuple(d for d in range(g.dim()) if threshold.dim() != g.dim() or threshold.shape[d] == 1)
            grad_threshold = g.sum(dim=dims, keepdim=threshold.dim() == g.dim()).reshape(threshold.shape)
        return grad_output, grad_threshold


def binary_activation_ste(x, threshold=0.0):
    return StraightThroughBinary.apply(x, threshold)


def _empty(shape, dtype, like):
    return torch.empty(shape, dtype=dtype, device=like.device)


class _FTIndexed(torch.autograd.Function):
    """FeatureTransformer.forward(idx, val) on explicit index lists (nnue.py:686-710)."""

    @staticmethod
    def forward(ctx, idx, val, weight, bias):
        if not weight.is_cuda:
            raise _lib.NnueError("FeatureTransformer runs on CUDA only (no CPU fallback); move the module to cuda")
        idx = idx.to(device=weight.device, dtype=torch.long).contiguous()
        val = val.to(device=weight.device, dtype=torch.float32).contiguous()
        B, K = idx.shape
        F, L1 = weight.shape
        out = _empty((B, L1), torch.float32, weight)
        w, b = weight.detach().contiguous(), bias.detach().contiguous()
        with _lib.on_device_of(weight):
            check(_lib.lib().nnue_ft_fwd_indexed(B, K, F, L1, dptr(idx), dptr(val), dptr(w), dptr(b), dptr(out),
                                                 stream_ptr()))
        ctx.save_for_backward(idx, val, w)
        return out

    @staticmethod
    def backward(ctx, g_out):
        idx, val, w = ctx.saved_tensors
        B, K = idx.shape
        F, L1 = w.shape
        g_out = g_out.contiguous().float()
        # the (row, sample, value) triples sorted by row -- the segment reduction's input -- by the library's own stable
        # counting sort on the device (invalid slots sort behind every row: always B * K triples, no host read-back)
        n = B * K
        rows = _empty((n,), torch.int32, w)
        samples = _empty((n,), torch.int32, w)
        vals = _empty((n,), torch.float32, w)
        L = _lib.lib()
        sort_bytes = int(L.nnue_ft_sort_pairs_workspace_bytes(B, K, F))
        sort_ws = _empty((sort_bytes,), torch.uint8, w)
        with _lib.on_device_of(w):
            check(L.nnue_ft_sort_pairs(B, K, F, dptr(idx), dptr(val), dptr(rows), dptr(samples), dptr(vals), dptr(sort_ws),
                                       sort_bytes, stream_ptr()))
        g_w = _empty((F, L1), torch.float32, w)
        g_b = _empty((L1,), torch.float32, w)
        g_val = _empty((B, K), torch.float32, w) if ctx.needs_input_grad[1] else None
        with _lib.on_device_of(w):
            check(_lib.lib().nnue_ft_bwd_indexed(B, K, F, L1, dptr(idx), dptr(w), dptr(g_out), n, dptr(rows) if n else None,
                                                 dptr(samples) if n else None, dptr(vals) if n else None, dptr(g_w),
                                                 dptr(g_b), dptr(g_val), stream_ptr()))
        return None, g_val, g_w, g_b


class FeatureTransformer(nn.Module):
    """Feature transformer (input layer), nnue.py:674-710."""

    def __init__(self, num_features: int, output_size: int):
        super().__init__()
        self.num_features = num_features
        self.output_size = output_size
        self.weight = nn.Parameter(torch.randn(num_features, output_size) * 0.1)
        self.bias = nn.Parameter(torch.zeros(output_size))

    def forward(self, feature_indices: torch.Tensor, feature_values: torch.Tensor):
        return _FTIndexed.apply(feature_indices, feature_values, self.weight, self.bias)


class SimpleClassifier(nn.Module):
    """nnue.py:713-738.  The `nn.Linear` modules hold the parameters (serialize.py picks them by
    isinstance); `NNUE.forward` runs them through the fused head kernels instead."""

    def __init__(self, l1_size: int, l2_size: int, l3_size: int, num_classes: int):
        super().__init__()
        self.num_classes = num_classes
        self.classifier = nn.Sequential(
            nn.Linear(l1_size, l2_size),
            nn.ReLU(),
            nn.Linear(l2_size, l3_size),
            nn.ReLU(),
            nn.Linear(l3_size, num_classes),
        )

    def forward(self, x: torch.Tensor):
        return self.classifier(x)


class _Saved:
    """Forward products the backward needs (kept on the autograd ctx)."""
    __slots__ = ("shape", "images", "bits_s", "bits_t", "xpad", "ft_out", "act1", "act2", "params")


def _mark(marks, name):
    """Stage boundary for bench.py's per-kernel timing: records a CUDA event on the current stream."""
    if marks is not None:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        marks.append((name, ev))


def _run_forward(shape, images, params, need_backward, marks=None):
    """extract -> feature transformer -> head.  `params` = (thr, conv_w, ft_w, ft_b, w1, b1, w2, b2, w3, b3)."""
    L = _lib.lib()
    st = stream_ptr()
    thr, conv_w, ft_w, ft_b, w1, b1, w2, b2, w3, b3 = params
    sp = ctypes.byref(shape)
    bits_s = _empty((shape.B, shape.NW), torch.int32, images)
    want_t = need_backward and L.nnue_wants_transposed_bits(sp)
    bits_t = _empty((shape.PP, shape.BW), torch.int32, images) if want_t else None
    xpad = None  # the backward recomputes the conv activations instead of storing them
    _mark(marks, "start")
    check(L.nnue_extract_fwd(sp, dptr(images), dptr(conv_w), dptr(thr), dptr(bits_s), dptr(bits_t), dptr(xpad), None,
                             None, st))
    _mark(marks, "extract_fwd")
    ft_out = _empty((shape.B, shape.L1), torch.float32, images)
    ws_bytes = _lib.workspace_bytes(shape)
    ws = _empty((ws_bytes,), torch.uint8, images)
    check(L.nnue_ft_fwd(sp, dptr(bits_s), dptr(ft_w), dptr(ft_b), dptr(ft_out), dptr(ws), ws_bytes, st))
    _mark(marks, "ft_fwd")
    act1 = _empty((shape.B, shape.L2), torch.float32, images)
    act2 = _empty((shape.B, shape.L3), torch.float32, images)
    logits = _empty((shape.B, shape.NC), torch.float32, images)
    check(L.nnue_head_fwd(sp, dptr(ft_out), dptr(w1), dptr(b1), dptr(w2), dptr(b2), dptr(w3), dptr(b3), dptr(act1),
                          dptr(act2), dptr(logits), st))
    _mark(marks, "head_fwd")
    return logits, bits_s, bits_t, xpad, ft_out, act1, act2


def _run_backward(shape, images, params, bits_s, bits_t, xpad, ft_out, act1, act2, g_logits, grads=None, marks=None):
    """All parameter gradients from g_logits.  `grads` (optional) are preallocated output tensors in
    parameter order (thr, conv_w, ft_w, ft_b, w1, b1, w2, b2, w3, b3) -- e.g. views of a flat
    data-parallel gradient buffer -- otherwise fresh tensors are returned."""
    L = _lib.lib()
    st = stream_ptr()
    thr, conv_w, ft_w, ft_b, w1, b1, w2, b2, w3, b3 = params
    sp = ctypes.byref(shape)
    if grads is None:
        grads = tuple(torch.empty_like(p) for p in params)
    g_thr, g_conv_w, g_ft_w, g_ft_b, g_w1, g_b1, g_w2, g_b2, g_w3, g_b3 = grads
    ws_bytes = _lib.workspace_bytes(shape)
    ws = _empty((ws_bytes,), torch.uint8, images)
    g_ft = _empty((shape.B, shape.L1), torch.float32, images)
    check(L.nnue_head_bwd(sp, dptr(g_logits), dptr(ft_out), dptr(act1), dptr(act2), dptr(w1), dptr(w2), dptr(w3),
                          dptr(g_w1), dptr(g_b1), dptr(g_w2), dptr(g_b2), dptr(g_w3), dptr(g_b3), dptr(g_ft),
                          dptr(ws), ws_bytes, st))
    _mark(marks, "head_bwd")
    check(L.nnue_ft_bwd_dw(sp, dptr(bits_s), dptr(bits_t), dptr(g_ft), dptr(g_ft_w), dptr(g_ft_b), dptr(ws), ws_bytes,
                           st))
    _mark(marks, "ft_bwd_dw")
    check(L.nnue_input_bwd(sp, dptr(images), dptr(bits_s), dptr(ft_w), dptr(g_ft), dptr(conv_w), dptr(thr),
                           dptr(g_conv_w), dptr(g_thr), dptr(ws), ws_bytes, st))
    _mark(marks, "input_bwd")
    return grads


class _NNUEForward(torch.autograd.Function):
    """images -> logits as one autograd node (nnue.py:637-671)."""

    @staticmethod
    def forward(ctx, images, stride, C, G, thr, conv_w, ft_w, ft_b, w1, b1, w2, b2, w3, b3):
        params = tuple(p.detach().contiguous() for p in (thr, conv_w, ft_w, ft_b, w1, b1, w2, b2, w3, b3))
        B, _, H, W = images.shape
        shape = _lib.make_shape(B, H, W, C, G, ft_w.shape[1], w1.shape[0], w2.shape[0], w3.shape[0], stride)
        need_bwd = any(ctx.needs_input_grad[4:])
        with _lib.on_device_of(images):
            logits, bits_s, bits_t, xpad, ft_out, act1, act2 = _run_forward(shape, images, params, need_bwd)
        if need_bwd:
            sv = _Saved()
            sv.shape, sv.images, sv.bits_s, sv.bits_t, sv.xpad = shape, images, bits_s, bits_t, xpad
            sv.ft_out, sv.act1, sv.act2, sv.params = ft_out, act1, act2, params
            ctx.sv = sv
        return logits

    @staticmethod
    def backward(ctx, g_logits):
        sv = ctx.sv
        with _lib.on_device_of(sv.images):
            g = _run_backward(sv.shape, sv.images, sv.params, sv.bits_s, sv.bits_t, sv.xpad, sv.ft_out, sv.act1, sv.act2,
                              g_logits.contiguous().float())
        return (None, None, None, None) + tuple(g)


# Training keeps the pre-threshold conv activations ([B, PP] fp32) for the threshold gradient instead of
# recomputing them in the backward (measured faster on B200; set False to trade the time for the memory).
STORE_ACTIVATIONS = os.environ.get("NNUE_STORE_ACTIVATIONS", "1") != "0"


OVERLAP_TABLE_GRADIENT = True
# wide stacks: the layer-1 weight-gradient chain of the head on the side stream (nnue_head_train_overlapped)
# -- measured at L1 = 1024, batch 16384 on one B200: 0.814 ms per step with, 0.795 ms without (the chain then competes with
# the value-gradient GEMM for the tensor pipe and delays the table gradient behind it on the same side stream): off by default
HEAD_SIDE_STREAM = os.environ.get("NNUE_HEAD_SIDE_STREAM", "0") == "1"
# Format the table tiles on the side stream while the images are extracted (tcgen05 shapes).  Off by default: at
# config D the two 4 us formatting kernels running beside the extraction cost more than they save (measured on one
# box, 200 steps each: 227 us per step with, 216 us without -- the extraction is issue-bound and shares its SMs).
PREFORMAT_TABLES = os.environ.get("NNUE_PREFORMAT_TABLES", "0") == "1"


# Under data parallelism the side stream runs at high priority: the table gradient and the push of the exchange behind it
# get their SMs first (2 GPUs, config D: 218.3 vs 220.3 us per step).  Without an exchange it stays at normal priority: at
# L1 = 1024 the 200 us table gradient in front of everything else costs the step 0.817 vs 0.795 ms.
SIDE_STREAM_PRIORITY = int(os.environ.get("NNUE_SIDE_PRIORITY", "-1"))
_SIDE_STREAMS = {}


def _side_stream(device, high=False):
    key = (device.type, device.index, bool(high))
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device, priority=SIDE_STREAM_PRIORITY if high else 0)
    return _SIDE_STREAMS[key]


def _run_train_step(shape, images, labels, params, inv_count, grads=None, loss_out=None, marks=None, exchange=None):
    """One whole training step of the hot path -- forward, mean cross-entropy, every parameter gradient --
    in seven launches: extract, feature transformer, fused head step (forward + loss + backward), the two
    feature-transformer gradients, conv / threshold gradients.  `grads` / `loss_out` may be preallocated
    (views of a flat data-parallel buffer).  Returns (loss [1], grads in parameter order).

    `exchange` (data parallelism, train.DataParallelStep): an object with early() / late() / full() that all-reduces
    the slices of the flat gradient buffer.  early() -- feature transformer, head, loss -- is issued on the side stream
    right behind the table gradient, so it travels while the value / conv gradients are still being computed; late()
    -- conv weights and thresholds, a few hundred floats -- at the end."""
    L = _lib.lib()
    st = stream_ptr()
    thr, conv_w, ft_w, ft_b, w1, b1, w2, b2, w3, b3 = params
    sp = ctypes.byref(shape)
    if grads is None:
        grads = tuple(torch.empty_like(p) for p in params)
    if loss_out is None:
        loss_out = _empty((1,), torch.float32, images)
    g_thr, g_conv_w, g_ft_w, g_ft_b, g_w1, g_b1, g_w2, g_b2, g_w3, g_b3 = grads
    bits_s = _empty((shape.B, shape.NW), torch.int32, images)
    bits_t = _empty((shape.PP, shape.BW), torch.int32, images) if L.nnue_wants_transposed_bits(sp) else None
    # the dense conv-gradient kernel takes the pre-threshold activations from the forward when asked to
    xpad = _empty((shape.B, shape.PP), torch.float32, images) if (
        (STORE_ACTIVATIONS and L.nnue_input_bwd_is_dense(sp)) or L.nnue_input_bwd_wants_activations(sp)) else None
    # tcgen05 shapes: the table's split-bf16 tiles depend only on the weights -- outside the per-stage timing mode they
    # are formatted on the side stream while the images are being extracted
    tables = None
    tb = int(L.nnue_ft_tables_bytes(sp)) if (marks is None and PREFORMAT_TABLES and OVERLAP_TABLE_GRADIENT and L.nnue_input_bwd_is_dense(sp)) else 0
    if tb:
        tables = _empty((tb,), torch.uint8, images)
        side0, main0 = _side_stream(images.device), torch.cuda.current_stream()
        side0.wait_stream(main0)
        with torch.cuda.stream(side0):
            check(L.nnue_ft_format_tables(sp, dptr(ft_w), dptr(tables), ctypes.c_void_p(side0.cuda_stream)))
    _mark(marks, "start")
    check(L.nnue_extract_fwd(sp, dptr(images), dptr(conv_w), dptr(thr), dptr(bits_s), dptr(bits_t), dptr(xpad), None,
                             None, st))
    _mark(marks, "extract_fwd")
    ft_out = _empty((shape.B, shape.L1), torch.float32, images)
    ws_bytes = _lib.workspace_bytes(shape)
    ws = _empty((ws_bytes,), torch.uint8, images)
    if tables is not None:
        torch.cuda.current_stream().wait_stream(side0)
        check(L.nnue_ft_fwd_tables(sp, dptr(bits_s), dptr(tables), dptr(ft_b), dptr(ft_out), st))
    else:
        check(L.nnue_ft_fwd(sp, dptr(bits_s), dptr(ft_w), dptr(ft_b), dptr(ft_out), dptr(ws), ws_bytes, st))
    _mark(marks, "ft_fwd")
    g_ft = _empty((shape.B, shape.L1), torch.float32, images)
    # wide stacks on the tensor cores: the layer-1 weight-gradient chain (independent of g_ft and of everything after it)
    # goes to the side stream with its own scratch; the side stream is joined at the end of the step
    head_side = (marks is None and OVERLAP_TABLE_GRADIENT and HEAD_SIDE_STREAM and L.nnue_ft_uses_mma(sp)
                 and L.nnue_input_bwd_is_dense(sp))
    side_bytes = int(L.nnue_head_side_workspace_bytes(sp)) if head_side else 0
    if side_bytes:
        side_ws = _empty((side_bytes,), torch.uint8, images)
        check(L.nnue_head_train_overlapped(sp, dptr(ft_out), dptr(labels), inv_count, dptr(w1), dptr(b1), dptr(w2), dptr(b2),
                                           dptr(w3), dptr(b3), dptr(loss_out), dptr(g_ft), dptr(g_w1), dptr(g_b1), dptr(g_w2),
                                           dptr(g_b2), dptr(g_w3), dptr(g_b3), dptr(ws), ws_bytes, st, dptr(side_ws), side_bytes,
                                           ctypes.c_void_p(_side_stream(images.device, high=exchange is not None).cuda_stream)))
    else:
        check(L.nnue_head_train(sp, dptr(ft_out), dptr(labels), inv_count, dptr(w1), dptr(b1), dptr(w2), dptr(b2),
                                dptr(w3), dptr(b3), dptr(loss_out), dptr(g_ft), dptr(g_w1), dptr(g_b1), dptr(g_w2),
                                dptr(g_b2), dptr(g_w3), dptr(g_b3), dptr(ws), ws_bytes, st))
    _mark(marks, "head_train")
    if L.nnue_ft_uses_mma(sp) and L.nnue_input_bwd_is_dense(sp):  # small tables: tensor-core contractions
        # The table gradient depends only on g_ft and nothing downstream depends on it, while the value gradient
        # and the conv gradient that follow are latency-bound and leave issue slots free: outside the per-stage
        # timing mode it runs on a side stream next to them (it uses its own part of the workspace).
        side = _side_stream(images.device, high=exchange is not None) if (marks is None and OVERLAP_TABLE_GRADIENT) else None
        if side is not None:
            main = torch.cuda.current_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                check(L.nnue_ft_bwd_dw(sp, dptr(bits_s), None, dptr(g_ft), dptr(g_ft_w), dptr(g_ft_b), dptr(ws), ws_bytes,
                                       ctypes.c_void_p(side.cuda_stream)))
                if exchange is not None:
                    exchange.early(ctypes.c_void_p(side.cuda_stream))
        else:
            check(L.nnue_ft_bwd_dw(sp, dptr(bits_s), None, dptr(g_ft), dptr(g_ft_w), dptr(g_ft_b), dptr(ws), ws_bytes, st))
            _mark(marks, "ft_bwd_dw")
            if exchange is not None:
                exchange.early(st)
        if xpad is not None and tables is None and L.nnue_input_bwd_fused_ok(sp):
            # one kernel: the value gradient stays in tensor memory, the conv-gradient consumers read it from there
            check(L.nnue_input_bwd_fused(sp, dptr(images), dptr(bits_s), dptr(xpad), dptr(ft_w), dptr(g_ft), dptr(thr),
                                         dptr(g_conv_w), dptr(g_thr), dptr(ws), ws_bytes, st))
            _mark(marks, "input_bwd_onchip")
        else:
            gbin = _empty((shape.B, shape.PP), torch.float32, images)
            if tables is not None:
                check(L.nnue_ft_bwd_gbin_tables(sp, dptr(bits_s), dptr(tables), dptr(g_ft), dptr(gbin), dptr(ws), ws_bytes, st))
            else:
                check(L.nnue_ft_bwd_gbin(sp, dptr(bits_s), dptr(ft_w), dptr(g_ft), dptr(gbin), dptr(ws), ws_bytes, st))
            _mark(marks, "ft_bwd_gbin")
            check(L.nnue_conv_bwd(sp, dptr(images), dptr(gbin), dptr(xpad), dptr(conv_w), dptr(thr), dptr(g_conv_w), dptr(g_thr),
                                  dptr(ws), ws_bytes, st))
            _mark(marks, "conv_bwd")
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)  # the step's gradients are complete on the caller's stream
        if exchange is not None:
            exchange.late(st)
            _mark(marks, "exchange")
        return loss_out, grads
    if L.nnue_ft_bwd_is_fused(sp):  # both feature-transformer gradients in one CUDA-core pass over g_ft
        gbin = _empty((shape.B, shape.PP), torch.float32, images)
        check(L.nnue_ft_bwd(sp, dptr(bits_s), dptr(ft_w), dptr(g_ft), dptr(g_ft_w), dptr(g_ft_b), dptr(gbin), dptr(ws),
                            ws_bytes, st))
        _mark(marks, "ft_bwd")
        check(L.nnue_conv_bwd(sp, dptr(images), dptr(gbin), dptr(xpad), dptr(conv_w), dptr(thr), dptr(g_conv_w), dptr(g_thr),
                              dptr(ws), ws_bytes, st))
        _mark(marks, "conv_bwd")
        if exchange is not None:
            exchange.full(st)
            _mark(marks, "exchange")
        return loss_out, grads
    # General shapes (ImageNet-sized images, large tables).  Under data parallelism the table gradient runs on the side
    # stream with its own scratch and its slice of the exchange (the 268 MB of SURVEY config I) follows it there, while
    # the value / conv gradients run on the caller's stream.
    side = _side_stream(images.device, high=True) if (exchange is not None and marks is None and OVERLAP_TABLE_GRADIENT) else None
    if side is not None:
        main = torch.cuda.current_stream()
        side.wait_stream(main)
        ws_dw = _empty((ws_bytes,), torch.uint8, images)  # (the general input gradient uses all of `ws`)
        with torch.cuda.stream(side):
            sst = ctypes.c_void_p(side.cuda_stream)
            check(L.nnue_ft_bwd_dw(sp, dptr(bits_s), dptr(bits_t), dptr(g_ft), dptr(g_ft_w), dptr(g_ft_b), dptr(ws_dw), ws_bytes, sst))
            exchange.early(sst)
    else:
        check(L.nnue_ft_bwd_dw(sp, dptr(bits_s), dptr(bits_t), dptr(g_ft), dptr(g_ft_w), dptr(g_ft_b), dptr(ws), ws_bytes,
                               st))
        _mark(marks, "ft_bwd_dw")
        if exchange is not None:
            exchange.early(st)
    if L.nnue_input_bwd_is_dense(sp):
        gbin = _empty((shape.B, shape.PP), torch.float32, images)
        check(L.nnue_ft_bwd_gbin(sp, dptr(bits_s), dptr(ft_w), dptr(g_ft), dptr(gbin), None, 0, st))
        _mark(marks, "ft_bwd_gbin")
        check(L.nnue_conv_bwd(sp, dptr(images), dptr(gbin), dptr(xpad), dptr(conv_w), dptr(thr), dptr(g_conv_w), dptr(g_thr),
                              dptr(ws), ws_bytes, st))
        _mark(marks, "conv_bwd")
    else:
        check(L.nnue_input_bwd_stored(sp, dptr(images), dptr(bits_s), dptr(xpad), dptr(ft_w), dptr(g_ft), dptr(conv_w),
                                      dptr(thr), dptr(g_conv_w), dptr(g_thr), dptr(ws), ws_bytes, st))
        _mark(marks, "input_bwd")
    if side is not None:
        torch.cuda.current_stream().wait_stream(side)
    if exchange is not None:
        exchange.late(st)
        _mark(marks, "exchange")
    return loss_out, grads


# The module-level loss (`model.loss`, `train.compute_loss`) replays the step as a CUDA graph once it has seen the same
# (input buffers, parameter storages, shape) twice -- the same cache DataParallelStep uses; set False to launch eagerly.
CUDA_GRAPHS = os.environ.get("NNUE_CUDA_GRAPHS", "1") != "0"


_GRAPH_RUNNERS = weakref.WeakKeyDictionary()  # model -> its private DataParallelStep (graph cache + static buffers)


def _flat_views(flat, shapes):
    views, off = [], 0
    for shp in shapes:
        n = 1
        for d in shp:
            n *= d
        views.append(flat[off:off + n].view(shp))
        off += n
    return views, off


class _NNUELoss(torch.autograd.Function):
    """images, labels -> mean cross-entropy (train.py:250-254).  The loss is a scalar, so every parameter
    gradient is linear in the upstream gradient: when gradients are wanted the whole step (forward, loss,
    backward) runs inside `forward` through the fused kernels -- every gradient lands in ONE flat buffer owned by this
    autograd node -- and `backward` is one kernel that scales that buffer by the incoming scalar into a fresh one."""

    @staticmethod
    def forward(ctx, images, labels, stride, C, G, inv_count, runner, thr, conv_w, ft_w, ft_b, w1, b1, w2, b2, w3, b3):
        params = tuple(p.detach().contiguous() for p in (thr, conv_w, ft_w, ft_b, w1, b1, w2, b2, w3, b3))
        B, _, H, W = images.shape
        shape = _lib.make_shape(B, H, W, C, G, ft_w.shape[1], w1.shape[0], w2.shape[0], w3.shape[0], stride)
        with _lib.on_device_of(images):
            if any(ctx.needs_input_grad[7:]):
                ctx.shapes = [tuple(p.shape) for p in params]
                if runner is not None:
                    # graph replay into the runner's static buffer, then ONE copy: this node owns its gradients even if
                    # the model is stepped again before backward() runs
                    flat = runner.run_local_flat(images, labels, inv_count).clone()
                else:
                    n = sum(p.numel() for p in params)
                    flat = _empty(((n + 1 + 3) // 4 * 4,), torch.float32, images)
                    views, n = _flat_views(flat, ctx.shapes)
                    _run_train_step(shape, images, labels, params, inv_count, grads=views, loss_out=flat[n:n + 1])
                ctx.flat = flat
                ctx.n = sum(p.numel() for p in params)
                return flat[ctx.n].clone()
            logits = _run_forward(shape, images, params, False)[0]
            loss = _empty((1,), torch.float32, images)
            per = _empty((B,), torch.float32, images)
            check(_lib.lib().nnue_ce_fwd_bwd(B, shape.NC, dptr(logits), dptr(labels), inv_count, None, dptr(loss),
                                             dptr(per), None, None, 0, stream_ptr()))
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g_loss):
        flat = ctx.flat
        g = g_loss.detach().to(device=flat.device, dtype=torch.float32).contiguous()
        out = torch.empty_like(flat)
        with _lib.on_device_of(flat):
            check(_lib.lib().nnue_scale_flat(ctx.n, dptr(flat), dptr(g), dptr(out), stream_ptr()))
        return (None,) * 7 + tuple(_flat_views(out, ctx.shapes)[0])


class NNUE(nn.Module):
    """NNUE model for computer vision (nnue.py:447-671): conv 3x3 -> learnable hard threshold ->
    sparse grid features -> feature transformer -> pairwise product -> 3-layer classifier."""

    def __init__(
        self,
        feature_set: Optional[GridFeatureSet] = None,
        l1_size: int = DEFAULT_L1,
        l2_size: int = DEFAULT_L2,
        l3_size: int = DEFAULT_L3,
        loss_params=LossParams(),
        num_classes=1,
        weight_decay=5e-4,
        input_size=32,
    ):
        super().__init__()
        if feature_set is None:
            feature_set = GridFeatureSet(grid_size=10, num_features_per_square=8)
        self.feature_set = feature_set
        self.l1_size = l1_size
        self.l2_size = l2_size
        self.l3_size = l3_size
        self.num_classes = num_classes
        self.loss_params = loss_params
        self.weight_decay = weight_decay
        self.input_size = input_size

        conv_out_channels, conv_stride = self._calculate_conv_params(
            input_size, feature_set.grid_size, feature_set.num_features_per_square)
        # registration order conv -> input -> classifier -> nnue2score -> visual_threshold keeps both the
        # RNG stream of the initialisers and named_parameters() identical to the reference (nnue.py:485-508)
        self.conv = nn.Conv2d(3, conv_out_channels, kernel_size=3, stride=conv_stride, padding=1, bias=False)
        self.input = FeatureTransformer(feature_set.num_features, l1_size)
        self.classifier = SimpleClassifier(l1_size, l2_size, l3_size, num_classes)
        self.nnue2score = nn.Parameter(torch.tensor(600.0))
        self.visual_threshold = nn.Parameter(torch.full((conv_out_channels,), 0.1))

    def _calculate_conv_params(self, input_size, target_grid_size, num_features_per_square):
        """stride = max(1, (input_size-1) // (grid-1)), nnue.py:510-526."""
        conv_stride = max(1, (input_size - 1) // (target_grid_size - 1))
        return num_features_per_square, conv_stride

    # ---- quantisation / serialisation surface (nnue.py:528-588) ---------------------------------
    def _clip_weights(self):
        with torch.no_grad():
            self.input.weight.clamp_(-1.0, 1.0)
            for module in self.classifier.modules():
                if isinstance(module, nn.Linear):
                    module.weight.clamp_(-1.0, 1.0)

    def get_quantized_model_data(self):
        from .serialize import quantize_conv_layer, quantize_linear_layer
        self.eval()
        self._clip_weights()
        linear_layers = [m for m in self.classifier.classifier if isinstance(m, nn.Linear)]
        return {
            "metadata": {
                "feature_set": self.feature_set,
                "L1": self.l1_size,
                "L2": self.l2_size,
                "L3": self.l3_size,
                "num_classes": self.num_classes,
                "nnue2score": self.nnue2score.item(),
                "quantized_one": 127.0,
                "visual_threshold": float(self.visual_threshold.detach().mean().cpu().item()),
            },
            "conv_layer": quantize_conv_layer(self.conv),
            "feature_transformer": quantize_linear_layer(self.input),
            "classifier": {"layers": [quantize_linear_layer(m) for m in linear_layers]},
        }

    # ---- hot path --------------------------------------------------------------------------------
    def _hot_params(self):
        lin = self.classifier.classifier
        return (self.visual_threshold, self.conv.weight, self.input.weight, self.input.bias,
                lin[0].weight, lin[0].bias, lin[2].weight, lin[2].bias, lin[4].weight, lin[4].bias)

    def _check_images(self, images):
        if not images.is_cuda or not self.input.weight.is_cuda:
            raise _lib.NnueError("NNUE runs on CUDA only (no CPU fallback): move the model and the images to a B200")
        if images.dim() != 4 or images.shape[1] != 3:
            raise ValueError(f"expected images [B,3,H,W], got {tuple(images.shape)}")
        return images.detach().float().contiguous()

    def extract_bits(self, images: torch.Tensor):
        """conv + threshold -> (shape, sample-major bitmask int32 [B, NW]).  Internal layout: bit k of
        word c*CW + j is position (channel c, cell 32j + k) of the conv raster."""
        images = self._check_images(images)
        B, _, H, W = images.shape
        fs = self.feature_set
        shape = _lib.make_shape(B, H, W, fs.num_features_per_square, fs.grid_size, self.l1_size, self.l2_size,
                                self.l3_size, self.num_classes, self.conv.stride[0])
        bits = _empty((B, shape.NW), torch.int32, images)
        with _lib.on_device_of(images):
            check(_lib.lib().nnue_extract_fwd(ctypes.byref(shape), dptr(images), dptr(self.conv.weight.detach().contiguous()),
                                              dptr(self.visual_threshold.detach().contiguous()), dptr(bits), None, None,
                                              None, None, stream_ptr()))
        return shape, bits

    def _to_sparse_features(self, binary_features: torch.Tensor):
        """[B,C,Gh,Gw] 0/1 map -> (indices int64 [B,K], values fp32 [B,K]), -1 / 0 padded, K = max(1, max nnz),
        ascending CHW order (nnue.py:590-635).  Values are gathered from the input so the autograd edge to
        `binary_features` is preserved, as in the reference."""
        B = binary_features.shape[0]
        flat = binary_features.reshape(B, -1)
        active = flat > 0.5
        counts = active.sum(dim=1)
        K = max(int(counts.max().item()) if B > 0 else 1, 1)
        order = torch.sort((~active).to(torch.uint8), dim=1, stable=True).indices[:, :K]  # active first, ascending
        keep = torch.arange(K, device=flat.device).unsqueeze(0) < counts.unsqueeze(1)
        idx = torch.where(keep, order, torch.full_like(order, -1))
        val = torch.where(keep, flat.gather(1, order), torch.zeros((), dtype=flat.dtype, device=flat.device))
        return idx, val.float()

    def forward(self, images: torch.Tensor):
        images = self._check_images(images)
        fs = self.feature_set
        return _NNUEForward.apply(images, self.conv.stride[0], fs.num_features_per_square, fs.grid_size,
                                  *self._hot_params())

    def loss(self, images: torch.Tensor, targets: torch.Tensor, global_batch: Optional[int] = None):
        """Mean cross-entropy of `forward(images)` against `targets` with the loss fused into the head
        (what train.compute_loss computes, train.py:250-254).  `global_batch` = total samples across
        data-parallel ranks (defaults to this rank's batch)."""
        images = self._check_images(images)
        fs = self.feature_set
        labels = targets.to(device=images.device, dtype=torch.long).contiguous()
        inv = 1.0 / float(global_batch if global_batch else images.shape[0])
        runner = None
        if CUDA_GRAPHS and torch.is_grad_enabled():
            runner = _GRAPH_RUNNERS.get(self)  # (kept outside the module: attributes, state_dict, deepcopy unchanged)
            if runner is None or runner.device != images.device:
                from .train import DataParallelStep
                runner = DataParallelStep(self, device=images.device, cuda_graphs=True, attach=False, single=True)
                _GRAPH_RUNNERS[self] = runner
        return _NNUELoss.apply(images, labels, self.conv.stride[0], fs.num_features_per_square, fs.grid_size, inv,
                               runner, *self._hot_params())
