#!/usr/bin/env python3
"""Both formulations of the feature transformer on one workload, timed alone with CUDA events.

    python tools/ft_compare.py --workload imagenet_large_b4096 [--batch B] [--reps 5]

  dense  : bitmask x split-bf16 table on the tensor cores (ft_umma.cu), operand formatting included
  gather : index-driven row gather / segment reduction (ft_gather.cu, ft.cu)

Prints one JSON line: per call and formulation the device time, the algorithmic bytes of SURVEY 8d
(rows actually gathered), the achieved GB/s against the measured HBM peak, and the largest relative
difference between the two formulations' results (they must agree to the float-path bar).
"""
import argparse
import ctypes
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2]


def rel_diff(a, b):
    scale = float(b.abs().max()) or 1.0
    return float((a - b).abs().max()) / scale


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="imagenet_large_b4096")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="", help="comma list of calls (fwd,dw,dval) to run")
    ap.add_argument("--forms", default="dense,gather")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE")
    a = ap.parse_args()
    from nnue_vision_b200 import _lib
    from nnue_vision_b200._lib import check, dptr
    L = _lib.lib()
    for kv in a.opt:
        key, val = kv.split("=")
        _lib.set_option(key, int(val))
    w = dict(bench.WORKLOADS[a.workload])
    B = a.batch or w["batch"]
    dev = torch.device("cuda", 0)
    model = bench.build_model(w, dev)
    images, labels = bench.synthetic_batch(w, B, seed=1, device=dev)
    st = _lib.stream_ptr()
    peak, peak_src = bench.peaks()
    thr, conv_w, ft_w, ft_b = (p.detach().contiguous() for p in model._hot_params()[:4])
    g_ft = torch.randn(B, w["L1"], device=dev) * 1e-3
    calls = [c for c in (a.only.split(",") if a.only else ("fwd", "dw", "dval"))]
    out = {"workload": a.workload, "batch": B, "peak_gbs": peak, "peak_source": peak_src, "calls": {}}
    results = {}
    for form in a.forms.split(","):
        _lib.set_option("ft_form", 1 if form == "dense" else 2)
        shape = _lib.make_shape(B, w["image"], w["image"], w["C"], w["grid"], w["L1"], w["L2"], w["L3"], w["NC"],
                                model.conv.stride[0])
        sp = ctypes.byref(shape)
        ws_bytes = _lib.workspace_bytes(shape)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        bits_s = torch.empty(B, shape.NW, dtype=torch.int32, device=dev)
        want_t = bool(L.nnue_wants_transposed_bits(sp))
        bits_t = torch.empty(shape.PP, shape.BW, dtype=torch.int32, device=dev) if want_t else None
        xpad = torch.empty(B, shape.PP, dtype=torch.float32, device=dev)
        nnz = torch.empty(B, dtype=torch.int32, device=dev)
        check(L.nnue_extract_fwd(sp, dptr(images), dptr(conv_w), dptr(thr), dptr(bits_s), dptr(bits_t), dptr(xpad), None,
                                 dptr(nnz), st))
        nnz_total = int(nnz.sum())
        L1, F = w["L1"], shape.F
        ft_out = torch.empty(B, L1, device=dev)
        g_w, g_b = torch.empty(F, L1, device=dev), torch.empty(L1, device=dev)
        dval = torch.zeros(B, shape.PP, device=dev)
        g_thr = torch.empty(w["C"], device=dev)
        algo = {  # rows gathered + results written (+ the indices, here one bit per position)
            "fwd": nnz_total * L1 * 4 + B * L1 * 4 + B * shape.NW * 4,
            "dw": nnz_total * L1 * 4 + F * L1 * 4 + B * shape.NW * 4,
            "dval": nnz_total * L1 * 4 + B * L1 * 4 + nnz_total * 4 + B * shape.NW * 4,
        }
        fns = {
            "fwd": lambda: check(L.nnue_ft_fwd(sp, dptr(bits_s), dptr(ft_w), dptr(ft_b), dptr(ft_out), dptr(ws), ws_bytes, st)),
            "dw": lambda: check(L.nnue_ft_bwd_dw(sp, dptr(bits_s), dptr(bits_t), dptr(g_ft), dptr(g_w), dptr(g_b), dptr(ws),
                                                 ws_bytes, st)),
            "dval": lambda: check(L.nnue_ft_bwd_dval(sp, dptr(bits_s), dptr(ft_w), dptr(g_ft), dptr(xpad), dptr(thr), dptr(dval),
                                                     dptr(g_thr), dptr(ws), ws_bytes, st)),
        }
        for c in calls:
            ms = timed(fns[c], a.reps)
            gbs = algo[c] / (ms * 1e-3) / 1e9
            out["calls"].setdefault(c, {})[form] = {"ms": ms, "algorithmic_bytes": algo[c], "achieved_gbs": gbs,
                                                    "frac_of_hbm_peak": gbs / peak}
        mask = None
        if "dval" in calls:  # the gather form writes only the active positions
            words = bits_s.view(B, shape.NW, 1) >> torch.arange(32, device=dev, dtype=torch.int32).view(1, 1, 32)
            mask = (words & 1).bool().view(B, shape.PP)
        results[form] = {"fwd": ft_out.clone(), "dw": g_w.clone(), "db": g_b.clone(),
                         "dval": torch.where(mask, dval, torch.zeros_like(dval)) if mask is not None else None,
                         "g_thr": g_thr.clone()}
        out["nnz_per_sample"] = nnz_total / B
        out["table_bytes"] = F * L1 * 4
        del ws, bits_t, xpad, dval
    forms = a.forms.split(",")
    if len(forms) == 2:
        x, y = results[forms[0]], results[forms[1]]
        agree = {}
        if "fwd" in calls:
            agree["fwd"] = rel_diff(x["fwd"], y["fwd"])
        if "dw" in calls:
            agree["dw"], agree["db"] = rel_diff(x["dw"], y["dw"]), rel_diff(x["db"], y["db"])
        if "dval" in calls:
            agree["dval"], agree["g_thr"] = rel_diff(x["dval"], y["dval"]), rel_diff(x["g_thr"], y["g_thr"])
        out["max_rel_diff_dense_vs_gather"] = agree
    _lib.set_option("ft_form", 0)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
