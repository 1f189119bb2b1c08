"""Check the tcgen05 feature-transformer kernels (option ft_umma=1) against the other kernel families
(ft_umma=0) on several shapes, and time the three contractions at the benchmark batch."""
import ctypes, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from nnue_vision_b200 import _lib, nnue
from nnue_vision_b200._lib import check, dptr, stream_ptr

CASES = [  # grid, C, L1, L2, L3, NC, model_input, image, batch
    ("T", 8, 4, 64, 4, 8, 10, 32, 32, 16),
    ("D", 10, 8, 64, 32, 8, 10, 32, 32, 300),
    ("D_full", 10, 8, 64, 32, 8, 10, 32, 32, 16384),
    ("L1_128", 6, 16, 128, 16, 32, 1000, 64, 64, 50),
    ("big_into_small", 4, 8, 64, 4, 4, 10, 32, 96, 33),
    ("D1k", 10, 8, 1024, 128, 32, 10, 32, 32, 2048),
]
only = sys.argv[1:]


def run(model, images, labels, umma):
    _lib.set_option("ft_umma", umma)
    model.zero_grad()
    loss = model.loss(images, labels)
    loss.backward()
    torch.cuda.synchronize()
    return loss.item(), {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}


for name, G, C, L1, L2, L3, NC, msize, isize, B in CASES:
    if only and name not in only:
        continue
    torch.manual_seed(1)
    model = nnue.NNUE(nnue.GridFeatureSet(G, C), L1, L2, L3, num_classes=NC, input_size=msize).cuda()
    with torch.no_grad():
        model.input.bias.normal_(0, 0.05)
    images = torch.randn(B, 3, isize, isize, device="cuda")
    labels = torch.randint(0, NC, (B,), device="cuda")
    l0, g0 = run(model, images, labels, 0)
    l1, g1 = run(model, images, labels, 1)
    worst = 0.0
    for k in g0:
        ref = g0[k].double()
        err = ((g1[k].double() - ref).abs() / (1e-5 * ref.abs() + 1e-5 * ref.abs().max() + 1e-30)).max().item()
        worst = max(worst, err)
        flag = "" if err <= 1.0 else "   <-- FAIL"
        print(f"{name:16s} {k:32s} err/tol {err:9.3f}{flag}")
    print(f"{name:16s} loss {l0:.7f} vs {l1:.7f}   worst err/tol {worst:.3f}", flush=True)

# timing of the three contractions at config D, batch 16384
if not only or "time" in only:
    for L1, B in ((64, 16384), (1024, 16384)):
        torch.manual_seed(1)
        model = nnue.NNUE(nnue.GridFeatureSet(10, 8), L1, 32, 8, num_classes=10, input_size=32).cuda()
        images = torch.randn(B, 3, 32, 32, device="cuda")
        shape, bits = model.extract_bits(images)
        L = _lib.lib()
        w, b = model.input.weight.detach().contiguous(), model.input.bias.detach().contiguous()
        g_ft = torch.randn(B, L1, device="cuda") * 1e-4
        out = torch.empty(B, L1, device="cuda")
        gw, gb = torch.empty_like(w), torch.empty_like(b)
        gbin = torch.empty(B, shape.PP, device="cuda")
        sp = ctypes.byref(shape)
        for umma in (0, 1):
            _lib.set_option("ft_umma", umma)
            if not L.nnue_ft_uses_mma(sp):
                continue
            ws_bytes = _lib.workspace_bytes(shape)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
            calls = {
                "ft_fwd": lambda: check(L.nnue_ft_fwd(sp, dptr(bits), dptr(w), dptr(b), dptr(out), dptr(ws), ws_bytes, stream_ptr())),
                "ft_bwd_dw": lambda: check(L.nnue_ft_bwd_dw(sp, dptr(bits), None, dptr(g_ft), dptr(gw), dptr(gb), dptr(ws), ws_bytes, stream_ptr())),
                "ft_bwd_gbin": lambda: check(L.nnue_ft_bwd_gbin(sp, dptr(bits), dptr(w), dptr(g_ft), dptr(gbin), dptr(ws), ws_bytes, stream_ptr())),
            }
            for cname, fn in calls.items():
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                print(f"L1={L1} B={B} umma={umma} {cname:12s} {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us per call", flush=True)
_lib.set_option("ft_umma", 0)
