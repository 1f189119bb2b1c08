"""Check the tcgen05 forward against the warp-level MMA forward on a small and a full batch."""
import ctypes, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from nnue_vision_b200 import _lib, nnue
from nnue_vision_b200._lib import check, dptr, stream_ptr

torch.manual_seed(1)
model = nnue.NNUE(nnue.GridFeatureSet(10, 8), 64, 32, 8, num_classes=10, input_size=32).cuda()
with torch.no_grad():
    model.input.bias.normal_(0, 0.05)
for B in (100, 16384):
    images = torch.randn(B, 3, 32, 32, device="cuda")
    shape, bits = model.extract_bits(images)
    L = _lib.lib()
    ws_bytes = _lib.workspace_bytes(shape)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    outs = {}
    for name, opt in (("mma", 0), ("umma", 1)):
        _lib.set_option("ft_umma", opt)
        ws_bytes = _lib.workspace_bytes(shape)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
        out = torch.full((B, 64), float("nan"), device="cuda")
        w, b = model.input.weight.detach().contiguous(), model.input.bias.detach().contiguous()
        for _ in range(3):
            check(L.nnue_ft_fwd(ctypes.byref(shape), dptr(bits), dptr(w), dptr(b), dptr(out), dptr(ws), ws_bytes, stream_ptr()))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            check(L.nnue_ft_fwd(ctypes.byref(shape), dptr(bits), dptr(w), dptr(b), dptr(out), dptr(ws), ws_bytes, stream_ptr()))
        e1.record(); torch.cuda.synchronize()
        outs[name] = out.clone()
        print(B, name, "us per call", e0.elapsed_time(e1) / 20 * 1e3, "finite", bool(torch.isfinite(out).all()))
    d = (outs["mma"] - outs["umma"]).abs().max().item()
    print(B, "max |mma - umma|", d, "max |out|", outs["mma"].abs().max().item())
_lib.set_option("ft_umma", 0)
