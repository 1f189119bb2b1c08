"""oracle/float_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (torch CPU tensors, fp32 or fp64) of the reference's FLOAT
training path: `NNUE.forward` (/root/reference/nnue.py:637-671), the
straight-through threshold (nnue.py:15-59), `_to_sparse_features`
(nnue.py:590-635), `FeatureTransformer.forward` (nnue.py:686-710), the pairwise
product (nnue.py:660-666), `SimpleClassifier` (nnue.py:713-738) and the mean
cross-entropy of train.py:250-254, with the BACKWARD written out in closed form
 instead of taken from autograd.

Two entry points:
  * `step(...)`            closed-form forward+backward, vectorised; the checker.
  * `reference_style_step` the reference's own algorithmic shape (per-sample
                           nonzero / gather-sum loops under autograd); used as the
                           `cpu_baseline` "port" timing leg and as a second opinion.

The float arithmetic of the reference lives in a third-party dependency, PyTorch
ATen (conv2d, index, sum, addmm, cross_entropy; the reference pins no version,
this image has torch 2.11.0+cu128).  Parity status: PINNED against golden
vectors produced by the reference itself in the build container
(tests/golden/make_golden.py -> tests/test_oracle_float.py), at
rtol 1e-5 / atol 1e-5*max|ref| per tensor.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this module.
"""
import torch
import torch.nn.functional as F

STE_SHARPNESS = 10.0  # nnue.py:41


def python_stride(input_size, grid_size):
    """Conv stride rule of the float model, nnue.py:519 (floor; differs from the engine's ceil)."""
    return max(1, (input_size - 1) // (grid_size - 1))


def as_tensors(state, dtype=torch.float32):
    return {k: torch.as_tensor(v).to(dtype).clone() for k, v in state.items()}


def feature_rows(P, num_features):
    """Row of the FT table hit by flat CHW position p: min(p, F-1) (the clamp at nnue.py:701)."""
    return torch.clamp(torch.arange(P), max=num_features - 1)


def extract(state, images, stride):
    """conv 3x3 pad 1 (nnue.py:640) + hard threshold (nnue.py:24) -> (x [B,C,Gh,Gw], bits bool)."""
    x = F.conv2d(images, state["conv.weight"], None, stride=stride, padding=1)
    bits = x > state["visual_threshold"].view(1, -1, 1, 1)
    return x, bits


def sparse_features(bits):
    """_to_sparse_features, nnue.py:590-635: ascending CHW indices, -1 / 0 padding, K >= 1."""
    B = bits.shape[0]
    flat = bits.reshape(B, -1)
    counts = flat.sum(1)
    K = max(1, int(counts.max().item()) if B else 1)
    idx = torch.full((B, K), -1, dtype=torch.long)
    val = torch.zeros((B, K), dtype=torch.float32)
    for b in range(B):
        nz = torch.nonzero(flat[b]).squeeze(-1)
        idx[b, : nz.numel()] = nz
        val[b, : nz.numel()] = 1.0
    return idx, val


def ft_forward(idx, val, weight, bias):
    """FeatureTransformer.forward for arbitrary (idx,val), nnue.py:686-710."""
    Fn = weight.shape[0]
    valid = (idx >= 0).to(weight.dtype)
    rows = torch.clamp(idx, 0, Fn - 1)
    return bias.unsqueeze(0) + (weight[rows] * (val.to(weight.dtype) * valid).unsqueeze(-1)).sum(1)


def ft_backward(idx, val, weight, grad_out):
    """Closed-form gradients of ft_forward w.r.t. weight, bias, val."""
    Fn, L1 = weight.shape
    valid = (idx >= 0).to(weight.dtype)
    rows = torch.clamp(idx, 0, Fn - 1)
    v = val.to(weight.dtype) * valid
    gw = torch.zeros_like(weight)
    gw.index_add_(0, rows.reshape(-1), (v.unsqueeze(-1) * grad_out.unsqueeze(1)).reshape(-1, L1))
    gb = grad_out.sum(0)
    gval = (weight[rows] * grad_out.unsqueeze(1)).sum(-1) * valid
    return gw, gb, gval


def head_forward(state, ft):
    """pairwise (nnue.py:660-666) + SimpleClassifier (nnue.py:728-734)."""
    h = ft.shape[1] // 2
    a, b = ft[:, :h], ft[:, h:2 * h]
    l0 = torch.cat([a * b, a], dim=1)
    z1 = l0 @ state["classifier.classifier.0.weight"].T + state["classifier.classifier.0.bias"]
    r1 = torch.relu(z1)
    z2 = r1 @ state["classifier.classifier.2.weight"].T + state["classifier.classifier.2.bias"]
    r2 = torch.relu(z2)
    logits = r2 @ state["classifier.classifier.4.weight"].T + state["classifier.classifier.4.bias"]
    return logits, (a, b, l0, z1, r1, z2, r2)


def step(state, images, labels, stride, dtype=torch.float32):
    """Closed-form forward + backward of one training step (mean CE loss).

    Returns a dict: logits, loss, conv_out, bits, ft_out, nnz, and `grads` keyed by
    the reference's parameter names (no entry for nnue2score: it is never in the
    graph, tests/test_model.py:179-182).
    """
    st = as_tensors(state, dtype)
    images = torch.as_tensor(images).to(dtype)
    labels = torch.as_tensor(labels).long()
    B = images.shape[0]
    W, bias = st["input.weight"], st["input.bias"]
    Fn, L1 = W.shape
    x, bits = extract(st, images, stride)
    C, Gh, Gw = x.shape[1:]
    P = C * Gh * Gw
    flat = bits.reshape(B, P).to(dtype)
    rows = feature_rows(P, Fn)
    Wext = W[rows]  # [P, L1]
    ft = bias.unsqueeze(0) + flat @ Wext
    logits, (a, b, l0, z1, r1, z2, r2) = head_forward(st, ft)
    logp = torch.log_softmax(logits, dim=1)
    loss = -logp[torch.arange(B), labels].mean()

    # ---- backward ----
    g_logits = torch.softmax(logits, dim=1)
    g_logits[torch.arange(B), labels] -= 1.0
    g_logits /= B
    W1, W2, W3 = (st[f"classifier.classifier.{i}.weight"] for i in (0, 2, 4))
    grads = {}
    grads["classifier.classifier.4.weight"] = g_logits.T @ r2
    grads["classifier.classifier.4.bias"] = g_logits.sum(0)
    g_z2 = (g_logits @ W3) * (z2 > 0).to(dtype)
    grads["classifier.classifier.2.weight"] = g_z2.T @ r1
    grads["classifier.classifier.2.bias"] = g_z2.sum(0)
    g_z1 = (g_z2 @ W2) * (z1 > 0).to(dtype)
    grads["classifier.classifier.0.weight"] = g_z1.T @ l0
    grads["classifier.classifier.0.bias"] = g_z1.sum(0)
    g_l0 = g_z1 @ W1
    h = L1 // 2
    g_ft = torch.zeros_like(ft)
    g_ft[:, :h] = g_l0[:, :h] * b + g_l0[:, h:2 * h]
    g_ft[:, h:2 * h] = g_l0[:, :h] * a
    grads["input.bias"] = g_ft.sum(0)
    gW = torch.zeros_like(W)
    gW.index_add_(0, rows, flat.T @ g_ft)  # clamp aliasing folds every p >= F onto row F-1
    grads["input.weight"] = gW
    g_bin = ((g_ft @ Wext.T) * flat).reshape(B, C, Gh, Gw)  # d out / d val at ACTIVE positions only
    thr = st["visual_threshold"].view(1, -1, 1, 1)
    sig = torch.sigmoid(STE_SHARPNESS * (x - thr))
    grads["visual_threshold"] = -(g_bin * STE_SHARPNESS * sig * (1 - sig)).sum(dim=(0, 2, 3))
    grads["conv.weight"] = torch.nn.grad.conv2d_weight(images, st["conv.weight"].shape, g_bin,
                                                       stride=stride, padding=1)
    return {
        "logits": logits, "loss": loss, "conv_out": x, "bits": bits, "ft_out": ft,
        "nnz": bits.reshape(B, -1).sum(1), "grads": grads, "g_ft": g_ft, "g_bin": g_bin,
    }


def reference_style_step(state, images, labels, stride):
    """The reference's algorithm with its own cost shape: per-sample `nonzero`, per-sample
    gather * value -> sum, autograd backward (nnue.py:601-606, 694-708; train.py:250-254).
    fp32, CPU.  Returns (loss, grads dict)."""
    st = {k: torch.as_tensor(v).float().clone().requires_grad_(k != "nnue2score") for k, v in state.items()}
    images = torch.as_tensor(images).float()
    labels = torch.as_tensor(labels).long()
    B = images.shape[0]
    Fn = st["input.weight"].shape[0]
    x = F.conv2d(images, st["conv.weight"], None, stride=stride, padding=1)
    thr = st["visual_threshold"].view(1, -1, 1, 1)
    hard = (x > thr).float()
    # straight-through value path: forward value `hard`, backward identity to x and the
    # k*sig*(1-sig) surrogate to thr (nnue.py:28-54)
    sig = torch.sigmoid(STE_SHARPNESS * (x - thr)).detach()
    surrogate = (STE_SHARPNESS * sig * (1 - sig)) * (thr - thr.detach())
    binv = hard + (x - x.detach()) - surrogate
    flat = binv.reshape(B, -1)
    outs = []
    for b in range(B):
        nz = torch.nonzero(hard.reshape(B, -1)[b] > 0.5).squeeze(-1)
        if nz.numel():
            rows = torch.clamp(nz, 0, Fn - 1)
            outs.append(st["input.bias"] + (st["input.weight"][rows] * flat[b][nz].unsqueeze(-1)).sum(0))
        else:
            outs.append(st["input.bias"] + 0.0)
    ft = torch.stack(outs)
    logits, _ = head_forward(st, ft)
    loss = F.cross_entropy(logits, labels)
    loss.backward()
    grads = {k: v.grad.detach() for k, v in st.items() if v.grad is not None}
    return loss.detach(), logits.detach(), grads
