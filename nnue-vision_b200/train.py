"""Training-step surface of the hot path: `compute_loss` (train.py:250-254 of the reference) and
the batch-sharded data-parallel step.

The reference trains on one device (train.py:263) and has no collective; the only coupling
between samples is the mean-reduced loss and the parameter gradients, so the batch is split
across ranks, every rank scales its cross-entropy by 1/(global batch), the kernels write the
local gradients straight into ONE flat fp32 buffer, and a single NCCL all-reduce (sum) over
NVLink makes it the global mean gradient.  `clip_grad_norm_`
(train.py:363-364) must run after `step()` on the reduced buffer.
"""
from typing import Callable, Optional

import torch
import torch.distributed as dist

# gradient-bearing parameters in the reference's registration order (nnue2score is never in the
# graph: tests/test_model.py:179-182)
GRAD_PARAM_NAMES = (
    "visual_threshold", "conv.weight", "input.weight", "input.bias",
    "classifier.classifier.0.weight", "classifier.classifier.0.bias",
    "classifier.classifier.2.weight", "classifier.classifier.2.bias",
    "classifier.classifier.4.weight", "classifier.classifier.4.bias",
)


def compute_loss(model, batch):
    """F.cross_entropy(model(images), targets.long()) with the CE fused into the head kernels."""
    images, targets = batch
    return model.loss(images, targets)


class FlatGradBuffer:
    """One contiguous fp32 buffer holding every gradient (+ one trailing slot for the loss) with
    per-parameter views; `attach()` makes the views the parameters' `.grad`."""

    def __init__(self, named_params, device):
        self.names = [n for n in GRAD_PARAM_NAMES if n in named_params]
        self.params = [named_params[n] for n in self.names]
        sizes = [p.numel() for p in self.params]
        self.flat = torch.zeros(sum(sizes) + 1, dtype=torch.float32, device=device)
        self.views, off = [], 0
        for p, n in zip(self.params, sizes):
            self.views.append(self.flat[off:off + n].view(p.shape))
            off += n
        self.loss = self.flat[off:off + 1]

    def attach(self):
        for p, v in zip(self.params, self.views):
            p.grad = v

    def numel(self):
        return self.flat.numel()


def _cuda_local_step(model):
    """The B200 hot path: forward, fused CE, backward -- every kernel through the C ABI."""
    from . import _lib, nnue as _nnue

    def run(images, labels, inv_count, buf: FlatGradBuffer, marks=None):
        images = model._check_images(images)
        labels = labels.to(device=images.device, dtype=torch.long).contiguous()
        fs = model.feature_set
        B, _, H, W = images.shape
        params = tuple(p.detach().contiguous() for p in model._hot_params())
        shape = _lib.make_shape(B, H, W, fs.num_features_per_square, fs.grid_size, model.l1_size, model.l2_size,
                                model.l3_size, model.num_classes, model.conv.stride[0])
        # gradients land directly in the flat buffer's views, the loss in its trailing slot
        _nnue._run_train_step(shape, images, labels, params, inv_count, grads=buf.views, loss_out=buf.loss,
                              marks=marks)

    return run


class DataParallelStep:
    """fwd + bwd (+ one all-reduce) of the NNUE over a batch sharded across ranks.

    step(images, labels) takes THIS rank's shard and returns the global mean loss as a 0-d device
    tensor; afterwards every parameter's `.grad` (a view of the flat buffer) holds the gradient of
    the global mean loss.  With world_size == 1 no collective is issued.

    `local_step` is injectable so the host logic (flat layout, scaling, all-reduce) can be tested
    on CPU with the gloo backend; the default is the CUDA hot path.
    """

    def __init__(self, model, process_group=None, local_step: Optional[Callable] = None, device=None):
        self.model = model
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        named = dict(model.named_parameters())
        device = device if device is not None else named["input.weight"].device
        self.buf = FlatGradBuffer(named, device)
        self.buf.attach()
        self._local = local_step if local_step is not None else _cuda_local_step(model)

    def step(self, images, labels, global_batch: Optional[int] = None, marks=None):
        if global_batch is None:
            global_batch = images.shape[0] * self.world  # equal shards
        if marks is not None:
            self._local(images, labels, 1.0 / float(global_batch), self.buf, marks)
        else:
            self._local(images, labels, 1.0 / float(global_batch), self.buf)
        if self.world > 1:
            dist.all_reduce(self.buf.flat, op=dist.ReduceOp.SUM, group=self.group)
        return self.buf.loss.reshape(())
