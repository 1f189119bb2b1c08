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

    def __init__(self, named_params, device, flat=None):
        self.names = [n for n in GRAD_PARAM_NAMES if n in named_params]
        self.params = [named_params[n] for n in self.names]
        sizes = [p.numel() for p in self.params]
        # `flat` may be supplied (e.g. a symmetric-memory tensor the peers can read); at least sum(sizes) + 1 floats
        total = (sum(sizes) + 1 + 3) // 4 * 4  # padded to whole float4s (the one-shot all-reduce moves 16-byte words)
        self.flat = torch.zeros(total, dtype=torch.float32, device=device) if flat is None else flat
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


class FlatParamBuffer:
    """Makes the gradient-bearing parameters views of ONE contiguous fp32 buffer (same order as
    FlatGradBuffer), so that the optimizer update is a single elementwise pass.  `state_dict()` keys,
    shapes and values are unchanged; only the storage moves."""

    def __init__(self, named_params, device):
        self.names = [n for n in GRAD_PARAM_NAMES if n in named_params]
        self.params = [named_params[n] for n in self.names]
        sizes = [p.numel() for p in self.params]
        self.flat = torch.empty(sum(sizes), dtype=torch.float32, device=device)
        off = 0
        with torch.no_grad():
            for p, n in zip(self.params, sizes):
                view = self.flat[off:off + n].view(p.shape)
                view.copy_(p.data)
                p.data = view
                off += n

    def numel(self):
        return self.flat.numel()


class FusedOptimizer:
    """clip_grad_norm_ + optimizer.step() of the reference's loop (train.py:363-366, 457-471) as two or
    three launches over flat buffers: parameters (FlatParamBuffer), gradients (the DataParallelStep's
    FlatGradBuffer, already all-reduced) and optimizer state.

        step = DataParallelStep(model)
        opt = FusedOptimizer(step, "sgd", lr=0.01, momentum=0.9, weight_decay=2e-4, max_grad_norm=1.0)
        loss = step.step(images, labels); opt.step()

    Semantics are torch's (`torch.optim.SGD` / `torch.optim.Adam` with L2 weight decay,
    `torch.nn.utils.clip_grad_norm_` with norm 2); tests/test_gpu_optim.py checks them against torch.
    The parameter the reference's optimizer never touches (`nnue2score`, no gradient) is left alone."""

    def __init__(self, dp_step, optimizer_type="sgd", lr=0.01, momentum=0.9, weight_decay=0.0, max_grad_norm=0.0,
                 betas=(0.9, 0.999), eps=1e-8):
        from . import _lib
        if optimizer_type not in ("sgd", "adam"):
            raise ValueError(f"optimizer_type must be 'sgd' or 'adam', got {optimizer_type!r}")
        self._lib = _lib
        self.kind, self.lr, self.momentum, self.weight_decay = optimizer_type, float(lr), float(momentum), float(weight_decay)
        self.max_grad_norm, self.betas, self.eps = float(max_grad_norm), (float(betas[0]), float(betas[1])), float(eps)
        self.grads = dp_step.buf
        device = self.grads.flat.device
        if not self.grads.flat.is_cuda:
            raise _lib.NnueError("FusedOptimizer runs on CUDA only (no CPU fallback)")
        self.params = FlatParamBuffer(dict(dp_step.model.named_parameters()), device)
        assert self.params.names == self.grads.names
        self.n = self.params.numel()
        self.state1 = torch.zeros(self.n, dtype=torch.float32, device=device)  # momentum buffer / exp_avg
        self.state2 = torch.zeros(self.n, dtype=torch.float32, device=device) if optimizer_type == "adam" else None
        self.sqnorm = torch.zeros(1, dtype=torch.float32, device=device)
        self.ws = torch.empty(int(_lib.lib().nnue_opt_workspace_bytes(self.n)), dtype=torch.uint8, device=device)
        self.steps = 0

    def grad_norm(self):
        """Total gradient 2-norm as a 0-d device tensor (what clip_grad_norm_ returns)."""
        return self.sqnorm.sqrt().reshape(())

    def step(self):
        L, d, st = self._lib.lib(), self._lib.dptr, self._lib.stream_ptr()
        g = self.grads.flat[: self.n]
        if self.max_grad_norm > 0:
            self._lib.check(L.nnue_opt_grad_sqnorm(self.n, d(g), d(self.sqnorm), d(self.ws), self.ws.numel(), st))
        self.steps += 1
        if self.kind == "sgd":
            self._lib.check(L.nnue_opt_sgd_step(self.n, d(self.params.flat), d(g), d(self.state1), self.lr, self.momentum,
                                                self.weight_decay, self.max_grad_norm, d(self.sqnorm),
                                                1 if self.steps == 1 else 0, st))
        else:
            self._lib.check(L.nnue_opt_adam_step(self.n, d(self.params.flat), d(g), d(self.state1), d(self.state2),
                                                 self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                                 self.steps, self.max_grad_norm, d(self.sqnorm), st))


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

    def __init__(self, model, process_group=None, local_step: Optional[Callable] = None, device=None,
                 cuda_graphs: Optional[bool] = None, max_graphs: int = 8, allreduce: str = "auto"):
        self.model = model
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        named = dict(model.named_parameters())
        device = device if device is not None else named["input.weight"].device
        self.buf = FlatGradBuffer(named, device)
        self.buf.attach()
        self._local = local_step if local_step is not None else _cuda_local_step(model)
        # The step is ~20 short launches on two streams: replaying it as ONE CUDA graph removes the host launch cost
        # and the gaps between dependent kernels.  A graph is captured per (input buffers, parameter storages, shape)
        # the second time that combination is seen -- training loops that cycle a few device staging buffers hit the
        # cache -- and everything it allocates comes from one shared pool.
        if cuda_graphs is None:
            cuda_graphs = local_step is None and torch.device(device).type == "cuda"
        self.cuda_graphs = bool(cuda_graphs)
        self.max_graphs = int(max_graphs)
        self._graphs = {}   # key -> [hits, CUDAGraph or None, (images, labels) kept alive]
        self._pool = None
        # Exchange: the library's one-shot all-reduce over NVLink peer memory when the ranks share a node and
        # symmetric memory is available (in place on the flat buffer), otherwise one NCCL / gloo all-reduce.
        self.allreduce = "none" if self.world == 1 else "collective"
        # measured on 8 x B200 (config D, 216 KB per step): one-shot 233 / 237 us per step at 4 / 8 GPUs against 239 / 253 us
        # with NCCL; at 2 GPUs NCCL's two-rank path is 4 us ahead, so "auto" keeps the collective there
        want = allreduce == "oneshot" or (allreduce == "auto" and self.world >= 3)
        if self.world > 1 and local_step is None and torch.device(device).type == "cuda" and want:
            self._setup_oneshot(named, device)

    def _setup_oneshot(self, named, device):
        import ctypes
        import sys
        from . import _lib
        try:
            import torch.distributed._symmetric_memory as symm
            L = _lib.lib()
            group = self.group if self.group is not None else dist.group.WORLD
            if self.world > int(L.nnue_allreduce_max_world()):
                return
            n = self.buf.numel()  # (FlatGradBuffer pads its buffer to a multiple of 4 floats)
            recv = symm.empty(int(L.nnue_allreduce_recv_floats(self.world, n)), dtype=torch.float32, device=device)
            flags = symm.empty(int(L.nnue_allreduce_max_world()), dtype=torch.int32, device=device)
            recv.zero_()
            flags.zero_()
            h_recv = symm.rendezvous(recv, group)
            h_flags = symm.rendezvous(flags, group)
            torch.cuda.synchronize(device)
            dist.barrier(group=self.group)  # every rank's flags are zero before anybody publishes an epoch
            self._ar = dict(
                rank=int(h_recv.rank), n=n, epoch=0, keep=(recv, flags, h_recv, h_flags),
                recv=(ctypes.c_void_p * self.world)(*[int(x) for x in h_recv.buffer_ptrs]),
                flags=(ctypes.c_void_p * self.world)(*[int(x) for x in h_flags.buffer_ptrs]),
                counter=torch.zeros(1, dtype=torch.int32, device=device))
            self.allreduce = "oneshot_p2p"
        except Exception as e:  # no peer access / no symmetric memory on this system: the collective stays
            print(f"nnue_vision_b200: one-shot all-reduce unavailable ({type(e).__name__}: {e}); using the collective",
                  file=sys.stderr)

    def _exchange(self):
        if self.allreduce == "oneshot_p2p":
            from . import _lib
            a = self._ar
            a["epoch"] += 1
            _lib.check(_lib.lib().nnue_allreduce_oneshot(
                self.world, a["rank"], a["recv"], a["flags"], _lib.dptr(a["counter"]), a["n"], _lib.dptr(self.buf.flat),
                a["epoch"], _lib.stream_ptr()))
        else:
            dist.all_reduce(self.buf.flat, op=dist.ReduceOp.SUM, group=self.group)

    def _graph_key(self, images, labels, inv_count):
        return (images.data_ptr(), labels.data_ptr(), tuple(images.shape), images.dtype, labels.dtype, inv_count,
                tuple(p.data_ptr() for p in self.model.parameters()))

    def _run_local(self, images, labels, inv_count):
        ok = (self.cuda_graphs and images.is_cuda and labels.is_cuda and images.is_contiguous() and labels.is_contiguous()
              and images.dtype == torch.float32)
        if not ok:
            self._local(images, labels, inv_count, self.buf)
            return
        key = self._graph_key(images, labels, inv_count)
        ent = self._graphs.get(key)
        if ent is None:
            if len(self._graphs) >= self.max_graphs:  # drop the least recently used entry
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = [1, None, (images, labels)]
            self._local(images, labels, inv_count, self.buf)
            return
        self._graphs[key] = self._graphs.pop(key)  # most recently used last
        ent[0] += 1
        if ent[1] is None:
            if self._pool is None:
                self._pool = torch.cuda.graph_pool_handle()
            from . import _lib
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(images.device)
            n0 = int(_lib.lib().nnue_launch_count(0))
            try:
                with torch.cuda.graph(g, pool=self._pool):
                    self._local(images, labels, inv_count, self.buf)
            except Exception as e:  # a capture that cannot be taken must not cost the step: stay eager for good
                import sys
                print(f"nnue_vision_b200: CUDA-graph capture failed ({type(e).__name__}: {e}); running eagerly", file=sys.stderr)
                self.cuda_graphs = False
                self._graphs.clear()
                torch.cuda.synchronize(images.device)
                self._local(images, labels, inv_count, self.buf)
                return
            ent[1] = g
            ent.append(int(_lib.lib().nnue_launch_count(0)) - n0)  # kernels in the graph
            self._count_add = _lib.lib().nnue_launch_count_add
            g.replay()  # (the capture pass launched nothing but was counted: it stands for this replay)
            return
        ent[1].replay()
        self._count_add(ent[3])  # keep nnue_launch_count() meaning "kernels launched", replayed or not

    def step(self, images, labels, global_batch: Optional[int] = None, marks=None):
        if global_batch is None:
            global_batch = images.shape[0] * self.world  # equal shards
        if marks is not None:
            self._local(images, labels, 1.0 / float(global_batch), self.buf, marks)
        else:
            self._run_local(images, labels, 1.0 / float(global_batch))
        if self.world > 1:
            self._exchange()
        # optimizer.zero_grad() sets .grad to None by default: make the views the gradients again (ten assignments)
        self.buf.attach()
        return self.buf.loss.reshape(())
