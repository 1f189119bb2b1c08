"""Training-step surface of the hot path: `compute_loss` (train.py:250-254 of the reference) and
the batch-sharded data-parallel step.

The reference trains on one device (train.py:263) and has no collective; the only coupling
between samples is the mean-reduced loss and the parameter gradients, so the batch is split
across ranks, every rank scales its cross-entropy by 1/(global batch), the kernels write the
local gradients straight into ONE flat fp32 buffer, and a single NCCL all-reduce (sum) over
NVLink makes it the global mean gradient.  `clip_grad_norm_`
(train.py:363-364) must run after `step()` on the reduced buffer.
"""
import os
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
        # The conv / threshold gradients (the first two tensors) are the last to become final in a step; everything
        # behind them -- feature transformer, head, loss -- is final earlier.  `split` = first float4-aligned offset
        # behind the two: [split, end) can be exchanged while the conv gradient is still running, [0, split) at the end.
        early = sum(n for name, n in zip(self.names, sizes) if name in ("visual_threshold", "conv.weight"))
        self.split = min(total, (early + 3) // 4 * 4) if self.names[:2] == ["visual_threshold", "conv.weight"] else total

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
        with torch.cuda.device(self.grads.flat.device):  # the C side launches on the current device
            self._step()

    def _step(self):
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


class OneShotExchange:
    """The library's one-shot all-reduce (csrc/allreduce.cu) over two slices of the flat gradient buffer, each with its own
    symmetric receive area, flags and device-side step counter: `early` = [split, end) (feature transformer, head, loss),
    `late` = [0, split) (conv weights, thresholds).  No host argument changes from step to step, so both launches can be
    captured in the step's CUDA graph."""

    def __init__(self, buf: "FlatGradBuffer", world, group, device):
        import ctypes
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        L = _lib.lib()
        self._lib, self.world, self.buf = _lib, world, buf
        self.overlap = os.environ.get("NNUE_EXCHANGE_OVERLAP", "1") != "0"
        self.split_phases = os.environ.get("NNUE_EXCHANGE_SPLIT", "1") != "0"  # push early, collect at the end (one wait point)
        group = group if group is not None else dist.group.WORLD
        self.parts = []
        for lo, hi in ((buf.split, buf.numel()), (0, buf.split)):
            n = hi - lo
            if n <= 0:
                self.parts.append(None)
                continue
            recv = symm.empty(int(L.nnue_allreduce_recv_floats(world, n)), dtype=torch.float32, device=device)
            flags = symm.empty(int(L.nnue_allreduce_max_world()), dtype=torch.int32, device=device)
            recv.zero_()
            flags.zero_()
            h_recv = symm.rendezvous(recv, group)
            h_flags = symm.rendezvous(flags, group)
            self.rank = int(h_recv.rank)
            self.parts.append(dict(
                lo=lo, n=n, keep=(recv, flags, h_recv, h_flags),
                recv=(ctypes.c_void_p * world)(*[int(x) for x in h_recv.buffer_ptrs]),
                flags=(ctypes.c_void_p * world)(*[int(x) for x in h_flags.buffer_ptrs]),
                state=torch.zeros(2, dtype=torch.int32, device=device)))
        torch.cuda.synchronize(device)
        dist.barrier(group=group)  # every rank's flags are zero before anybody publishes an epoch

    def _launch(self, part, stream):
        if part is None:
            return
        _lib = self._lib
        _lib.check(_lib.lib().nnue_allreduce_oneshot(
            self.world, self.rank, part["recv"], part["flags"], _lib.dptr(part["state"]), part["n"],
            _lib.dptr(self.buf.flat[part["lo"]:part["lo"] + part["n"]]), stream))

    def _slices(self, jobs, stream):
        """One launch over (part, phase) jobs of the flagged form: phase 0 = whole exchange, 1 = push, 2 = collect."""
        import ctypes
        _lib = self._lib
        arr = (_lib.AllreduceSlice * len(jobs))()
        for a, (part, phase) in zip(arr, jobs):
            a.peer_recv_h = ctypes.cast(part["recv"], ctypes.c_void_p)
            a.peer_flags_h = ctypes.cast(part["flags"], ctypes.c_void_p)
            a.state_d = part["state"].data_ptr()
            a.n = part["n"]
            a.buf_d = self.buf.flat[part["lo"]:part["lo"] + part["n"]].data_ptr()
            a.phase = phase
        _lib.check(_lib.lib().nnue_allreduce_oneshot_slices(self.world, self.rank, arr, len(jobs), stream))

    def _split_ok(self):
        # the early slice is pushed where it becomes final and collected at the end of the step together with the late
        # slice's whole exchange (csrc/allreduce.cu): needs the flagged form for both
        ll = int(self._lib.lib().nnue_allreduce_ll_max_floats())
        return (self.overlap and self.split_phases and all(p is not None and p["n"] <= ll for p in self.parts))

    def early(self, stream=None):
        stream = stream if stream is not None else self._lib.stream_ptr()
        if self._split_ok():
            self._slices([(self.parts[0], 1)], stream)
        elif self.overlap:
            self._launch(self.parts[0], stream)

    def late(self, stream=None):
        stream = stream if stream is not None else self._lib.stream_ptr()
        if self._split_ok():
            self._slices([(self.parts[1], 0), (self.parts[0], 2)], stream)
            return
        if not self.overlap:  # both slices at the end, on the caller's stream
            self._launch(self.parts[0], stream)
        self._launch(self.parts[1], stream)

    def full(self, stream=None):
        self._launch(self.parts[0], stream if stream is not None else self._lib.stream_ptr())
        self._launch(self.parts[1], stream if stream is not None else self._lib.stream_ptr())


class CollectiveExchange:
    """NCCL all-reduce of the same two slices (large buffers: the 268 MB table gradient of SURVEY config I rides a ring /
    NVLS reduction while the value and conv gradients are still being computed)."""

    def __init__(self, buf: "FlatGradBuffer", group):
        self.buf, self.group = buf, group

    def early(self, stream=None):
        if self.buf.split < self.buf.numel():
            dist.all_reduce(self.buf.flat[self.buf.split:], op=dist.ReduceOp.SUM, group=self.group)

    def late(self, stream=None):
        if self.buf.split > 0:
            dist.all_reduce(self.buf.flat[:self.buf.split], op=dist.ReduceOp.SUM, group=self.group)

    def full(self, stream=None):
        dist.all_reduce(self.buf.flat, op=dist.ReduceOp.SUM, group=self.group)


def _cuda_local_step(model):
    """The B200 hot path: forward, fused CE, backward -- every kernel through the C ABI."""
    from . import _lib, nnue as _nnue

    def run(images, labels, inv_count, buf: FlatGradBuffer, marks=None, exchange=None):
        images = model._check_images(images)
        labels = labels.to(device=images.device, dtype=torch.long).contiguous()
        fs = model.feature_set
        B, _, H, W = images.shape
        params = tuple(p.detach().contiguous() for p in model._hot_params())
        shape = _lib.make_shape(B, H, W, fs.num_features_per_square, fs.grid_size, model.l1_size, model.l2_size,
                                model.l3_size, model.num_classes, model.conv.stride[0])
        # gradients land directly in the flat buffer's views, the loss in its trailing slot
        _nnue._run_train_step(shape, images, labels, params, inv_count, grads=buf.views, loss_out=buf.loss,
                              marks=marks, exchange=exchange)

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
                 cuda_graphs: Optional[bool] = None, max_graphs: int = 8, allreduce: str = "auto", attach: bool = True,
                 single: bool = False):
        self.model = model
        self.group = process_group
        self.world = 1 if single else (dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1)
        named = dict(model.named_parameters())
        device = torch.device(device if device is not None else named["input.weight"].device)
        self.device = device
        self.buf = FlatGradBuffer(named, device)
        self._attach = bool(attach)  # False: the buffer is private scratch of the module-level loss (nnue.NNUE.loss)
        if self._attach:
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
        self._graphs = {}   # key -> [hits, CUDAGraph or None, kernels in the graph]
        self._pool = None
        # Exchange (inside the step, overlapped: see _run_train_step): the library's one-shot all-reduce over NVLink peer
        # memory for small buffers when the ranks share a node and symmetric memory is available -- it has no per-step
        # host argument and is captured in the step's CUDA graph -- otherwise NCCL (gloo on CPU) all-reduces of the same
        # slices, issued eagerly on the step's streams (a step with a 268 MB gradient is not launch-bound).
        self.allreduce = "none" if self.world == 1 else "collective"
        self._xchg = None
        self._inline = local_step is None and device.type == "cuda" and self.world > 1  # the exchange runs inside the step
        small = self.buf.numel() * 4 <= (4 << 20)
        want = allreduce == "oneshot" or (allreduce == "auto" and small)
        if self._inline and want:
            self._setup_oneshot(named, device)
        if self._inline and self._xchg is None:
            self._xchg = CollectiveExchange(self.buf, self.group)
            self.cuda_graphs = False  # (NCCL calls stay outside graph capture)

    def _setup_oneshot(self, named, device):
        import sys
        from . import _lib
        try:
            if self.world > int(_lib.lib().nnue_allreduce_max_world()):
                return
            self._xchg = OneShotExchange(self.buf, self.world, self.group, device)
            self.allreduce = "oneshot_p2p"
        except Exception as e:  # no peer access / no symmetric memory on this system: the collective stays
            print(f"nnue_vision_b200: one-shot all-reduce unavailable ({type(e).__name__}: {e}); using the collective",
                  file=sys.stderr)

    def _exchange(self):
        """The whole buffer, now, on the current stream (the step itself issues the two slices where they become final)."""
        if self._xchg is not None:
            self._xchg.full()
        else:
            dist.all_reduce(self.buf.flat, op=dist.ReduceOp.SUM, group=self.group)

    def _graph_key(self, images, labels, inv_count):
        # Everything the captured launches depend on: buffer addresses, shapes, the scaling, the parameter storages, and
        # the library / module switches that select kernels.  No tensor is kept alive by the cache: a replay only needs
        # the addresses to hold the caller's CURRENT inputs, which the key guarantees (when the allocator hands a new
        # batch the address of a freed one, the graph is simply reused).
        from . import _lib, nnue as _nnue
        return (images.data_ptr(), labels.data_ptr(), tuple(images.shape), images.dtype, labels.dtype, inv_count,
                tuple(p.data_ptr() for p in self.model.parameters()), _lib.options_epoch(), _nnue.STORE_ACTIVATIONS,
                _nnue.PREFORMAT_TABLES, _nnue.OVERLAP_TABLE_GRADIENT, _nnue.HEAD_SIDE_STREAM)

    def _call_local(self, images, labels, inv_count, marks=None):
        if self._inline:  # the step issues the exchange itself, slice by slice, where the gradients become final
            self._local(images, labels, inv_count, self.buf, marks, self._xchg)
        elif marks is not None:
            self._local(images, labels, inv_count, self.buf, marks)
        else:
            self._local(images, labels, inv_count, self.buf)

    def _run_local(self, images, labels, inv_count):
        ok = (self.cuda_graphs and images.is_cuda and labels.is_cuda and images.is_contiguous() and labels.is_contiguous()
              and images.dtype == torch.float32)
        if not ok:
            self._call_local(images, labels, inv_count)
            return
        key = self._graph_key(images, labels, inv_count)
        ent = self._graphs.get(key)
        if ent is None:
            if len(self._graphs) >= self.max_graphs:  # drop the least recently used entry
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = [1, None, 0]
            self._call_local(images, labels, inv_count)
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
                    self._call_local(images, labels, inv_count)
            except Exception as e:  # a capture that cannot be taken must not cost the step: stay eager for good
                import sys
                print(f"nnue_vision_b200: CUDA-graph capture failed ({type(e).__name__}: {e}); running eagerly", file=sys.stderr)
                self.cuda_graphs = False
                self._graphs.clear()
                torch.cuda.synchronize(images.device)
                self._call_local(images, labels, inv_count)
                return
            ent[1] = g
            ent[2] = int(_lib.lib().nnue_launch_count(0)) - n0  # kernels in the graph
            self._count_add = _lib.lib().nnue_launch_count_add
            g.replay()  # (the capture pass launched nothing but was counted: it stands for this replay)
            return
        ent[1].replay()
        self._count_add(ent[2])  # keep nnue_launch_count() meaning "kernels launched", replayed or not

    def run_local_flat(self, images, labels, inv_count):
        """This rank's step (graph replay when possible, no exchange); returns the flat buffer: every gradient in
        parameter order, then the loss.  Used by the module-level loss, which copies what it needs."""
        labels = labels.to(device=images.device, dtype=torch.long).contiguous()
        with torch.cuda.device(self.device):
            self._run_local(images, labels, inv_count)
        return self.buf.flat

    def step(self, images, labels, global_batch: Optional[int] = None, marks=None):
        if global_batch is None:
            global_batch = images.shape[0] * self.world  # equal shards
        import contextlib
        guard = torch.cuda.device(self.device) if self.device.type == "cuda" else contextlib.nullcontext()
        with guard:  # (the C side launches on the current device: a model on cuda:1 must work while cuda:0 is current)
            if marks is not None:
                self._call_local(images, labels, 1.0 / float(global_batch), marks)
            else:
                self._run_local(images, labels, 1.0 / float(global_batch))
            if self.world > 1 and not self._inline:
                self._exchange()
        # optimizer.zero_grad() sets .grad to None by default: make the views the gradients again (ten assignments)
        if self._attach:
            self.buf.attach()
        return self.buf.loss.reshape(())
