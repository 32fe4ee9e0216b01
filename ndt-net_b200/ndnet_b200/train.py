"""Training step of the reference's networks on the CUDA library (include/ndnet_b200.h (4)).

`NetTrainer(module)` (alias `SegTrainer`) binds an `ndnet.models.ndtnet.NDTNetSegmentation` / `NDTNetClassification` or
an `ndnet.models.pointnet.PointNetSegmentation` / `PointNetClassification` (the reference's class layouts and state_dict
keys, /root/reference/ndnet/models/ndtnet.py:166-243, pointnet.py:137-214) to `ndnet_b200_trainer_*`; calling it in place
of `module(points, covs)` / `module(points)` inside the reference's training loops (/root/reference/tools/train.py:66-76,
tools/train_pointnet.py) gives the same train-mode forward (batch-statistics BatchNorm, running statistics updated in
place) and, through one `torch.autograd.Function`, the gradients of every parameter from our kernels instead of torch
autograd.  The loss and the optimizer stay ordinary torch code.

Multi-GPU (one process per GPU, scans sharded over the ranks): the gradient average is the training path's only collective
(SURVEY.md §8e).  `SegTrainer(..., overlap_allreduce=True)` (or `module.b200_overlap_allreduce = True`) launches it from
INSIDE the backward pass: the library lays the gradients out in the order backward finishes them and records an event per
bucket (head | trunk + feature T-Net | first layer + input T-Net); each bucket's NCCL all-reduce starts on a side stream
as soon as its event fires, while the kernels of the earlier layers still run.  `allreduce_gradients(module)` is the
plain version (one flat all-reduce after backward) and a no-op for gradients that were already averaged in backward.
BatchNorm: batch statistics are per replica, and every rank updates its own running statistics (stock DDP instead
broadcasts rank 0's buffers at every forward); call `sync_batchnorm_buffers(module)` before evaluating or saving a
checkpoint so that all ranks hold the same (averaged) running statistics.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class SegTrainer:
    def __init__(self, module: torch.nn.Module, device: torch.device | int | None = None, tf32: bool = False, graph: bool = True,
                 overlap_allreduce: bool = False):
        """tf32=False: fp32 FMA everywhere (parity configuration).  tf32=True: the large GEMMs run on the tensor cores
        (tcgen05 kind::tf32, fp32 accumulation) — the counterpart of torch.backends.cuda.matmul.allow_tf32.
        graph=True: each pass is captured into a CUDA graph per (shape, parameter storage) and replayed.
        overlap_allreduce=True: with an initialised process group of more than one rank, backward averages the gradients
        over the ranks itself, bucket by bucket, overlapped with the rest of the backward pass."""
        self.module = module
        p0 = next(module.parameters())
        dev = p0.device if device is None else (torch.device("cuda", device) if isinstance(device, int) else torch.device(device))
        if dev.type != "cuda":
            raise RuntimeError("ndnet_b200 trains on a CUDA device only: move the module to the GPU first")
        self.device = dev
        self._L = _lib.lib()
        self.param_names = [n for n, _ in module.named_parameters()]
        self.buffer_names = [n for n, _ in module.named_buffers()]
        self.names = self.param_names + self.buffer_names
        tensors = self._tensors()
        for n, t in zip(self.names, tensors):
            if n in self.param_names and t.dtype != torch.float32:
                raise RuntimeError(f"{n}: the training kernels are fp32")
        n = len(self.names)
        shapes = [np.array(tuple(t.shape) if t.dim() else (1,), np.int64) for t in tensors]
        c_names = (C.c_char_p * n)(*[k.encode() for k in self.names])
        c_shapes = (C.c_void_p * n)(*[s.ctypes.data for s in shapes])
        c_nd = (C.c_int * n)(*[len(s) for s in shapes])
        h = C.c_void_p()
        rc = self._L.ndnet_b200_trainer_create(dev.index or 0, n, c_names, c_shapes, c_nd, C.byref(h))
        if rc != 0:
            raise RuntimeError(f"ndnet_b200_trainer_create failed ({rc}): not an NDTNet / PointNet segmentation or classification state_dict?")
        self._h = h
        kind, width, outs = C.c_int(0), C.c_int(0), C.c_int(0)
        self._L.ndnet_b200_trainer_info(h, C.byref(kind), C.byref(width), C.byref(outs))
        self.kind, self.point_width = int(kind.value), int(width.value)
        self.segmentation, self.takes_covariances = self.kind in (1, 3), self.kind in (0, 1)
        self.tf32 = bool(tf32)
        self._L.ndnet_b200_trainer_set_precision(h, int(self.tf32))
        self._L.ndnet_b200_trainer_set_graph(h, int(bool(graph)))
        offs = np.full(n, -1, np.int64)
        self._flat_elems = int(self._L.ndnet_b200_trainer_grad_layout(h, offs.ctypes.data, n))
        self._grad_off = [int(o) for o in offs]
        assert all((o >= 0) == (k in set(self.param_names)) for o, k in zip(self._grad_off, self.names)), "gradient layout mismatch"
        self.num_out = int(outs.value)              # segmentation: num_classes + 1 per row; classification: num_classes per cloud
        self.overlap_allreduce = bool(overlap_allreduce)
        self._buckets = []
        for i in range(int(self._L.ndnet_b200_trainer_num_buckets(h))):
            b, e = C.c_long(0), C.c_long(0)
            self._L.ndnet_b200_trainer_bucket_range(h, i, C.byref(b), C.byref(e))
            self._buckets.append((int(b.value), int(e.value)))
        self._side = None                      # stream the bucket all-reduces are launched from
        self.last_allreduce = None             # (bytes, number of collectives) of the last overlapped backward

    def _tensors(self):
        params = dict(self.module.named_parameters())
        buffers = dict(self.module.named_buffers())
        return [params[n] if n in params else buffers[n] for n in self.names]

    def _ptr_array(self, tensors):
        return (C.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self._L.ndnet_b200_trainer_last_error(self._h).decode()}")

    def __call__(self, points: torch.Tensor, covariances: torch.Tensor | None = None) -> torch.Tensor:
        """NDT networks: (B,N,3), (B,N,9); PointNet: (B,N,point_dim).  Returns what the module's forward returns -
        segmentation: log-probabilities (B,N,num_classes+1); classification: probabilities (B,num_classes,1) -
        differentiable w.r.t. the module's parameters."""
        if not self.module.training:
            raise RuntimeError("the trainer normalises with batch statistics and updates the running ones: the module is in eval() "
                               "mode - call module(...) / forward_b200 (the folded inference kernels) instead")
        if self.takes_covariances:
            if covariances is None:
                raise TypeError("this network takes (points, covariances)")
            feat = torch.cat((points, covariances), dim=2).float().contiguous()
        else:
            feat = points.float().contiguous()
        if feat.shape[2] != self.point_width:
            raise RuntimeError(f"rows of width {feat.shape[2]} for a network built for {self.point_width}")
        out = _SegTrainFn.apply(self, feat, *[p for _, p in self.module.named_parameters()])
        return out if self.segmentation else out.unsqueeze(2)

    def debug_buffer(self, name: str) -> torch.Tensor:
        """Flat copy of an internal buffer of the last pass (test hook), e.g. "h3.dA", "t2.c1.Y", "t1.T"."""
        n = self._L.ndnet_b200_trainer_debug_buffer(self._h, name.encode(), None, None)
        if n < 0:
            raise KeyError(name)
        out = torch.empty((n,), dtype=torch.float32, device=self.device)
        self._L.ndnet_b200_trainer_debug_buffer(self._h, name.encode(), out.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
        return out

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.ndnet_b200_trainer_destroy(self._h)
                self._h = None
        except Exception:
            pass


NetTrainer = SegTrainer


class _SegTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, trainer: SegTrainer, feat: torch.Tensor, *params):
        B, N, _ = feat.shape
        tensors = trainer._tensors()
        for t in tensors:
            if not (t.is_cuda and t.is_contiguous()):
                raise RuntimeError("parameters and buffers must be contiguous CUDA tensors")
        out = torch.empty((B, N, trainer.num_out) if trainer.segmentation else (B, trainer.num_out), dtype=torch.float32, device=feat.device)
        stream = torch.cuda.current_stream(feat.device).cuda_stream
        rc = trainer._L.ndnet_b200_trainer_forward(trainer._h, feat.data_ptr(), B, N, trainer._ptr_array(tensors), out.data_ptr(),
                                                   int(trainer.module.training), stream)
        trainer._check(rc, "ndnet_b200_trainer_forward")
        trainer._pass = getattr(trainer, "_pass", 0) + 1      # the library keeps the activations of ONE forward pass
        ctx.pass_id = trainer._pass
        ctx.trainer = trainer
        ctx.save_for_backward(feat)
        return out

    @staticmethod
    def backward(ctx, dout: torch.Tensor):
        trainer: SegTrainer = ctx.trainer
        if ctx.pass_id != trainer._pass:
            raise RuntimeError("ndnet_b200 trainer: backward of an earlier forward pass - its activations were overwritten by a "
                               "later forward through the same module (run backward before the next forward, or use one "
                               "SegTrainer per in-flight pass)")
        (feat,) = ctx.saved_tensors                      # keeps the input alive: the library reads it again
        tensors = trainer._tensors()
        n_params = len(trainer.param_names)
        flat = torch.empty((trainer._flat_elems,), dtype=torch.float32, device=dout.device)    # fresh storage every pass
        dout = dout.float().contiguous()
        stream = torch.cuda.current_stream(dout.device).cuda_stream
        import torch.distributed as dist
        overlap = trainer.overlap_allreduce and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        trainer._L.ndnet_b200_trainer_set_deferred_copy(trainer._h, int(overlap))
        rc = trainer._L.ndnet_b200_trainer_backward_flat(trainer._h, dout.data_ptr(), trainer._ptr_array(tensors), flat.data_ptr(), stream)
        trainer._check(rc, "ndnet_b200_trainer_backward_flat")
        if overlap:
            # everything above is only enqueued.  Per bucket: the side stream waits for the bucket's event (and, with
            # CUDA graphs, copies the bucket out of the library's buffer), then NCCL averages it - while the main
            # stream is still running the backward kernels of the earlier layers
            if trainer._side is None:
                trainer._side = torch.cuda.Stream(dout.device)
            side = trainer._side
            flat.record_stream(side)
            works = []
            for i, (b, e) in enumerate(trainer._buckets):
                rc = trainer._L.ndnet_b200_trainer_bucket_ready(trainer._h, i, flat.data_ptr(), side.cuda_stream)
                trainer._check(rc, "ndnet_b200_trainer_bucket_ready")
                with torch.cuda.stream(side):
                    works.append(dist.all_reduce(flat[b:e], op=dist.ReduceOp.AVG, async_op=True))
            for w in works:
                w.wait()                          # the CURRENT stream waits for the collective; the host does not block
            trainer.last_allreduce = (4 * trainer._flat_elems, len(works))
            for t in tensors[:n_params]:
                t._b200_grad_averaged = True      # allreduce_gradients() skips these
        grads = [flat[o:o + t.numel()].view_as(t) for o, t in zip(trainer._grad_off[:n_params], tensors[:n_params])]
        return (None, None, *grads)


def debug_gemm(mode: int, A: torch.Tensor, B: torch.Tensor, C: torch.Tensor, bias: torch.Tensor | None = None,
               accumulate: bool = False) -> int:
    """Test hook: C (+)= A @ B.T (+ bias) through the training GEMM kernels (0 = fp32 FMA, 1 = tcgen05 TF32).  2-D CUDA fp32
    tensors whose last dimension is contiguous (row strides may exceed the width).  Returns the library's status code."""
    L = _lib.lib()
    M, K = A.shape
    N = B.shape[0]
    assert B.shape[1] == K and tuple(C.shape) == (M, N) and A.stride(1) == B.stride(1) == C.stride(1) == 1
    return L.ndnet_b200_debug_train_gemm(mode, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), C.data_ptr(), C.stride(0), M, N, K,
                                         bias.data_ptr() if bias is not None else None, int(accumulate),
                                         torch.cuda.current_stream(A.device).cuda_stream)


def allreduce_gradients(module: torch.nn.Module, world_size: int | None = None) -> int:
    """Averages the gradients of `module` over the ranks with ONE all-reduce of a flat fp32 buffer (≈3.5 M values for
    NDTNetSegmentation) — the training path's only collective.  Works on NCCL (GPU) and gloo (the CPU tests).
    Returns the number of values reduced; a no-op without an initialised process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    world = dist.get_world_size() if world_size is None else world_size
    # every parameter that takes gradients takes part on every rank (a missing gradient counts as zeros), so that the ranks
    # always issue the same collective; parameters whose gradients backward already averaged are left alone
    params = [p for p in module.parameters() if p.requires_grad]
    if all(getattr(p, "_b200_grad_averaged", False) for p in params):
        for p in params:
            p._b200_grad_averaged = False
        return 0
    if not params or world == 1:
        return 0
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= world
    off = 0
    for p in params:
        n = p.numel()
        if p.grad is None:
            p.grad = flat[off:off + n].view_as(p).clone()
        else:
            p.grad.copy_(flat[off:off + n].view_as(p))
        off += n
    return off


def sync_batchnorm_buffers(module: torch.nn.Module) -> int:
    """Averages the BatchNorm running statistics over the ranks (num_batches_tracked: maximum).  Training uses per-replica
    batch statistics and every rank updates its own running statistics, so the ranks drift apart; call this before
    evaluating or saving a checkpoint (stock DDP broadcasts rank 0's buffers at every forward instead).  Returns the number
    of values reduced; a no-op without an initialised process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    floats = [b for b in module.buffers() if b.is_floating_point()]
    ints = [b for b in module.buffers() if not b.is_floating_point()]
    n = 0
    if floats:
        flat = torch.cat([b.reshape(-1).float() for b in floats])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat /= dist.get_world_size()
        off = 0
        for b in floats:
            b.copy_(flat[off:off + b.numel()].view_as(b))
            off += b.numel()
        n += off
    if ints:
        flat = torch.stack([b.reshape(()).long() for b in ints])
        dist.all_reduce(flat, op=dist.ReduceOp.MAX)
        for b, v in zip(ints, flat):
            b.copy_(v)
        n += len(ints)
    return n


def reference_loss(pred: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """The loss the reference's loop applies to the model output (tools/train.py:73):
    `cross_entropy(pred, gt)` on (B, N, C+1) tensors, i.e. class-probability targets with dim 1 as the class axis."""
    return torch.nn.functional.cross_entropy(pred, gt)
