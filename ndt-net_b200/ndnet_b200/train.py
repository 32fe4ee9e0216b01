"""Training step of NDTNetSegmentation on the CUDA library (include/ndnet_b200.h (4)).

`SegTrainer(module)` binds an `ndnet.models.ndtnet.NDTNetSegmentation` (the reference's class layout and state_dict
keys, /root/reference/ndnet/models/ndtnet.py:198-243) to `ndnet_b200_trainer_*`; calling it in place of
`module(points, covs)` inside the reference's training loop (/root/reference/tools/train.py:66-76) gives the same
train-mode forward (batch-statistics BatchNorm, running statistics updated in place) and, through one
`torch.autograd.Function`, the gradients of every parameter from our kernels instead of torch autograd.  The loss and
the optimizer stay ordinary torch code.  Multi-GPU: one process per GPU, `allreduce_gradients` averages the gradients
with one flat NCCL all-reduce (the only collective of the training path).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class SegTrainer:
    def __init__(self, module: torch.nn.Module, device: torch.device | int | None = None, tf32: bool = False, graph: bool = True):
        """tf32=False: fp32 FMA everywhere (parity configuration).  tf32=True: the large GEMMs run on the tensor cores
        (tcgen05 kind::tf32, fp32 accumulation) — the counterpart of torch.backends.cuda.matmul.allow_tf32.
        graph=True: each pass is captured into a CUDA graph per (shape, parameter storage) and replayed."""
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
            raise RuntimeError(f"ndnet_b200_trainer_create failed ({rc}): not an NDTNetSegmentation state_dict?")
        self._h = h
        self.tf32 = bool(tf32)
        self._L.ndnet_b200_trainer_set_precision(h, int(self.tf32))
        self._L.ndnet_b200_trainer_set_graph(h, int(bool(graph)))
        offs = np.full(n, -1, np.int64)
        self._flat_elems = int(self._L.ndnet_b200_trainer_grad_layout(h, offs.ctypes.data, n))
        self._grad_off = [int(o) for o in offs]
        assert all((o >= 0) == (k in set(self.param_names)) for o, k in zip(self._grad_off, self.names)), "gradient layout mismatch"
        self.num_out = int(module.num_classes) + 1

    def _tensors(self):
        params = dict(self.module.named_parameters())
        buffers = dict(self.module.named_buffers())
        return [params[n] if n in params else buffers[n] for n in self.names]

    def _ptr_array(self, tensors):
        return (C.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self._L.ndnet_b200_trainer_last_error(self._h).decode()}")

    def __call__(self, points: torch.Tensor, covariances: torch.Tensor) -> torch.Tensor:
        """(B,N,3), (B,N,9) -> log-probabilities (B,N,num_classes+1), differentiable w.r.t. the module's parameters."""
        feat = torch.cat((points, covariances), dim=2).float().contiguous()
        return _SegTrainFn.apply(self, feat, *[p for _, p in self.module.named_parameters()])

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


class _SegTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, trainer: SegTrainer, feat: torch.Tensor, *params):
        B, N, _ = feat.shape
        tensors = trainer._tensors()
        for t in tensors:
            if not (t.is_cuda and t.is_contiguous()):
                raise RuntimeError("parameters and buffers must be contiguous CUDA tensors")
        out = torch.empty((B, N, trainer.num_out), dtype=torch.float32, device=feat.device)
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
        rc = trainer._L.ndnet_b200_trainer_backward_flat(trainer._h, dout.data_ptr(), trainer._ptr_array(tensors), flat.data_ptr(), stream)
        trainer._check(rc, "ndnet_b200_trainer_backward_flat")
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
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if not grads or world == 1:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= world
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return off


def reference_loss(pred: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """The loss the reference's loop applies to the model output (tools/train.py:73):
    `cross_entropy(pred, gt)` on (B, N, C+1) tensors, i.e. class-probability targets with dim 1 as the class axis."""
    return torch.nn.functional.cross_entropy(pred, gt)
