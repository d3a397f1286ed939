"""Synchronous data-parallel training plumbing: every rank rolls out and samples its own
episodes, computes the loss gradient of its minibatch, the gradients are summed with ONE NCCL all-reduce over a single
flat fp32 buffer (L-DGN: 1 005 315 parameters = 4.0 MB), divided by the world size inside the fused Adam kernel, and
every rank applies the identical update -- weights stay bit-identical across ranks without ever being broadcast again
after the start.  The reference is single-process (one model, l_dgn.py:66-76); this is what its trainer becomes when
the rollout is sharded over the GPUs of a box.

``FlatParameters`` re-homes a module's parameters (and their gradients) as views into two flat buffers, so autograd
accumulates straight into the all-reduce buffer; ``GradSync`` is the collective (NCCL on GPUs, gloo in the CPU tests);
``FusedAdam`` is torch.optim.Adam's update as one CUDA kernel over the flat buffers (``mls_adam_step``).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib


class FlatParameters:
    """All trainable parameters of ``module`` as views of one flat fp32 buffer (``.flat``), their gradients as views of
    another (``.grad``).  state_dict keys / shapes are unchanged."""

    def __init__(self, module: torch.nn.Module):
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise ValueError("module has no trainable parameters")
        dev, dt = params[0].device, params[0].dtype
        if any(p.device != dev or p.dtype != dt for p in params):
            raise ValueError("parameters must share one device and dtype")
        self.module, self.params = module, params
        # every parameter starts on a 64-byte boundary (the CUDA kernels read weight rows with 16-byte loads); the
        # padding stays zero in both buffers, so the optimiser never moves it
        ALIGN = 16
        self.offsets, off = [], 0
        for p in params:
            self.offsets.append(off)
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.numel = off
        self.flat = torch.zeros(self.numel, dtype=dt, device=dev)
        self.grad = torch.zeros(self.numel, dtype=dt, device=dev)
        with torch.no_grad():
            for p, off in zip(params, self.offsets):
                n = p.numel()
                self.flat[off:off + n].copy_(p.detach().reshape(-1))
                p.data = self.flat[off:off + n].view_as(p)
                p.grad = self.grad[off:off + n].view_as(p)

    def zero_grad(self):
        self.grad.zero_()
        for p in self.params:                      # a backward that replaced .grad (it should not) is re-homed
            if p.grad is None or p.grad.data_ptr() < self.grad.data_ptr() or \
                    p.grad.data_ptr() >= self.grad.data_ptr() + self.grad.numel() * self.grad.element_size():
                self._rehome()
                break

    def _rehome(self):
        for p, off in zip(self.params, self.offsets):
            p.grad = self.grad[off:off + p.numel()].view_as(p)

    def bump_versions(self):
        """The fused Adam kernel writes the flat buffer behind torch's back: tell caches keyed on the parameter
        versions (the packed bf16 copy of the rollout forward) that the weights changed."""
        mark = getattr(self.module, "mark_parameters_changed", None)
        if mark is not None:
            mark()
        else:
            with torch.no_grad():
                for p in self.params:
                    p.add_(0)


class GradSync:
    """Sum of the flat gradient over all ranks: one collective per optimiser step.  ``start()`` launches it
    asynchronously (NCCL runs it on its own stream, so rollout rounds issued meanwhile overlap it), ``finish()`` makes
    the current stream wait for it.  The mean is taken by the optimiser (``grad_scale = 1 / world``)."""

    def __init__(self, flat_grad: torch.Tensor, group=None):
        self.buf, self.group = flat_grad, group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self._work = None

    @property
    def grad_scale(self) -> float:
        return 1.0 / self.world

    def start(self):
        if self.world > 1:
            self._work = dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        if self._work is not None:
            self._work.wait()
            self._work = None

    def broadcast_parameters(self, flat_param: torch.Tensor, src: int = 0):
        """Once, at the start: every rank adopts rank ``src``'s initial weights."""
        if self.world > 1:
            dist.broadcast(flat_param, src=src, group=self.group)


class FusedAdam:
    """torch.optim.Adam(params, lr, betas, eps, weight_decay) over ``FlatParameters`` as one kernel launch."""

    def __init__(self, flat: FlatParameters, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        if flat.flat.device.type != "cuda" or flat.flat.dtype != torch.float32:
            raise _lib.MelissaLibraryError("FusedAdam needs float32 CUDA parameters (no CPU fallback)")
        self.lib = _lib.lib()
        self.flat = flat
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.exp_avg = torch.zeros_like(flat.flat)
        self.exp_avg_sq = torch.zeros_like(flat.flat)
        self.step_count = 0

    def zero_grad(self, set_to_none: bool = False):
        self.flat.zero_grad()

    def step(self, grad_scale: float = 1.0):
        self.step_count += 1
        f = self.flat
        _lib.check(self.lib.mls_adam_step(f.flat.data_ptr(), f.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                          f.numel, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                          self.step_count, float(grad_scale), _lib.current_stream_ptr()))
        f.bump_versions()

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "lr": self.lr,
                "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
