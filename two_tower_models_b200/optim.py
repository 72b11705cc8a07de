"""Fused Adam for the drop-in modules (the step of the training loop next to the hot path).

The reference trains with `torch.optim.Adam(model.parameters(), lr)` on DENSE gradients (train/train.py:179,
123-125): every row of the two [hash, D] embedding tables moves every step (momentum), so the update is a plain
HBM-bound elementwise pass over all parameters.  `FusedAdam` keeps torch's semantics and state layout
(`state[p] = {step, exp_avg, exp_avg_sq}`) and does the whole update in ONE launch of `tt_adam_step`
(28 bytes per element; ~0.7 GB per step at the benchmark config).  The step counter is a device scalar, so the
optimizer step can be captured in the same CUDA graph as forward + backward.
"""
from typing import Iterable, Tuple

import torch

from . import _native


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr, betas, eps, weight_decay) with amsgrad / maximize / foreach off."""

    def __init__(self, params: Iterable, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 0.0) -> None:
        if lr < 0.0 or eps < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or weight_decay < 0.0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._sync = {}  # device -> (step int64[1], ticket int32[1])

    def _device_state(self, device, group):
        """(step int64[1], ticket int32[1]) of one parameter group, on its device."""
        key = (device, id(group))
        st = self._sync.get(key)
        if st is None:
            st = (torch.zeros(1, dtype=torch.int64, device=device), torch.zeros(1, dtype=torch.int32, device=device))
            self._sync[key] = st
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _native.lib()
        for group in self.param_groups:
            todo = []
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.is_sparse:
                    raise RuntimeError("FusedAdam: fp32 CUDA parameters with dense gradients only (no CPU fallback)")
                if not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("FusedAdam: parameters and gradients must be contiguous")
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = self._device_state(p.device, group)[0]  # shared int64 device counter
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                todo.append(p)
            if not todo:
                continue
            dev = todo[0].device
            if any(p.device != dev for p in todo):
                raise RuntimeError("FusedAdam: one device per parameter group")
            step_dev, ticket = self._device_state(dev, group)
            beta1, beta2 = group["betas"]
            stream = torch.cuda.current_stream(dev).cuda_stream
            for i0 in range(0, len(todo), 32):
                if i0:  # later launches of the same step re-use step t: rewind the counter the previous one bumped
                    step_dev.sub_(1)
                chunk = todo[i0:i0 + 32]
                arr = (_native.AdamTensor * len(chunk))()
                for a, p in zip(arr, chunk):
                    st = self.state[p]
                    a.param, a.grad = p.data_ptr(), p.grad.data_ptr()
                    a.exp_avg, a.exp_avg_sq, a.numel = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()
                _native.check(L.tt_adam_step(arr, len(chunk), group["lr"], beta1, beta2, group["eps"],
                                             group["weight_decay"], step_dev.data_ptr(), ticket.data_ptr(), stream),
                              "adam_step")
                # the kernel wrote the parameters through raw pointers: tell autograd (and the bf16 operand caches of
                # ops.PackedWeights, which key on `_version`) that they changed
                torch.autograd.graph.increment_version(chunk)
        return loss
