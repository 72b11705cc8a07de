"""CUDA-graph capture of one training step (train_forward + backward) of a drop-in model.

A BASELINE config-2 step is ~60 GFLOP, i.e. tens of microseconds of device work spread over a few dozen
kernels, so launching them one by one from Python leaves the GPU idle most of the time.  Every native
entry point enqueues on the current stream and never synchronises the host, so the whole step can be
captured once and replayed with a single launch.

    step = GraphedTrainStep(model, example_batch)      # example_batch: the 7 train_forward tensors (on device)
    loss = step(batch)                                 # copies the batch into the static inputs, replays
    # gradients are in p.grad of every parameter (static tensors, overwritten by the next replay)

Shapes, dtypes and the parameter set are frozen at capture time; the bf16 operand copies of the weights are
re-packed inside the graph, so optimizer updates of the fp32 master weights between replays are honoured.  Pass
`optimizer=FusedAdam(...)` to capture the parameter update in the same graph (warm-up steps update the weights too).
"""
from typing import Dict, Sequence

import torch

ORDER = ("user_id", "user_features", "user_history", "item_id", "item_features", "position", "labels")


class GraphedTrainStep:
    def __init__(self, model, example_batch: Dict[str, torch.Tensor], warmup: int = 3, post_backward=None,
                 optimizer=None):
        self.model = model
        self.post_backward = post_backward  # e.g. DataParallelContext.sync_gradients
        self.optimizer = optimizer  # e.g. FusedAdam: its step is captured after the backward (device-side step counter)
        dev = example_batch["user_id"].device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTrainStep needs CUDA tensors (two_tower_models_b200 has no CPU path)")
        # the captured inputs are views of ONE flat buffer: a batch that was packed the same way (pack() / new_slot())
        # is loaded with a single copy instead of seven
        self._layout, off = [], 0
        for k in ORDER:
            t = example_batch[k]
            n = t.numel() * t.element_size()
            self._layout.append((k, off, n, t.dtype, tuple(t.shape)))
            off += (n + 255) // 256 * 256
        self._flat_bytes = off
        self._flat = torch.empty(off, dtype=torch.uint8, device=dev)
        self.static = self._views(self._flat)
        for k in ORDER:
            self.static[k].copy_(example_batch[k])
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        for p in model.parameters():
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._eager()
        self.grads = [p.grad for p in model.parameters()]

    def _eager(self):
        m = self.model
        packed = getattr(m, "_packed", None)
        if packed is not None:
            packed.invalidate()  # weights are re-cast to bf16 inside every step
        enc = getattr(m, "user_history_encoder", None)
        if enc is not None:
            enc._packed.invalidate()
        for p in m.parameters():
            p.grad = None
        loss = m.train_forward(*[self.static[k] for k in ORDER])
        loss.backward()
        if self.post_backward is not None:
            self.post_backward(m)
        if self.optimizer is not None:
            self.optimizer.step()
        return loss.detach()

    def _views(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        return {k: flat[o:o + n].view(dt).view(shape) for (k, o, n, dt, shape) in self._layout}

    def new_slot(self, batch: Dict[str, torch.Tensor] = None, device=None, pin_memory: bool = False) -> Dict[str, torch.Tensor]:
        """A batch-shaped set of tensors backed by one flat buffer with the layout of the captured inputs (on `device`,
        default the model's; pin_memory=True for a host staging slot), optionally filled from `batch`.  load() moves
        such a slot with ONE copy."""
        device = self._flat.device if device is None else torch.device(device)
        flat = torch.empty(self._flat_bytes, dtype=torch.uint8, device=device, pin_memory=pin_memory and device.type == "cpu")
        slot = self._views(flat)
        slot["_flat"] = flat
        if batch is not None:
            for k in ORDER:
                slot[k].copy_(batch[k])
        return slot

    def load(self, batch: Dict[str, torch.Tensor], non_blocking: bool = True) -> None:
        """Copy a batch (device or pinned host tensors) into the static input buffers."""
        flat = batch.get("_flat")
        if flat is not None and flat.numel() == self._flat_bytes:
            self._flat.copy_(flat, non_blocking=non_blocking)
            return
        for k in ORDER:
            self.static[k].copy_(batch[k], non_blocking=non_blocking)

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.loss

    def __call__(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        self.load(batch)
        return self.replay()
