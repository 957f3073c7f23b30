"""CUDA-graph replay of the forward + backward of a training step.

A training step of the decoder launches ~750 kernels, ~330 of them launch-bound (sub-10 us node-level and scalar
ops); stream-ordered they leave ~1.7 ms of gaps in a 58 ms step.  :class:`GraphedStep` captures ``step_fn`` -- decoder
forward, losses and ``backward()`` -- once into a CUDA graph and replays it with the next batch copied into fixed input
buffers.  The optimizer step (and, data parallel, the gradient all-reduce) stay outside and run eagerly on the
gradients the graph writes.

Everything whose result depends on HOST state at capture time is frozen into the graph: the packed sizes (valid lengths
of ``mask``), the band graph and the loss weights.  Replays are therefore only valid for batches with the same shapes and
the same mask layout as ``example``; :meth:`GraphedStep.matches` checks a host batch against it.
"""
from __future__ import annotations

import torch


class GraphedStep:
    """``loss = graphed(batch)``: ``step_fn(batch) -> loss`` (forward + ``loss.backward()``, no optimizer, no host
    synchronisation) captured once on ``example`` (dict of device tensors) and replayed.

    ``params``: the parameters whose ``.grad`` the step produces; after a call they hold this step's gradients (the same
    tensors every time -- do not ``zero_grad(set_to_none=True)`` between calls, or call :meth:`attach_grads`).
    """

    def __init__(self, step_fn, example: dict, params, warmup: int = 2):
        self.params = [p for p in params if p.requires_grad]
        self.inputs = {k: v.clone() for k, v in example.items()}
        self._mask_host = example["mask"].detach().cpu() if "mask" in example else None
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):                       # warm-up off the default stream: fills every host-side cache
            for _ in range(max(1, warmup)):
                for p in self.params:
                    p.grad = None
                step_fn(self.inputs)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        for p in self.params:
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = step_fn(self.inputs)
        self.grads = [p.grad for p in self.params]          # allocated inside the graph's pool: rewritten by every replay

    def matches(self, host_batch: dict) -> bool:
        """Same shapes / dtypes and the same mask as the captured batch (host-side check, no device sync)."""
        for k, v in self.inputs.items():
            h = host_batch.get(k)
            if h is None or h.shape != v.shape or h.dtype != v.dtype:
                return False
        if self._mask_host is not None and not host_batch["mask"].is_cuda:
            return bool(torch.equal(host_batch["mask"], self._mask_host))
        return True

    def attach_grads(self):
        for p, g in zip(self.params, self.grads):
            p.grad = g

    def __call__(self, batch: dict) -> torch.Tensor:
        for k, v in self.inputs.items():
            src = batch[k]
            if src.data_ptr() != v.data_ptr():
                v.copy_(src, non_blocking=True)
        self.attach_grads()
        self.graph.replay()
        return self.loss
