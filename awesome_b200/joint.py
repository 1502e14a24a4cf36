"""Data-parallel joint UNet + prior training step (BASELINE config 5): the body of
``TorchAgent._perform_step`` (``awesome/agent/torch_agent.py:428-551``) for the spatio-temporal configs, one process
per GPU.  Frames of the global batch are split across ranks; the model is replicated; the ONE exchange step of the
path is a single NCCL all-reduce between ``backward`` and ``optimizer.step`` (``torch_agent.py:491-492``).

* ``GradBucket``: every parameter's ``.grad`` is a view into one flat fp32 buffer (UNet 53.6 MB first, the prior's
  <= 175 KB of gradients in its tail), so autograd and the native prior backward accumulate straight into the
  message buffer -- no pack/unpack copies, one collective per step.
* the optimizer is ``FusedAdam``: UNet parameters by torch, the prior arena in one native pass incl. the clamp."""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


class GradBucket:
    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.attach()

    def attach(self) -> None:
        """(Re)point every ``.grad`` into the flat buffer (needed again after ``zero_grad(set_to_none=True)``)."""
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view(p.shape)
            off += n

    def zero(self) -> None:
        self.flat.zero_()
        if any(p.grad is None or p.grad.data_ptr() < self.flat.data_ptr()
               or p.grad.data_ptr() >= self.flat.data_ptr() + 4 * self.flat.numel() for p in self.params):
            self.attach()

    def allreduce_mean(self, group=None) -> None:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))

    @property
    def nbytes(self) -> int:
        return 4 * self.flat.numel()


class JointTrainer:
    """``step(images, grids, labels)``: ``out = cat([sigmoid(seg(images)), sigmoid(prior(grids))], 1)``;
    ``loss(out, labels).backward()``; gradient all-reduce (mean over ranks); ``optimizer.step()`` (+ clamp).

    A non-finite loss raises ``StopTraining``-style ``ValueError`` before the collective, like the agent's NaN guard
    (``torch_agent.py:484-487``) -- decided collectively so that no rank is left waiting in the all-reduce."""

    def __init__(self, seg_net: torch.nn.Module, prior: torch.nn.Module, loss: Callable, optimizer_cls=None,
                 optimizer_args: Optional[dict] = None, group=None):
        from .optim import FusedAdam
        self.seg_net, self.prior, self.loss, self.group = seg_net, prior, loss, group
        params = list(seg_net.parameters()) + list(prior.parameters())
        cls = optimizer_cls or FusedAdam
        self.optimizer = cls([dict(params=list(seg_net.parameters())), dict(params=list(prior.parameters()))],
                             **(optimizer_args or dict(lr=1e-4)))
        self.bucket = GradBucket(params)
        self.steps = 0

    def broadcast_parameters(self, src: int = 0) -> None:
        """Replicas start identical (the reference seeds every process the same: ``awesome/run/runner.py:19-25``)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            for t in list(self.seg_net.parameters()) + list(self.seg_net.buffers()) + list(self.prior.parameters()):
                dist.broadcast(t.data, src=src, group=self.group)

    def step(self, images: torch.Tensor, grids: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        self.bucket.zero()
        seg = torch.sigmoid(self.seg_net(images))
        pri = torch.sigmoid(self.prior(grids))
        out = torch.cat([seg, pri], dim=1)
        loss = self.loss(out, labels)
        bad = (~torch.isfinite(loss.detach())).float().reshape(1)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=self.group)
        if float(bad) > 0:
            raise ValueError("Loss is nan or inf!")
        loss.backward()
        self.bucket.allreduce_mean(self.group)
        self.optimizer.step()
        if hasattr(self.prior, "enforce_convexity"):
            self.prior.enforce_convexity()      # batch_processed hook of the runner (awesome_runner.py:294-297); idempotent
        self.steps += 1
        return loss.detach()
