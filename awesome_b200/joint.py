"""Data-parallel joint UNet + prior training step (BASELINE config 5): the body of
``TorchAgent._perform_step`` (``awesome/agent/torch_agent.py:428-551``) for the spatio-temporal configs, one process
per GPU.  Frames of the global batch are split across ranks; the model is replicated; the ONE exchange step of the
path is a single NCCL all-reduce between ``backward`` and ``optimizer.step`` (``torch_agent.py:491-492``).

* ``GradBucket``: every parameter's ``.grad`` is a view into one flat fp32 buffer (UNet 53.6 MB first, the prior's
  <= 175 KB of gradients in its tail), so autograd and the native prior backward accumulate straight into the
  message buffer -- no pack/unpack copies.  The buffer is cut into a few buckets whose all-reduce starts as soon as
  the bucket's gradients are complete, i.e. under the backward pass of the earlier layers.
* the optimizer is ``FusedAdam``: UNet parameters by torch, the prior arena in one native pass incl. the clamp."""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


class GradBucket:
    """Flat fp32 gradient buffer of a parameter list; ``.grad`` of every parameter is a view into it.

    ``n_buckets > 1`` cuts the buffer (at parameter boundaries) into contiguous buckets of about equal size.  With
    ``overlap`` hooks installed, a bucket's all-reduce is launched (asynchronously, on NCCL's stream) the moment autograd
    has accumulated the gradient of its last outstanding parameter, so the collective of the layers that finish their
    backward first runs under the backward of the rest; ``finish()`` waits for all of them.  NCCL's AVG reduction is used:
    every rank receives the same bits and no separate division pass is needed."""

    def __init__(self, params: Iterable[torch.nn.Parameter], n_buckets: int = 1):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        # bucket boundaries (element offsets) at parameter boundaries
        n_buckets = max(1, min(int(n_buckets), len(self.params)))
        target, self.bounds, self._bucket_of, acc, off, b = n / n_buckets, [0], {}, 0, 0, 0
        for i, p in enumerate(self.params):
            self._bucket_of[id(p)] = b
            off += p.numel()
            acc += p.numel()
            if acc >= target * (b + 1) - 1e-9 and b < n_buckets - 1 and i < len(self.params) - 1:
                self.bounds.append(off)
                b += 1
        self.bounds.append(n)
        self.n_buckets = len(self.bounds) - 1
        self._count = [0] * self.n_buckets
        for p in self.params:
            self._count[self._bucket_of[id(p)]] += 1
        self._pending = list(self._count)
        self._works: list = []
        self._hooks: list = []
        self._group = None
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
        self._pending = list(self._count)
        self._works = []

    def _active(self, group) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1

    # ---- overlapped: one asynchronous all-reduce per bucket, launched from autograd's post-accumulate hooks
    def install_overlap_hooks(self, group=None) -> None:
        self.remove_hooks()
        self._group = group

        def make(b):
            def hook(_p):
                self._pending[b] -= 1
                if self._pending[b] == 0 and self._active(self._group):
                    sl = self.flat[self.bounds[b]:self.bounds[b + 1]]
                    self._works.append(dist.all_reduce(sl, op=dist.ReduceOp.AVG, group=self._group, async_op=True))
            return hook
        for p in self.params:
            self._hooks.append(p.register_post_accumulate_grad_hook(make(self._bucket_of[id(p)])))

    def remove_hooks(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []

    def finish(self) -> None:
        """Wait for the bucket collectives launched during backward; buckets that never completed (a parameter without
        gradient this step) are reduced now."""
        if self._active(self._group):
            for b in range(self.n_buckets):
                if self._pending[b] > 0:
                    sl = self.flat[self.bounds[b]:self.bounds[b + 1]]
                    self._works.append(dist.all_reduce(sl, op=dist.ReduceOp.AVG, group=self._group, async_op=True))
        for w in self._works:
            w.wait()
        self._works = []

    # ---- not overlapped: one collective after backward
    def allreduce_mean(self, group=None) -> None:
        if self._active(group):
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)

    @property
    def nbytes(self) -> int:
        return 4 * self.flat.numel()


class JointTrainer:
    """``step(images, grids, labels)``: ``out = cat([sigmoid(seg(images)), sigmoid(prior(grids))], 1)``;
    ``loss(out, labels).backward()``; gradient all-reduce (mean over ranks); ``optimizer.step()`` (+ clamp).

    A non-finite loss raises ``StopTraining``-style ``ValueError`` before the collective, like the agent's NaN guard
    (``torch_agent.py:484-487``) -- decided collectively so that no rank is left waiting in the all-reduce."""

    def __init__(self, seg_net: torch.nn.Module, prior: torch.nn.Module, loss: Callable, optimizer_cls=None,
                 optimizer_args: Optional[dict] = None, group=None, n_buckets: int = 4, comm: str = "after"):
        """``comm``: "after" (one all-reduce of the flat bucket after backward, default), "overlap" (bucketed all-reduce
        launched from autograd hooks under the backward pass) or "off" (no exchange: single-replica timing baseline).
        Measured on 2 x B200 (bench.py joint leg): the 53.8 MB collective takes 0.13 ms of a 27 ms step; the overlapped
        variant is 1.2 ms SLOWER (60 Python hooks per step, NCCL kernels competing with cuDNN for SMs), hence the default."""
        from .optim import FusedAdam
        if comm not in ("overlap", "after", "off"):
            raise ValueError("comm must be 'overlap', 'after' or 'off'")
        self.seg_net, self.prior, self.loss, self.group, self.comm = seg_net, prior, loss, group, comm
        params = list(seg_net.parameters()) + list(prior.parameters())
        cls = optimizer_cls or FusedAdam
        self.optimizer = cls([dict(params=list(seg_net.parameters())), dict(params=list(prior.parameters()))],
                             **(optimizer_args or dict(lr=1e-4)))
        self.bucket = GradBucket(params, n_buckets=n_buckets if comm == "overlap" else 1)
        if comm == "overlap":
            self.bucket.install_overlap_hooks(group)
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
        if self.comm == "overlap":
            self.bucket.finish()
        elif self.comm == "after":
            self.bucket.allreduce_mean(self.group)
        self.optimizer.step()
        if hasattr(self.prior, "enforce_convexity"):
            self.prior.enforce_convexity()      # batch_processed hook of the runner (awesome_runner.py:294-297); idempotent
        self.steps += 1
        return loss.detach()
