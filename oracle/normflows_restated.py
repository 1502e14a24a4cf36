"""Restated subset of the third-party package ``normflows==1.7.3``.

TEST INFRASTRUCTURE ONLY.  Nothing under ``awesome_b200/`` may import this file.

The reference builds its RealNVP through ``normflows`` (reference
``awesome/model/net_factory.py:70-114``; pin ``poetry.lock:2407-2413``).  The
package is not vendored in ``/root/reference`` and cannot be installed here (no
network), so its published arithmetic is restated below, small enough to audit by
eye.  ``oracle/ref_shim.py`` registers this module as ``sys.modules["normflows"]``
so that the reference's own ``init_realnvp`` / ``real_nvp_path_connected_net`` run
unmodified on top of it.

PARITY UNPINNED at this boundary: the reference ships no test or golden vector
for the flow, and the wheel is not available offline.  The restatement follows the
package's public semantics:

* ``nets.MLP(layers, init_zeros, output_fn, output_scale)``: ``Linear ->
  LeakyReLU(leaky=0.0) -> ... -> Linear`` (+ ``Tanh``/``Sigmoid``/``ReLU`` when
  ``output_fn`` is given, then an optional constant scale); the last Linear is
  zero-initialised when ``init_zeros``.  Sequential indices give the state-dict
  keys ``net.0.*`` / ``net.2.*``.
* ``flows.MaskedAffineFlow(b, t, s)``: ``zm = b*z``; ``z' = zm + (1-b)*(z*exp(s(zm))
  + t(zm))`` with non-finite s/t replaced by NaN; inverse ``zm + (1-b)*(z-t)*exp(-s)``.
* ``flows.ActNorm(C)``: ``z*exp(s)+t`` with ``s,t`` of shape ``[1,C]``; the first
  forward call sets ``s=-log(std(z, dim=0)+1e-6)`` (unbiased std) and
  ``t=-mean*exp(s)`` and flips the buffer ``data_dep_init_done``.
* ``NormalizingFlow(q0, flows, p)``: ``forward`` applies the flows in order and
  discards the log-determinants; ``inverse`` walks them in reverse.
"""
from __future__ import annotations

import types

import numpy as np
import torch
from torch import nn


class _ConstScaleLayer(nn.Module):
    def __init__(self, scale=1.0):
        super().__init__()
        self.scale_cpu = torch.tensor(scale)
        self.register_buffer("scale", self.scale_cpu)

    def forward(self, x):
        return x * self.scale


class MLP(nn.Module):
    def __init__(self, layers, leaky=0.0, score_scale=None, output_fn=None,
                 output_scale=None, init_zeros=False, dropout=None):
        super().__init__()
        net = nn.ModuleList([])
        for k in range(len(layers) - 2):
            net.append(nn.Linear(layers[k], layers[k + 1]))
            net.append(nn.LeakyReLU(leaky))
        if dropout is not None:
            net.append(nn.Dropout(p=dropout))
        net.append(nn.Linear(layers[-2], layers[-1]))
        if init_zeros:
            nn.init.zeros_(net[-1].weight)
            nn.init.zeros_(net[-1].bias)
        if output_fn is not None:
            if score_scale is not None:
                net.append(_ConstScaleLayer(score_scale))
            if output_fn == "sigmoid":
                net.append(nn.Sigmoid())
            elif output_fn == "relu":
                net.append(nn.ReLU())
            elif output_fn == "tanh":
                net.append(nn.Tanh())
            else:
                raise NotImplementedError("output function not restated: %s" % output_fn)
            if output_scale is not None:
                net.append(_ConstScaleLayer(output_scale))
        self.net = nn.Sequential(*net)

    def forward(self, x):
        return self.net(x)


class Flow(nn.Module):
    def forward(self, z):
        raise NotImplementedError

    def inverse(self, z):
        raise NotImplementedError


class MaskedAffineFlow(Flow):
    def __init__(self, b, t=None, s=None):
        super().__init__()
        self.b_cpu = b.view(1, *b.size())
        self.register_buffer("b", self.b_cpu)
        if s is None:
            self.s = torch.zeros_like
        else:
            self.add_module("s", s)
        if t is None:
            self.t = torch.zeros_like
        else:
            self.add_module("t", t)

    def forward(self, z):
        z_masked = self.b * z
        scale = self.s(z_masked)
        nan = torch.tensor(np.nan, dtype=z.dtype, device=z.device)
        scale = torch.where(torch.isfinite(scale), scale, nan)
        trans = self.t(z_masked)
        trans = torch.where(torch.isfinite(trans), trans, nan)
        z_ = z_masked + (1 - self.b) * (z * torch.exp(scale) + trans)
        log_det = torch.sum((1 - self.b) * scale, dim=list(range(1, self.b.dim())))
        return z_, log_det

    def inverse(self, z):
        z_masked = self.b * z
        scale = self.s(z_masked)
        nan = torch.tensor(np.nan, dtype=z.dtype, device=z.device)
        scale = torch.where(torch.isfinite(scale), scale, nan)
        trans = self.t(z_masked)
        trans = torch.where(torch.isfinite(trans), trans, nan)
        z_ = z_masked + (1 - self.b) * (z - trans) * torch.exp(-scale)
        log_det = -torch.sum((1 - self.b) * scale, dim=list(range(1, self.b.dim())))
        return z_, log_det


class AffineConstFlow(Flow):
    def __init__(self, shape, scale=True, shift=True):
        super().__init__()
        if isinstance(shape, int):
            shape = (shape,)
        if scale:
            self.s = nn.Parameter(torch.zeros(shape)[None])
        else:
            self.register_buffer("s", torch.zeros(shape)[None])
        if shift:
            self.t = nn.Parameter(torch.zeros(shape)[None])
        else:
            self.register_buffer("t", torch.zeros(shape)[None])
        self.n_dim = self.s.dim()
        self.batch_dims = torch.nonzero(
            torch.tensor(self.s.shape) == 1, as_tuple=False)[:, 0].tolist()

    def forward(self, z):
        z_ = z * torch.exp(self.s) + self.t
        if len(self.batch_dims) > 1:
            prod_batch_dims = np.prod([z.size(i) for i in self.batch_dims[1:]])
        else:
            prod_batch_dims = 1
        log_det = prod_batch_dims * torch.sum(self.s)
        return z_, log_det

    def inverse(self, z):
        z_ = (z - self.t) * torch.exp(-self.s)
        if len(self.batch_dims) > 1:
            prod_batch_dims = np.prod([z.size(i) for i in self.batch_dims[1:]])
        else:
            prod_batch_dims = 1
        log_det = -prod_batch_dims * torch.sum(self.s)
        return z_, log_det


class ActNorm(AffineConstFlow):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.data_dep_init_done_cpu = torch.tensor(0.0)
        self.register_buffer("data_dep_init_done", self.data_dep_init_done_cpu)

    def forward(self, z):
        if not self.data_dep_init_done > 0.0:
            assert self.s is not None and self.t is not None
            s_init = -torch.log(z.std(dim=self.batch_dims, keepdim=True) + 1e-6)
            self.s.data = s_init.data
            self.t.data = (-z.mean(dim=self.batch_dims, keepdim=True)
                           * torch.exp(self.s)).data
            self.data_dep_init_done = torch.tensor(1.0)
        return super().forward(z)

    def inverse(self, z):
        if not self.data_dep_init_done:
            assert self.s is not None and self.t is not None
            s_init = torch.log(z.std(dim=self.batch_dims, keepdim=True) + 1e-6)
            self.s.data = s_init.data
            self.t.data = z.mean(dim=self.batch_dims, keepdim=True).data
            self.data_dep_init_done = torch.tensor(1.0)
        return super().inverse(z)


class _BaseDistribution(nn.Module):
    pass


class Uniform(_BaseDistribution):
    """Placeholder base distribution: never sampled on the prior-fit path."""

    def __init__(self, shape, low=-1.0, high=1.0):
        super().__init__()
        if isinstance(shape, int):
            shape = (shape,)
        if isinstance(shape, list):
            shape = tuple(shape)
        self.shape = shape
        self.d = np.prod(shape)
        self.low = torch.tensor(low)
        self.high = torch.tensor(high)


class NormalizingFlow(nn.Module):
    def __init__(self, q0, flows, p=None):
        super().__init__()
        self.q0 = q0
        self.flows = nn.ModuleList(flows)
        self.p = p

    def forward(self, z):
        for flow in self.flows:
            z, _ = flow(z)
        return z

    def inverse(self, x):
        for i in range(len(self.flows) - 1, -1, -1):
            x, _ = self.flows[i].inverse(x)
        return x


def as_module() -> types.ModuleType:
    """Assemble a module object shaped like the ``normflows`` package."""
    nf = types.ModuleType("normflows")
    nf.NormalizingFlow = NormalizingFlow
    nets = types.ModuleType("normflows.nets")
    nets.MLP = MLP
    flows = types.ModuleType("normflows.flows")
    flows.MaskedAffineFlow = MaskedAffineFlow
    flows.ActNorm = ActNorm
    flows.AffineConstFlow = AffineConstFlow
    flows.Flow = Flow
    dist = types.ModuleType("normflows.distributions")
    base = types.ModuleType("normflows.distributions.base")
    base.Uniform = Uniform
    dist.base = base
    nf.nets, nf.flows, nf.distributions = nets, flows, dist
    nf.__restated__ = True
    return nf
