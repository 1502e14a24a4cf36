"""Drop-in for the reference's "N priors in one module" container:
``NumberBasedMultiPriorModule`` (``awesome/model/number_based_multi_prior_module.py:15-49``) on top of
``AbstractMultiPriorModule`` (``awesome/model/abstract_multi_prior_module.py:30-100``) -- ``priors.{k}.`` state-dict
keys, ``assure_prior_count``, ``load_state_dict`` resizing, ``forward(..., num_priors=O) -> [B,O,1,H,W]``
(``torch.stack(dim=1)`` of the per-prior outputs, exactly like ``_wrapped``).

The reference evaluates and fits the O priors one after another in Python.  Here their parameters are rows of
one ``[O,P]`` arena and a fit step of all O objects is ONE grouped launch per kernel (object = ``blockIdx.y``):
``make_fitter(grid, unaries[O,N])``.  The reference's multi-object pretrain is not runnable as written
(SURVEY a13); the semantics implemented are O independent fits, object k against unaries channel k."""
from __future__ import annotations

import copy
from typing import Any, List, Mapping, Optional

import torch
from torch import nn

from .. import _lib as L
from ..core import GridSpecHost, Prior
from .base import ArenaPriorModule


class NumberBasedMultiPriorModule(nn.Module):
    def __init__(self, prior: Optional[nn.Module] = None, prior_type=None, prior_args: Optional[dict] = None,
                 min_priors: int = 1):
        super().__init__()
        self.native_prior = prior
        self.prior_type = prior_type if prior_type is not None else type(prior)
        self.prior_args = prior_args if prior_args is not None else {}
        self.priors = nn.ModuleList()
        self._arena_all: Optional[torch.Tensor] = None
        self._group_prior: Optional[Prior] = None
        self.assure_prior_count(min_priors)

    # ---- container management (abstract_multi_prior_module.py:50-96)
    def create_prior(self) -> nn.Module:
        if self.native_prior is not None:
            p = copy.deepcopy(self.native_prior)
            if hasattr(p, "_flatten_"):
                p._flatten_()
        else:
            p = self.prior_type(**copy.deepcopy(self.prior_args))
        dev = next(self.parameters()).device if len(self.priors) else None
        return p.to(dev) if dev is not None else p

    def assure_prior_count(self, num: int) -> None:
        while len(self.priors) > num:
            del self.priors[len(self.priors) - 1]
            self._arena_all = None
        while len(self.priors) < num:
            self.priors.append(self.create_prior())
            self._arena_all = None

    def load_state_dict(self, state_dict: Mapping[str, Any], strict: bool = True):
        n = len({k.split(".")[1] for k in state_dict.keys() if k.startswith("priors.")})
        self.assure_prior_count(n)
        out = super().load_state_dict(state_dict, strict)
        self._arena_all = None
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._arena_all = None
        return out

    # ---- reference forward: stack on the channel dim
    def forward(self, *args, num_priors: Optional[int] = None, **kwargs) -> torch.Tensor:
        num_priors = len(self.priors) if num_priors is None else num_priors
        self.assure_prior_count(num_priors)
        return torch.stack([self.priors[i](*args, **kwargs) for i in range(num_priors)], dim=1)

    def enforce_convexity(self) -> None:
        for p in self.priors:
            if hasattr(p, "enforce_convexity"):
                p.enforce_convexity()

    def reset_parameters(self) -> None:
        for p in self.priors:
            p.reset_parameters()

    # ---- grouped native fit
    def _group_arena(self) -> torch.Tensor:
        """Re-point every prior's parameters into consecutive rows of one ``[O,P]`` tensor."""
        ps: List[ArenaPriorModule] = list(self.priors)
        if not ps or not all(isinstance(p, ArenaPriorModule) for p in ps):
            raise TypeError("grouped fits need awesome_b200 prior modules")
        rows = [p._ensure_flat() for p in ps]
        P = rows[0].numel()
        if any(r.numel() != P or r.device != rows[0].device for r in rows):
            raise ValueError("all priors of a grouped fit must have the same architecture and device")
        big = self._arena_all
        ok = big is not None and big.shape == (len(ps), P) and all(
            rows[k].data_ptr() == big.data_ptr() + 4 * k * P for k in range(len(ps)))
        if not ok:
            big = torch.empty((len(ps), P), dtype=torch.float32, device=rows[0].device)
            with torch.no_grad():
                for k, p in enumerate(ps):
                    big[k].copy_(rows[k])
                    off = 0
                    for q in p._arena_params():
                        n = q.numel()
                        q.data = big[k, off:off + n].view(q.shape)
                        off += n
                    p._arena = big[k]
            self._arena_all = big
            self._group_prior = None
        return big

    def make_fitter(self, grid, target: torch.Tensor, loss=None, optim=None, **kw):
        """``PriorFitter`` over all objects: ``target`` ``[O,N]`` (or ``[O,H,W]``), object k fitted to row k."""
        from ..fit import LossConfig, OptimConfig, PriorFitter
        big = self._group_arena()
        p0 = self.priors[0]
        if self._group_prior is None:
            with torch.cuda.device(big.device):
                single = p0._prior_for(big.device)
                d = single.desc
                self._group_prior = Prior(d.kind, d.C, d.h, d.L, d.F, d.m, bool(d.flow_tanh), len(self.priors), d.precision)
                if hasattr(p0, "_push_consts"):
                    p0._push_consts(self._group_prior)
        if isinstance(grid, torch.Tensor):
            grid = GridSpecHost.from_tensor(grid)
        if hasattr(p0, "_maybe_actnorm_init"):
            x = grid.materialize(p0.in_channels, big.device)
            for p in self.priors:
                p._maybe_actnorm_init(x)
        return PriorFitter(self._group_prior, big.reshape(-1), grid, target.reshape(len(self.priors), -1),
                           loss or LossConfig(), optim or OptimConfig(), **kw)
