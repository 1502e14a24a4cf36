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


class BatchSizeMultiPriorModule(NumberBasedMultiPriorModule):
    """``awesome/model/batch_size_multi_prior_module.py:13-18``: one prior per batch item; ``forward(..., batch_size=B)``
    evaluates ``B`` priors on the same arguments and stacks them on dim 1."""

    def forward(self, *args, batch_size: int, **kwargs) -> torch.Tensor:       # noqa: D401
        return super().forward(*args, num_priors=batch_size, **kwargs)


class MultipleObjectsAwarePathConnectedNet(NumberBasedMultiPriorModule):
    """``awesome/model/multiple_object_aware_path_connected_net.py``: one path-connectedness prior per object of a frame.

    The reference's loop (``:186-367``) is experimental and not runnable as written; the semantics kept are
    the ones its structure spells out: ``n_priors = unaries.shape[1] - 1`` (channel 0 is the background, ``:186-192``),
    object ``k`` is fitted against unaries channel ``k + 1``, every object carries its OWN warm-start chain from frame to
    frame (``previous_image_object_states``, ``:206-219``), fresh optimizer per object and attempt (``:289-298``), IoU check
    and retry per object (``:330-355``), checkpoints ``pretrain_checkpoint_{i}_{k}.pth`` holding the OBJECT's prior
    (``:220-229, 364-367``).  The O objects of a frame are independent, so a frame is ONE grouped launch per kernel whenever
    its objects need the same number of steps (all cold or all warm); otherwise object by object."""

    def pretrain_load_state(self, train_set, test_set, device, agent, state, use_progress_bar: bool = True,
                            wrapper_module=None, **kwargs):
        agent.training_dataset.__prior_cache__.set_state(state)

    def pretrain(self, train_set, test_set=None, device=None, agent=None, use_progress_bar: bool = True,
                 do_pretrain_checkpoints: bool = False, use_pretrain_checkpoints: bool = False,
                 pretrain_checkpoint_dir: Optional[str] = None, wrapper_module=None, **kwargs) -> Any:
        from .. import pretrain as P
        ds = getattr(agent, "training_dataset", None)
        cache = getattr(ds, "__prior_cache__", None)
        if cache is None or not bool(getattr(ds, "has_prior", getattr(ds, "__has_prior__", False))):
            raise NotImplementedError("Fixed pretraining not implemented.")
        if wrapper_module is None:
            raise ValueError("Wrapper model must be provided for pretraining.")
        sched = P.FitSchedule.from_pretrain_args(kwargs)
        dev = torch.device(device) if device is not None else next(self.parameters()).device
        was = wrapper_module.training
        wrapper_module.eval()
        try:
            _, grids, uns, keys = P.collect_unaries(wrapper_module, agent, train_set, dev, sched.unet_batch_size)

            def keep(i, _results):
                state = {k: v.detach().clone() for k, v in self.state_dict().items()}
                if hasattr(cache, "__setitem__"):
                    cache[keys[i]] = state
            self.fit_frames_multi_object(grids, uns, sched, on_frame=keep, do_pretrain_checkpoints=do_pretrain_checkpoints,
                                         use_pretrain_checkpoints=use_pretrain_checkpoints,
                                         pretrain_checkpoint_dir=pretrain_checkpoint_dir)
            return cache.get_state()
        finally:
            wrapper_module.train(was)

    def fit_frames_multi_object(self, grids, unaries, schedule=None, on_frame=None, do_pretrain_checkpoints: bool = False,
                                use_pretrain_checkpoints: bool = False, pretrain_checkpoint_dir: Optional[str] = None):
        """``unaries[i]``: ``[1, O+1, H, W]`` (or ``[1,1,H,W]``: one object).  Returns ``results[i][k]`` (``FrameResult`` of
        object k in frame i, or ``None`` for a skipped frame)."""
        import dataclasses
        import logging
        import os
        from .. import pretrain as P
        from ..core import iou_counts, target_counts
        s = schedule or P.FitSchedule()
        if do_pretrain_checkpoints:
            if pretrain_checkpoint_dir is None:
                raise ValueError("Pretrain checkpoint dir must be provided.")
            os.makedirs(pretrain_checkpoint_dir, exist_ok=True)
        all_results = []
        prev: dict = {}                     # object -> state (device row) of the previous frame, when its fit was proper
        prev_frame = -2
        for i, (grid, un) in enumerate(zip(grids, unaries)):
            dev = next(self.parameters()).device if len(self.priors) else torch.device("cuda")
            un = un.detach().float()
            if un.dim() == 3:
                un = un.unsqueeze(0)
            O_ = un.shape[1] - 1 if un.shape[1] > 1 else 1
            self.assure_prior_count(O_)
            dev = next(self.parameters()).device
            un = un.to(dev)
            obj_un = un[0, 1:] if un.shape[1] > 1 else un[0]                 # [O,H,W]: channel 0 is the background
            cnt = target_counts(un.reshape(1, -1), L.AWB_CLS_UNARY_LT_HALF).cpu()[0]
            if int(cnt[0]) == 0 or int(cnt[1]) == 0:                          # torch.unique(unaries >= 0.5) has one value
                logging.warning("Unaries of segmentation model contain no foreground. Skipping image. %s", i)
                all_results.append(None)
                continue
            if prev_frame != i - 1:
                prev = {}                                                      # only the last image's states are kept (:209-212)
            spec = P._as_grid(grid, dev)
            x = spec.materialize(self.priors[0].in_channels, dev)
            results: List[Any] = [None] * O_
            todo = list(range(O_))
            if use_pretrain_checkpoints and pretrain_checkpoint_dir:
                for k in list(todo):
                    path = os.path.join(pretrain_checkpoint_dir, f"pretrain_checkpoint_{i}_{k}.pth")
                    if os.path.exists(path) and P.load_pretrain_checkpoint(self.priors[k], path, device=dev):
                        results[k] = P.FrameResult(index=i, proper_fit=True, state=self.priors[k]._ensure_flat().detach().clone())
                        todo.remove(k)
            warm = {k: (s.reuse_state and k in prev) for k in todo}
            for k in todo:
                pk = self.priors[k]
                if warm[k]:
                    with torch.no_grad():
                        pk._ensure_flat().copy_(prev[k])
                else:
                    if s.prefit_flow_net_identity:
                        pk.learn_flow_identity(x, lr=s.prefit_flow_net_identity_lr, weight_decay=s.prefit_flow_net_identity_weight_decay,
                                               max_iter=s.prefit_flow_net_identity_num_epochs, use_progress_bar=False)
                    if s.prefit_convex_net:
                        pk.learn_convex_net(x, obj_un[k][None, None], lr=s.prefit_convex_net_lr,
                                            weight_decay=s.prefit_convex_net_weight_decay,
                                            max_iter=s.prefit_convex_net_num_epochs, use_progress_bar=False)
            grouped = len(todo) == O_ and O_ > 1 and len({warm[k] for k in todo}) == 1
            if grouped:
                epochs = s.reuse_state_epochs if warm[0] else s.num_epochs
                fitter = self.make_fitter(spec, obj_un.reshape(O_, -1), s.criterion, s.optim(True), steps_per_graph=s.steps_per_graph)
                hist = fitter.run(epochs)
                fitter.raise_if_nonfinite()
                with torch.no_grad():
                    logits = self(x, num_priors=O_)
                cnts = iou_counts(logits.reshape(O_, -1), obj_un.reshape(O_, -1), pred_is_logit=True, n_objects=O_).cpu()
                for k in range(O_):
                    inter, pf, tf = int(cnts[k, 0]), int(cnts[k, 1]), int(cnts[k, 2])
                    r = P.FrameResult(index=i, steps=epochs, final_loss=float(hist[-1, k]) if epochs > 0 else float("nan"))
                    r.iou = 0.0 if tf == 0 else inter / float(pf + tf - inter)
                    r.proper_fit = r.iou >= s.proper_prior_fit_threshold
                    results[k] = r
                redo = [k for k in range(O_) if not results[k].proper_fit and s.proper_prior_fit_retrys > 0]
            else:
                redo = list(todo)
            for k in redo:                     # per-object path: mixed cold / warm frames, and the reset + refit retry
                pk = self.priors[k]
                first_try = results[k] is None
                if not first_try:
                    logging.info("Prior fit not proper on image index: %s object %s. Retrying. Metric: %s", i, k, results[k].iou)
                    pk.reset_parameters()
                    pk._ensure_flat()
                sch_k = dataclasses.replace(s, prefit_flow_net_identity=False, prefit_convex_net=False, reuse_state=True,
                                            proper_prior_fit_retrys=s.proper_prior_fit_retrys - (0 if first_try else 1))
                init_prev = pk._ensure_flat().detach().clone() if (first_try and warm.get(k, False)) else None
                again = P.fit_frames(pk, [spec], [obj_un[k]], sch_k, frame_indices=[i], initial_previous=init_prev)[0]
                if not first_try:
                    again.retries += 1
                    again.steps += results[k].steps
                results[k] = again
                self._arena_all = None
            for k in range(O_):
                results[k].state = self.priors[k]._ensure_flat().detach().clone()
                if s.reuse_state and results[k].proper_fit:
                    prev[k] = results[k].state
                elif k in prev:
                    del prev[k]
                if do_pretrain_checkpoints and k in todo:
                    P.save_pretrain_checkpoint(self.priors[k], os.path.join(pretrain_checkpoint_dir, f"pretrain_checkpoint_{i}_{k}.pth"))
            prev_frame = i
            all_results.append(results)
            if on_frame:
                on_frame(i, results)
        return all_results
