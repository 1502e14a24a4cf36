"""Device-resident prior store: the reference swaps per-frame weights with
``load_state_dict`` + ``deepcopy(state_dict())`` + a CPU round trip on every step
(``awesome/util/prior_cache.py:10-91``, ``awesome/dataset/prior_dataset.py:70-110``).  Here every frame's
parameters are one row of a ``[capacity, P]`` fp32 arena on the model's device; entering a frame is one
device-to-device row copy into the module's flat arena, leaving it is the copy back.  ``get_state`` /
``set_state`` / ``save`` / ``load`` keep the reference's on-disk format (``{model_type, model_args,
store_device, cache{key -> state_dict}}``), so files are interchangeable."""
from __future__ import annotations

import copy
import json
import math
from typing import Any, Dict, Optional

import torch


def _class_name(t) -> Optional[str]:
    if t is None:
        return None
    return f"{t.__module__}.{getattr(t, '__qualname__', getattr(t, '__name__', str(t)))}"


def _import(name: str):
    import importlib
    mod, _, attr = name.rpartition(".")
    return getattr(importlib.import_module(mod), attr)


class DevicePriorCache:
    """Same public surface as the reference ``PriorCache`` (``__contains__``, ``__getitem__``, ``__setitem__``,
    ``generate_prior``, ``extract_prior``, ``apply_prior``, ``get_state``, ``set_state``, ``save``, ``load``) plus
    the row-copy fast path ``load_into`` / ``store_from`` used by ``PriorManager``."""

    def __init__(self, model_type=None, model_args: Optional[Dict[str, Any]] = None,
                 store_device: Optional[torch.device] = None, capacity: int = 16):
        self.model_type = model_type
        self.model_args = copy.deepcopy(model_args) if model_args is not None else {}
        self.store_device = torch.device(store_device) if store_device is not None else None
        self._rows: Optional[torch.Tensor] = None      # [capacity, P]
        self._index: Dict[int, int] = {}               # key -> row
        self._extra: Dict[int, Dict[str, torch.Tensor]] = {}   # key -> non-parameter state entries (buffers)
        self._layout = None                            # [(state key, shape, offset)] of the parameters, arena order
        self._capacity0 = max(1, int(capacity))
        self._generic: Dict[int, Any] = {}             # states of models without a flat arena

    # ---- layout helpers
    def _ensure_layout(self, model) -> bool:
        if self._layout is not None:
            return True
        if not hasattr(model, "_arena_params"):
            return False
        names = {id(p): k for k, p in model.named_parameters()}
        layout, off = [], 0
        for p in model._arena_params():
            layout.append((names[id(p)], tuple(p.shape), off))
            off += p.numel()
        self._layout = layout
        self._P = off
        return True

    def _row(self, key: int, device) -> torch.Tensor:
        if self._rows is None:
            self._rows = torch.empty((self._capacity0, self._P), dtype=torch.float32, device=device)
        if key not in self._index:
            r = len(self._index)
            if r >= self._rows.shape[0]:
                grown = torch.empty((2 * self._rows.shape[0], self._P), dtype=torch.float32, device=self._rows.device)
                grown[:self._rows.shape[0]].copy_(self._rows)
                self._rows = grown
            self._index[key] = r
        return self._rows[self._index[key]]

    # ---- reference surface
    def __contains__(self, key: int) -> bool:
        return key in self._index or key in self._generic

    def __len__(self) -> int:
        return len(self._index) + len(self._generic)

    def generate_prior(self, key: int) -> Any:
        """A fresh model's state dict, like the reference (``prior_cache.py:29-32``)."""
        return self.model_type(**copy.deepcopy(self.model_args)).state_dict()

    @staticmethod
    def extract_prior(model: torch.nn.Module) -> Any:
        if hasattr(model, "extract_prior") and callable(model.extract_prior):
            return model.extract_prior()
        return copy.deepcopy(model.state_dict())

    @staticmethod
    def apply_prior(model: torch.nn.Module, prior: Any) -> None:
        if hasattr(model, "apply_prior") and callable(model.apply_prior):
            return model.apply_prior(prior)
        model.load_state_dict(prior)

    def __getitem__(self, key: int) -> Dict[str, torch.Tensor]:
        if key in self._generic:
            return self._generic[key]
        if key not in self._index:
            self[key] = self.generate_prior(key)
            if key in self._generic:
                return self._generic[key]
        row = self._rows[self._index[key]]
        out = {k: row[off:off + math.prod(shape)].view(shape) for k, shape, off in self._layout}
        out.update(self._extra.get(key, {}))
        return out

    def __setitem__(self, key: int, value: Dict[str, torch.Tensor]) -> None:
        if self._layout is None:
            # learn the parameter layout from a model instance when the type is known, else store generically
            proto = None
            if self.model_type is not None:
                try:
                    proto = self.model_type(**copy.deepcopy(self.model_args))
                except Exception:
                    proto = None
            if proto is None or not self._ensure_layout(proto):
                self._generic[key] = {k: (v.detach().clone().to(self.store_device) if self.store_device is not None
                                          else v.detach().clone()) for k, v in value.items()}
                return
        pkeys = {k for k, _, _ in self._layout}
        dev = self.store_device or next(iter(value.values())).device
        row = self._row(key, dev)
        for k, shape, off in self._layout:
            row[off:off + value[k].numel()].copy_(value[k].detach().reshape(-1))
        self._extra[key] = {k: v.detach().clone().to(row.device) for k, v in value.items() if k not in pkeys}

    # ---- fast path used by PriorManager
    def load_into(self, model, key: int) -> None:
        """Frame ``key`` -> the module's arena: one device-to-device copy (mints a fresh state on first use)."""
        if not self._ensure_layout(model):
            return self.apply_prior(model, self[key])
        arena = model._ensure_flat()
        if key not in self._index:
            self[key] = self.generate_prior(key)
        with torch.no_grad():
            arena.copy_(self._rows[self._index[key]].to(arena.device, non_blocking=True))
            extra = self._extra.get(key)
            if extra:
                sd = model.state_dict()
                for k, v in extra.items():
                    if k in sd and sd[k].shape == v.shape:
                        sd[k].copy_(v)

    def store_from(self, model, key: int) -> None:
        if not self._ensure_layout(model):
            self[key] = self.extract_prior(model)
            return
        arena = model._ensure_flat()
        with torch.no_grad():
            self._row(key, self.store_device or arena.device).copy_(arena)
            pkeys = {k for k, _, _ in self._layout}
            self._extra[key] = {k: v.detach().clone() for k, v in model.state_dict().items() if k not in pkeys}

    # ---- on-disk format of the reference (prior_cache.py:61-91)
    def get_state(self) -> Dict[str, Any]:
        cache = {}
        for key in list(self._index) + list(self._generic):
            cache[str(key)] = {k: v.detach().to("cpu").clone() for k, v in self[key].items()}
        return {"model_type": _class_name(self.model_type), "model_args": json.dumps(self.model_args, default=str),
                "store_device": str(self.store_device), "cache": cache}

    def set_state(self, state: Dict[str, Any]) -> None:
        if state.get("model_type"):
            try:
                self.model_type = _import(state["model_type"])
            except Exception:
                pass
        ma = state.get("model_args")
        if isinstance(ma, str):
            try:
                self.model_args = json.loads(ma)
            except Exception:
                pass
        dev = state.get("store_device")
        self.store_device = None if dev in (None, "None") else torch.device(dev)
        self._rows, self._index, self._extra, self._generic = None, {}, {}, {}
        for k, v in state["cache"].items():
            self[int(k)] = v

    def save(self, f) -> None:
        torch.save(self.get_state(), f)

    @classmethod
    def load(cls, f) -> "DevicePriorCache":
        res = torch.load(f, map_location="cpu", weights_only=False)
        c = cls(None, None)
        c.set_state(res)
        return c


class PriorManager:
    """Context manager with the reference's signature (``prior_dataset.py:70-110``): applies the frame's state on
    enter and stores the (possibly trained) state on exit.  With a ``DevicePriorCache`` both are row copies."""

    def __init__(self, model: torch.nn.Module, prior_state=None, prior_cache=None, model_device=None,
                 store_device=None, training: bool = False):
        self.model, self.state, self.training = model, prior_state, training
        if prior_cache is not None and hasattr(prior_cache, "__prior_cache__"):
            prior_cache = prior_cache.__prior_cache__
        self.prior_cache = prior_cache
        self.model_device = model_device

    def _prior_module(self):
        m = self.model
        return getattr(m, "prior_module", m) if not hasattr(m, "_arena_params") else m

    def __enter__(self):
        if self.state is None or self.prior_cache is None:
            return
        key = self.state[0] if isinstance(self.state, (tuple, list)) else self.state
        if isinstance(self.prior_cache, DevicePriorCache):
            self.prior_cache.load_into(self._prior_module(), int(key))
        else:
            self.prior_cache.apply_prior(self.model, self.state[1])

    def __exit__(self, exc_type, exc_value, traceback):
        if self.state is None or self.prior_cache is None:
            return False
        key = self.state[0] if isinstance(self.state, (tuple, list)) else self.state
        if isinstance(self.prior_cache, DevicePriorCache):
            self.prior_cache.store_from(self._prior_module(), int(key))
        else:
            self.prior_cache[key] = self.prior_cache.extract_prior(self.model)
        return False
