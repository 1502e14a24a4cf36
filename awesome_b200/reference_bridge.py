"""Bridge to the reference package when both live in one process (``scripts/run.py`` with ``prior_model_type:
awesome_b200....``).

The reference gates its pretrain protocol on ``isinstance(prior_module, PretrainableModule)``
(``awesome/model/wrapper_module.py:325-340``, ``awesome/agent/torch_agent.py:562-563``) -- a plain marker class, not an
ABC, so a drop-in module has to have it among its bases.  ``integrate_with_reference()`` appends the reference's
``PretrainableModule`` to the bases of the drop-in prior classes (idempotent; a no-op when the reference is not
importable).  It runs on ``import awesome_b200`` when ``awesome`` is already imported -- the order ``AwesomeRunner``
produces (it resolves the dotted ``prior_model_type`` after its own imports) -- and on the first construction of a prior
module otherwise."""
from __future__ import annotations

import sys

_done = False


def integrate_with_reference(force: bool = False) -> bool:
    global _done
    if _done and not force:
        return True
    if not force and "awesome" not in sys.modules and "awesome.model.pretrainable_module" not in sys.modules:
        return False
    try:
        from awesome.model.pretrainable_module import PretrainableModule
    except Exception:
        return False
    from .model.convex_diffeomorphism_net import ConvexDiffeomorphismNet
    from .model.multi_prior import NumberBasedMultiPriorModule
    from .model.path_connected_net import PathConnectedNet
    for cls in (PathConnectedNet, ConvexDiffeomorphismNet, NumberBasedMultiPriorModule):
        if PretrainableModule not in cls.__mro__:
            cls.__bases__ = cls.__bases__ + (PretrainableModule,)
    _done = True
    return True
