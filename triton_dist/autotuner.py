"""Contextual autotuner: tune a *whole function* that contains several tunable ops as one unit.

Reference: /root/reference/python/triton_dist/autotuner.py:43-250 (``contextual_autotune``: re-runs the wrapped
function once per candidate of every inner ``triton.autotune`` kernel, keeps ranks in lock step).  Our tunable ops
take explicit ``GemmConfig`` objects, so the context manager installs per-op overrides that the ops consult."""
from __future__ import annotations

import contextlib
import itertools
from typing import Any, Callable, Dict, List

from .tune import _time_call

_ACTIVE: Dict[str, Any] = {}


def override_for(op_name: str, default=None):
    """Ops call this to pick up a config installed by an enclosing ``contextual_autotune`` run."""
    return _ACTIVE.get(op_name, default)


@contextlib.contextmanager
def _install(assign: Dict[str, Any]):
    old = dict(_ACTIVE)
    _ACTIVE.update(assign)
    try:
        yield
    finally:
        _ACTIVE.clear()
        _ACTIVE.update(old)


def contextual_autotune(spaces: Dict[str, List[Any]], is_dist: bool = False, pg=None, warmup: int = 3, rep: int = 5):
    """``spaces``: {op_name: [candidate configs]}.  The decorated function is timed for every combination (cartesian
    product, as in the reference's n_repeat x n_configs sweep) and afterwards always runs with the best one."""
    def deco(fn: Callable):
        best: Dict[str, Any] = {}

        def wrapped(*args, **kwargs):
            if not best:
                names = list(spaces)
                results = []
                for combo in itertools.product(*[spaces[n] for n in names]):
                    assign = dict(zip(names, combo))
                    with _install(assign):
                        ms = _time_call(lambda: fn(*args, **kwargs), warmup, rep, pg if is_dist else None)
                    results.append((ms, assign))
                best.update(min(results, key=lambda r: r[0])[1])
            with _install(best):
                return fn(*args, **kwargs)
        wrapped.best = best
        return wrapped
    return deco


class ContextualAutoTuner:
    """Class form of :func:`contextual_autotune` (reference: autotuner.py ``ContextualAutoTuner``): wraps a function, tunes on the first
    call, exposes the winning assignment and the whole timing table."""

    def __init__(self, fn: Callable, spaces: Dict[str, List[Any]], is_dist: bool = False, pg=None, warmup: int = 3, rep: int = 5):
        self.fn, self.spaces, self.is_dist, self.pg, self.warmup, self.rep = fn, spaces, is_dist, pg, warmup, rep
        self.best: Dict[str, Any] = {}
        self.results: List[Any] = []

    def tune(self, *args, **kwargs):
        names = list(self.spaces)
        self.results = []
        for combo in itertools.product(*[self.spaces[n] for n in names]):
            assign = dict(zip(names, combo))
            with _install(assign):
                ms = _time_call(lambda: self.fn(*args, **kwargs), self.warmup, self.rep, self.pg if self.is_dist else None)
            self.results.append((ms, assign))
        self.best = dict(min(self.results, key=lambda r: r[0])[1])
        return self.best

    def __call__(self, *args, **kwargs):
        if not self.best:
            self.tune(*args, **kwargs)
        with _install(self.best):
            return self.fn(*args, **kwargs)
