"""The little module instrumentation the EMCID hot path needs.

The reference's util/nethook.py is a general tracing toolkit (Trace / TraceDict with output editing, gradient retention,
subsequence extraction, ...).  The edit path touches three things of it, restated here from what callers observe:

  get_module(model, "a.b.c")        the submodule, LookupError when there is none            (util/nethook.py:375-382)
  get_parameter(model, "a.b.weight") the parameter, LookupError when there is none          (:385-392)
  TraceDict(model, [names], retain_input=True, retain_output=True)                           (:131-200)
                                    context manager; td[name].input / td[name].output hold what the named submodule
                                    received (its first positional argument) and returned during the forward passes inside
                                    the `with` block.  Used by the traced fallback of the key extraction only; the
                                    statistics pass hooks the MLP itself (layer_stats.TextEncoderMom2Pass).
  StopForward                       raised by a hook to abandon the rest of a forward pass    (:203-213)
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Iterable

import torch


class StopForward(Exception):
    """Raised inside a forward hook: nothing after the hooked module needs to run."""


def get_module(model: torch.nn.Module, name: str) -> torch.nn.Module:
    """Walks the dotted path (O(depth)): the edit loop resolves two names per layer and edit."""
    try:
        return model.get_submodule(name)
    except AttributeError:
        raise LookupError(name) from None


def get_parameter(model: torch.nn.Module, name: str) -> torch.nn.Parameter:
    owner, _, leaf = name.rpartition(".")
    try:
        p = (model.get_submodule(owner) if owner else model)._parameters.get(leaf)
    except AttributeError:
        p = None
    if p is None:
        raise LookupError(name)
    return p


class TraceDict(dict):
    """{name: record} with record.input / record.output of every named submodule, refreshed by each forward pass run inside
    the `with` block; `stop=True` abandons a forward once the LAST named submodule has run."""

    def __init__(self, module: torch.nn.Module, layers: Iterable[str] = (), retain_output: bool = True,
                 retain_input: bool = False, stop: bool = False):
        super().__init__()
        self._handles = []
        self._stop = stop
        names = list(dict.fromkeys(layers))
        for i, name in enumerate(names):
            record = self[name] = SimpleNamespace(input=None, output=None)
            last = stop and i == len(names) - 1

            def hook(_mod, args, out, record=record, last=last):
                if retain_input:
                    record.input = args[0] if len(args) == 1 else args
                if retain_output:
                    record.output = out
                if last:
                    raise StopForward()

            self._handles.append(get_module(module, name).register_forward_hook(hook))

    def __enter__(self) -> "TraceDict":
        return self

    def __exit__(self, exc_type, exc, tb) -> bool:
        self.close()
        return bool(self._stop and exc_type is not None and issubclass(exc_type, StopForward))

    def close(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []
