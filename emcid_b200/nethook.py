"""Minimal module-instrumentation helpers used on the hot path.

Behavioural mirror of the pieces of util/nethook.py the EMCID edit path touches:
`get_module` (:375-382), `get_parameter` (:385-392), `Trace(..., stop=True)` (:22-128) and
`TraceDict` (:131-200).  Implemented with plain forward hooks.
"""
from __future__ import annotations

import contextlib
from collections import OrderedDict

import torch


class StopForward(Exception):
    """Raised by a hook to abandon the rest of a forward pass (reference :203-213)."""


def get_module(model: torch.nn.Module, name: str) -> torch.nn.Module:
    """The submodule with that dotted name; LookupError if there is none (walks the path instead of listing every module:
    the edit loop resolves two names per layer and edit)."""
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


def set_requires_grad(requires_grad: bool, *models) -> None:
    for model in models:
        if isinstance(model, torch.nn.Module):
            for p in model.parameters():
                p.requires_grad = requires_grad
        elif isinstance(model, (torch.nn.Parameter, torch.Tensor)):
            model.requires_grad = requires_grad
        else:
            raise AssertionError("unknown type %r" % type(model))


def _keep(x, clone: bool, detach: bool):
    if isinstance(x, torch.Tensor):
        if detach:
            x = x.detach()
        if clone:
            x = x.clone()
        return x
    if isinstance(x, (tuple, list)):
        return type(x)(_keep(v, clone, detach) for v in x)
    if isinstance(x, dict):
        return type(x)((k, _keep(v, clone, detach)) for k, v in x.items())
    return x


class Trace(contextlib.AbstractContextManager):
    """Retain input and/or output of one named submodule during a forward pass; with stop=True the
    forward is abandoned right after that submodule ran and the StopForward is swallowed on exit."""

    def __init__(self, module, layer=None, retain_output=True, retain_input=False, clone=False,
                 detach=False, retain_grad=False, edit_output=None, stop=False):
        self.layer = layer
        self.stop = stop
        target = get_module(module, layer) if layer is not None else module

        def hook(_m, inputs, output):
            if retain_input:
                self.input = _keep(inputs[0] if len(inputs) == 1 else inputs, clone, detach)
            if edit_output is not None:
                try:
                    output = edit_output(output=output, layer=self.layer)
                except TypeError:
                    output = edit_output(output)
            if retain_output:
                self.output = _keep(output, clone, detach)
                if retain_grad and isinstance(self.output, torch.Tensor):
                    self.output.requires_grad_(True)
                    self.output.retain_grad()
            if stop:
                raise StopForward()
            return output

        self._handle = target.register_forward_hook(hook)

    def __exit__(self, exc_type, exc, tb):
        self.close()
        if self.stop and exc_type is not None and issubclass(exc_type, StopForward):
            return True

    def close(self):
        self._handle.remove()


class TraceDict(OrderedDict, contextlib.AbstractContextManager):
    """One Trace per named layer; the last layer carries the stop flag."""

    def __init__(self, module, layers=None, retain_output=True, retain_input=False, clone=False,
                 detach=False, retain_grad=False, edit_output=None, stop=False):
        super().__init__()
        self.stop = stop
        layers = list(dict.fromkeys(layers or []))
        for i, name in enumerate(layers):
            self[name] = Trace(module, name, retain_output=retain_output, retain_input=retain_input,
                               clone=clone, detach=detach, retain_grad=retain_grad, edit_output=edit_output,
                               stop=stop and i == len(layers) - 1)

    def __exit__(self, exc_type, exc, tb):
        self.close()
        if self.stop and exc_type is not None and issubclass(exc_type, StopForward):
            return True

    def close(self):
        for tr in reversed(list(self.values())):
            tr.close()
