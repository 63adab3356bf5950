"""Hyper-parameters of the edit, loadable from the reference's own ``hparams/*.json`` files.

Mirror of emcid/emcid_hparams.py::EMCIDHyperParams / EMCIDXLHyperParams (:55-276) and util/hparams.py::
HyperParams.from_json (:11-16) as far as the hot path is concerned: every key of the JSON becomes an attribute
(the stage-1 optimisation fields are carried along untouched), and the fields the stage-2 loop reads
(emcid_main.py:846-1065) have the reference's defaults when a file omits them.
"""
from __future__ import annotations

import json
from typing import Any, Dict, List

_REQUIRED = ("layers", "mom2_update_weight", "rewrite_module_tmp", "mom2_dataset", "mom2_n_samples", "mom2_dtype")
_DEFAULTS: Dict[str, Any] = {
    # emcid_hparams.py:87-105
    "use_new_compute_z": False, "num_edit_tokens": 1, "edit_weight": 0.5, "objective": "ablate-dest",
    "sld_supervision": False,
}


class EMCIDHyperParams:
    """Attribute bag with the reference's field names.  ``layers`` / ``mom2_update_weight`` drive text encoder 1,
    ``layers_2`` / ``mom2_update_weight_2`` (SDXL files) text encoder 2."""

    def __init__(self, **fields: Any):
        missing = [k for k in _REQUIRED if k not in fields]
        if missing:
            raise TypeError(f"EMCIDHyperParams: missing field(s) {missing}")   # dataclass __init__ raises TypeError too
        for k, v in {**_DEFAULTS, **fields}.items():
            setattr(self, k, v)
        self.layers: List[int] = list(self.layers)
        if hasattr(self, "layers_2"):
            self.layers_2 = list(self.layers_2)

    @classmethod
    def from_json(cls, fpath):
        with open(fpath, "r") as f:
            return cls(**json.load(f))

    def __repr__(self) -> str:
        return f"{type(self).__name__}({', '.join(f'{k}={v!r}' for k, v in sorted(vars(self).items()))})"


class EMCIDXLHyperParams(EMCIDHyperParams):
    """SDXL variant (emcid_hparams.py:166-276): additionally requires the encoder-2 fields."""

    def __init__(self, **fields: Any):
        for k in ("layers_2", "mom2_update_weight_2"):
            if k not in fields:
                raise TypeError(f"EMCIDXLHyperParams: missing field {k!r}")
        super().__init__(**fields)
