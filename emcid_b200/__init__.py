"""emcid_b200 — B200-native (sm_100a) implementation of EMCID's data-parallel hot path.

Two pieces, behind the reference's own Python API (SilentView/EMCID):
  * the mom2 statistics pass  (emcid/layer_stats.py, util/runningstats.py::SecondMoment)
  * the closed-form multi-layer weight update (emcid/emcid_main.py, the upd_matrix block)
The arithmetic lives in hand-written CUDA (``csrc/``) behind the C ABI in ``include/emcid_b200.h``.
"""

__version__ = "0.1.0"

# the package root re-exports the edit entry points like the reference's emcid/__init__.py:1
from .emcid_main import (apply_emcid_to_clip, apply_emcid_to_cross_attn, apply_emcid_to_sdxl_text_encoders,  # noqa: E402,F401
                         apply_emcid_to_text_encoder, clear_factor_cache, execute_emcid_clip, execute_emcid_cross_attn,
                         execute_emcid_text_encoder)
from .emcid_hparams import EMCIDHyperParams, EMCIDXLHyperParams  # noqa: E402,F401
from .layer_stats import (layer_stats_cross_attn_kv, layer_stats_text_encoder,  # noqa: E402,F401
                          layer_stats_text_encoder_multi)
