"""emcid_b200 — B200-native (sm_100a) implementation of EMCID's data-parallel hot path.

Two pieces, behind the reference's own Python API (SilentView/EMCID):
  * the mom2 statistics pass  (emcid/layer_stats.py, util/runningstats.py::SecondMoment)
  * the closed-form multi-layer weight update (emcid/emcid_main.py, the upd_matrix block)
The arithmetic lives in hand-written CUDA (``csrc/``) behind the C ABI in ``include/emcid_b200.h``.
"""

__version__ = "0.1.0"
