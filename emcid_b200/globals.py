"""Path defaults of the reference (globals.yml / util/globals.py:8-38).  A globals.yml in the CWD
overrides them, as it does for the reference."""
from pathlib import Path

_DEFAULTS = dict(RESULTS_DIR="results", DATA_DIR="data", STATS_DIR="data/stats", HPARAMS_DIR="hparams",
                 CACHE_DIR="cache", XL_STATS_DIR1="data/stats/sdxl/text1", XL_STATS_DIR2="data/stats/sdxl/text2",
                 EDITING_PROMPTS_CNT=3)


def _load():
    vals = dict(_DEFAULTS)
    p = Path("globals.yml")
    if p.exists():
        try:
            import yaml

            data = yaml.safe_load(p.read_text()) or {}
            vals.update({k: data[k] for k in _DEFAULTS if k in data})
        except Exception:
            pass
    return vals


_v = _load()
RESULTS_DIR, DATA_DIR, STATS_DIR, HPARAMS_DIR, CACHE_DIR, XL_STATS_DIR1, XL_STATS_DIR2 = (
    Path(_v[k]) for k in ("RESULTS_DIR", "DATA_DIR", "STATS_DIR", "HPARAMS_DIR", "CACHE_DIR", "XL_STATS_DIR1",
                          "XL_STATS_DIR2"))
EDITING_PROMPTS_CNT = _v["EDITING_PROMPTS_CNT"]
