"""Where the host time of the edit calls goes (cProfile over bench.py's edit workloads); run on the GPU box."""
import cProfile
import io
import os
import pstats
import sys
import tempfile
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from emcid_b200 import emcid_main, synth  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
LAYERS = [7, 8, 9, 10, 11]
model = synth.make_text_encoder("sd-text", seed=0).to(dev)
names = [f"text_model.encoder.layers.{l}.mlp.fc2" for l in LAYERS]
g = torch.Generator(device=dev).manual_seed(0)
for nm in names:   # any SPD covariance will do for a host-time profile
    A = torch.randn(3072, 4096, device=dev, generator=g)
    emcid_main.COV_CACHE[(model.config._name_or_path.replace("/", "_"), nm)] = (A @ A.T / 4096).float()
tok = synth.WordHashTokenizer(49408)
pipe = SimpleNamespace(text_encoder=model, tokenizer=tok, device=dev)
tmp = tempfile.mkdtemp()
hp = synth.make_edit_hparams(LAYERS, mom2_n_samples=1)


def run(n_edits, per_edit, label):
    reqs = [[dict(r, source=f"edit{e} {r['source']}") for r in synth.make_edit_requests(per_edit)] for e in range(n_edits)]
    cache = os.path.join(tmp, label, "c_")
    for rq in reqs:
        synth.write_vstar_cache(cache, rq, 768, seed=3)
    for timed in (False, True):
        prof = cProfile.Profile()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if timed:
            prof.enable()
        for rq in reqs:
            emcid_main.apply_emcid_to_text_encoder(pipe, rq, hp, device=dev, cache_name=cache, stats_dir=tmp, verbose=False)
        torch.cuda.synchronize()
        if timed:
            prof.disable()
        dt = time.perf_counter() - t0
    out = io.StringIO()
    pstats.Stats(prof, stream=out).sort_stats("cumulative").print_stats(38)
    print(f"==== {label}: {n_edits} x {per_edit} concepts: {1e3 * dt:.1f} ms wall (profiled pass)")
    print(out.getvalue()[:9000])


import contextlib  # noqa: E402

with contextlib.redirect_stdout(sys.stderr):
    pass
run(10, 100, "sequential")
run(1, 1000, "single")
