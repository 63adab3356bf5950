"""Where does the end-to-end statistics pass (host captions -> host mom2) spend its time?
Times the stages of layer_stats_text_encoder_multi separately (each bracketed by a device synchronize)."""
import json, os, sys, tempfile, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emcid_b200 import layer_stats, synth
from emcid_b200.runningstats import FixedSubsetSampler, subset_indices
from emcid_b200.stat_dataset import packed_collation

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
BLK = int(sys.argv[2]) if len(sys.argv) > 2 else 512
WORKERS = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0")
model = synth.make_text_encoder("sd-text", seed=0).to(dev)
names = [f"text_model.encoder.layers.{l}.mlp.fc2" for l in (7, 8, 9, 10, 11)]
out = {"captions": N, "block": BLK, "workers": WORKERS}


def tick():
    torch.cuda.synchronize()
    return time.perf_counter()


t0 = tick()
caps = synth.make_caption_ids(N, seed=7, full=True)
out["make_captions_s"] = tick() - t0
ds = synth.CaptionIdDataset(caps)

# (1) the loader alone
t0 = tick()
idx = subset_indices(len(ds), N, random_sample=1)
loader = torch.utils.data.DataLoader(ds, sampler=FixedSubsetSampler(idx), batch_size=BLK, collate_fn=packed_collation(),
                                     num_workers=WORKERS, pin_memory=True)
it = iter(loader)
first = next(it)
out["loader_first_batch_s"] = tick() - t0
n = 1
for b in it:
    n += 1
out["loader_all_batches_s"] = tick() - t0
out["loader_batches"] = n

# (2) runner construction (accumulators + native encoder is lazy: first run_batch builds it)
t0 = tick()
runner = layer_stats.TextEncoderMom2Pass(model, names)
out["runner_ctor_s"] = tick() - t0
t0 = tick()
runner.run_batch(first)
out["first_block_s"] = tick() - t0
t0 = tick()
for _ in range(4):
    runner.run_batch(first)
out["steady_block_s"] = (tick() - t0) / 4
t0 = tick()
res = runner.finalize()
out["finalize_s"] = tick() - t0
t0 = tick()
host = {k: v[0].to("cpu") for k, v in res.items()}
out["d2h_pageable_s"] = tick() - t0
t0 = tick()
runner.close()
out["close_s"] = tick() - t0

# (3) the public call, twice (second call: allocator warm)
layer_stats.get_ccs_filtered_ds = lambda tokenizer: ds
for rep in range(2):
    tmp = tempfile.mkdtemp(prefix="emcid_e2e_")
    t0 = tick()
    stats = layer_stats.layer_stats_text_encoder_multi(model, None, names, stats_dir=tmp, sample_size=N, precision="float32",
                                                       progress=None, force_recompute=True, captions_per_batch=BLK,
                                                       num_workers=WORKERS)
    dt = tick() - t0
    out[f"public_call_{rep}_s"] = dt
    out[f"public_call_{rep}_tokens_per_s"] = N * 77 / dt
    out[f"public_call_{rep}_timing"] = dict(layer_stats.LAST_PASS_INFO.get("timing", {}))
print(json.dumps(out, indent=1))
