"""End-to-end statistics pass on RAGGED captions (len ~ U{4..W}, W = 77 by default; `probe_ragged.py N W` for shorter ones — real
caption data averages a dozen tokens): tokens/s with loader batches of 256 captions regrouped into token-budget blocks
(default) vs one block per loader batch (block_tokens=0)."""
import json, os, sys, tempfile, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emcid_b200 import layer_stats, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 77
dev = torch.device("cuda:0")
model = synth.make_text_encoder("sd-text", seed=0).to(dev)
names = [f"text_model.encoder.layers.{l}.mlp.fc2" for l in (7, 8, 9, 10, 11)]
caps = synth.make_caption_ids(N, seed=11, full=False, min_len=4, width=W)
tokens = sum(len(c) for c in caps)
layer_stats.get_ccs_filtered_ds = lambda tokenizer: synth.CaptionIdDataset(caps)
out = {"captions": N, "tokens": tokens, "mean_len": tokens / N}
for label, bt in (("warmup", 37888), ("token_budget_blocks", 37888), ("one_block_per_loader_batch", 0)):
    tmp = tempfile.mkdtemp(prefix="emcid_ragged_")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    stats = layer_stats.layer_stats_text_encoder_multi(model, None, names, stats_dir=tmp, sample_size=N, precision="float32",
                                                       progress=None, force_recompute=True, captions_per_batch=256,
                                                       block_tokens=bt)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    assert stats[names[0]].mom2.count == tokens
    out[label] = {"seconds": dt, "tokens_per_s": tokens / dt, "timeline": dict(layer_stats.LAST_PASS_INFO.get("timing", {}))}
print(json.dumps(out, indent=1))
