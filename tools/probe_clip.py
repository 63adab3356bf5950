"""GPU probe of the native text-encoder forward: hidden states vs HF (fp32 and fp64), statistics vs the
hook path and an fp64 reference, and throughput of the whole statistics step."""
import copy
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emcid_b200 import clip_forward, layer_stats, synth  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
res = []


def tiny_model(act="quick_gelu", seed=0):
    from transformers import CLIPTextConfig, CLIPTextModel
    torch.manual_seed(seed)
    cfg = CLIPTextConfig(vocab_size=1000, hidden_size=64, intermediate_size=256, num_hidden_layers=2,
                         num_attention_heads=4, max_position_embeddings=77, hidden_act=act, bos_token_id=998,
                         eos_token_id=999)
    return CLIPTextModel(cfg).eval()


def padded(caps):
    B, L = len(caps), max(len(c) for c in caps)
    ids = torch.zeros(B, L, dtype=torch.long)
    pos = torch.zeros(B, L, dtype=torch.long)
    mask = torch.zeros(B, L, dtype=torch.long)
    for i, c in enumerate(caps):
        ids[i, :len(c)] = c
        pos[i, :len(c)] = torch.arange(len(c))
        mask[i, :len(c)] = 1
    return {"input_ids": ids, "position_ids": pos, "attention_mask": mask}


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def check_hidden(name, model, caps):
    model = model.to(dev)
    m64 = copy.deepcopy(model).double()
    batch = {k: v.to(dev) for k, v in padded(caps).items()}
    keep = batch["attention_mask"].bool()
    with torch.no_grad():
        hs32 = model(**batch, output_hidden_states=True).hidden_states
        hs64 = m64(**batch, output_hidden_states=True).hidden_states
    ids, pos, cu, S, T = clip_forward.pack_batch(batch, 77)
    nat = clip_forward.NativeClipTextEncoder(model, T, S)
    row = {"name": name, "T": T, "S": S}
    for n in sorted({0, 1, len(hs32) - 1}):
        h = nat.forward_hidden(ids, pos, cu, S, T, n)
        torch.cuda.synchronize()
        row[f"native_vs_fp64_L{n}"] = rel(h, hs64[n][keep])
        row[f"hf32_vs_fp64_L{n}"] = rel(hs32[n][keep], hs64[n][keep])
    nat.close()
    print(json.dumps(row), flush=True)
    res.append(row)


def check_stats(name, model, caps, layers, block=64):
    model = model.to(dev)
    m64 = copy.deepcopy(model).double()
    names = [f"text_model.encoder.layers.{l}.mlp.fc2" for l in layers]
    # fp64 reference of the statistics
    ref = {}
    feats = {}
    hooks = [m64.text_model.encoder.layers[l].mlp.fc2.register_forward_pre_hook(
        lambda m, a, l=l: feats.__setitem__(l, a[0])) for l in layers]
    tot = {l: 0 for l in layers}
    cnt = 0
    with torch.no_grad():
        for i in range(0, len(caps), block):
            batch = {k: v.to(dev) for k, v in padded(caps[i:i + block]).items()}
            m64(**batch)
            keep = batch["attention_mask"].bool()
            cnt += int(keep.sum())
            for l in layers:
                a = feats[l][keep]
                tot[l] = tot[l] + a.T @ a
    for h in hooks:
        h.remove()
    row = {"name": name, "count_ref": cnt}
    for mode in ("native", "hooks"):
        runner = layer_stats.TextEncoderMom2Pass(model, names, native=(mode == "native"))
        for i in range(0, len(caps), block):
            runner.run_batch(padded(caps[i:i + block]))
        out = runner.finalize()
        torch.cuda.synchronize()
        row[mode + "_is_native"] = runner._native is not None
        for l, n in zip(layers, names):
            row[f"{mode}_L{l}"] = rel(out[n][0], tot[l])
            row[f"{mode}_count_L{l}"] = int(out[n][1])
        runner.close()
    print(json.dumps(row), flush=True)
    res.append(row)


caps_tiny = synth.make_caption_ids(150, vocab=1000, seed=3, full=False, min_len=2)
check_hidden("tiny_quick_gelu", tiny_model("quick_gelu"), caps_tiny)
check_hidden("tiny_gelu", tiny_model("gelu", 1), caps_tiny[:37])
check_stats("tiny_stats", tiny_model("quick_gelu"), caps_tiny, [0, 1])
check_stats("tiny_gelu_stats", tiny_model("gelu", 1), caps_tiny, [1])

clipl = synth.make_text_encoder("sd-text", seed=0)
caps_l = synth.make_caption_ids(96, seed=5, full=False)
check_hidden("clipl", clipl, caps_l)
check_stats("clipl_stats", clipl, caps_l, [7, 8, 9, 10, 11], block=48)

# ---- throughput of the statistics step, native vs hooks
model = clipl.to(dev)
names = [f"text_model.encoder.layers.{l}.mlp.fc2" for l in (7, 8, 9, 10, 11)]
for mode, C in (("native", 512), ("native", 1024), ("hooks", 512)):
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(0, 49406, (C, 77), generator=g)
    ids[:, 0] = 49406
    ids[:, -1] = 49407
    batch = {"input_ids": ids.to(dev), "position_ids": torch.arange(77, device=dev).expand(C, 77).contiguous(),
             "attention_mask": torch.ones(C, 77, dtype=torch.long, device=dev)}
    runner = layer_stats.TextEncoderMom2Pass(model, names, native=(mode == "native"))
    for _ in range(2):
        runner.run_batch(batch)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 5
    t0 = time.perf_counter()
    s.record()
    for _ in range(iters):
        runner.run_batch(batch)
    e.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = s.elapsed_time(e) / iters
    row = {"timing": mode, "captions": C, "ms_per_block": ms, "tokens_per_s": C * 77 / ms * 1e3, "wall_ms_per_block": wall / iters * 1e3}
    print(json.dumps(row), flush=True)
    res.append(row)
    runner.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/probe_clip.json", "w"), indent=1)
from emcid_b200 import _lib  # noqa: E402
print("hang_code", _lib.lib().emcid_hang_code())
print("PROBE DONE")
