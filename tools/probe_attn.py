"""Minimal check of the tensor-core attention path: CLIP-L, a few captions, one layer vs HF fp64."""
import copy, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emcid_b200 import clip_forward, synth, _lib
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
n_caps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
model = synth.make_text_encoder("sd-text", seed=0).to(dev)
caps = synth.make_caption_ids(n_caps, seed=5, full=False)
B, L = len(caps), max(len(c) for c in caps)
ids = torch.zeros(B, L, dtype=torch.long); pos = torch.zeros(B, L, dtype=torch.long); mask = torch.zeros(B, L, dtype=torch.long)
for i, c in enumerate(caps):
    ids[i, :len(c)] = c; pos[i, :len(c)] = torch.arange(len(c)); mask[i, :len(c)] = 1
batch = {"input_ids": ids.to(dev), "position_ids": pos.to(dev), "attention_mask": mask.to(dev)}
keep = batch["attention_mask"].bool()
with torch.no_grad():
    hs64 = copy.deepcopy(model).double()(**batch, output_hidden_states=True).hidden_states
p = clip_forward.pack_batch(batch, 77)
nat = clip_forward.NativeClipTextEncoder(model, p[4], p[3])
for n in (1, 2, 12):
    h = nat.forward_hidden(*p, n)
    torch.cuda.synchronize()
    ref = hs64[n][keep]
    print(json.dumps({"layers": n, "T": p[4], "rel": float((h.double() - ref).norm() / ref.norm()),
                      "nan": int(torch.isnan(h).sum())}), flush=True)
print("hang_code", hex(_lib.lib().emcid_hang_code()))
