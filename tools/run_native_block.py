"""One statistics block (C captions x 77 tokens, sd-text layers 7-11) through the native forward; for ncu."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emcid_b200 import layer_stats, synth
C = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
model = synth.make_text_encoder("sd-text", seed=0).to(dev)
names = [f"text_model.encoder.layers.{l}.mlp.fc2" for l in (7, 8, 9, 10, 11)]
g = torch.Generator().manual_seed(1)
ids = torch.randint(0, 49406, (C, 77), generator=g); ids[:, 0] = 49406; ids[:, -1] = 49407
batch = {"input_ids": ids.to(dev), "position_ids": torch.arange(77, device=dev).expand(C, 77).contiguous(),
         "attention_mask": torch.ones(C, 77, dtype=torch.long, device=dev)}
runner = layer_stats.TextEncoderMom2Pass(model, names)
for _ in range(reps):
    runner.run_batch(batch)
torch.cuda.synchronize()
out = runner.finalize()
torch.cuda.synchronize()
print("count", int(out[names[0]][1]))
