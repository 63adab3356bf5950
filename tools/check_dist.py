"""Multi-GPU checks of the statistics entry points on the NCCL path (run under torchrun on a GPU box):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_dist.py

Every rank computes the single-process answer first (distributed=False), then the sharded one, and compares:
  1. layer_stats_text_encoder (reference signature): every rank gets the reduced matrix (broadcast), counts exact;
  2. layer_stats_text_encoder_multi: layer i only on rank i mod world, counts everywhere, npz written once;
  3. the same with checkpoints every 2 blocks: per-rank .resume files appear and are gone at the end, same numbers;
  4. layer_stats_cross_attn_kv: one pass, every K/V module's file written, every rank gets the matrix.
Prints one line per check on rank 0 and exits non-zero on a mismatch."""
import os
import sys
import tempfile
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from emcid_b200 import layer_stats, synth  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402  (test infrastructure: the miniature UNet stand-in)

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
ok = True


def report(name, good, detail=""):
    global ok
    flag = torch.tensor([1 if good else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok = ok and bool(flag.item())
    if rank == 0:
        print(f"{'ok  ' if flag.item() else 'FAIL'} {name} {detail}", flush=True)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


model = rh.make_clip_text_model("clip-l", seed=0, num_hidden_layers=9).to(dev)
caps = synth.make_caption_ids(3000, seed=41, full=False, min_len=4)
layer_stats.get_ccs_filtered_ds = lambda tokenizer: synth.CaptionIdDataset(caps)
names = [f"text_model.encoder.layers.{l}.mlp.fc2" for l in (6, 7, 8)]
base = tempfile.mkdtemp(prefix=f"check_dist_r{rank}_")
shared = [tempfile.mkdtemp(prefix="check_dist_shared_") if rank == 0 else None]
dist.broadcast_object_list(shared, src=0)
shared = shared[0]
kw = dict(sample_size=2500, precision="float32", progress=None, num_workers=0, block_tokens=8192)

single = layer_stats.layer_stats_text_encoder_multi(model, None, names, stats_dir=os.path.join(base, "single"),
                                                    distributed=False, **kw)
total = single[names[0]].mom2.count

# 1. reference-shaped call: every rank ends with the matrix
st = layer_stats.layer_stats_text_encoder(model, None, names[0], stats_dir=os.path.join(shared, "one"), **kw)
report("layer_stats_text_encoder (broadcast)", st.mom2.count == total and st.mom2.mom2 is not None
       and rel(st.mom2.mom2, single[names[0]].mom2.mom2) < 2e-6, layer_stats.LAST_PASS_INFO.get("exchange", ""))

# 2. multi-layer call: roots only
multi = layer_stats.layer_stats_text_encoder_multi(model, None, names, stats_dir=os.path.join(shared, "multi"), **kw)
good = True
for i, n in enumerate(names):
    good &= multi[n].mom2.count == total
    if rank == i % world:
        good &= multi[n].mom2.mom2 is not None and rel(multi[n].mom2.mom2, single[n].mom2.mom2) < 2e-6
    else:
        good &= multi[n].mom2.mom2 is None
dist.barrier()
files = [f for _, _, fs in os.walk(os.path.join(shared, "multi")) for f in fs]
report("layer_stats_text_encoder_multi (roots only)", good and len(files) == len(names), f"{len(files)} files")

# 3. checkpointed pass
ck = layer_stats.layer_stats_text_encoder_multi(model, None, names, stats_dir=os.path.join(shared, "ckpt"),
                                                checkpoint_every=2, broadcast=True, **kw)
dist.barrier()
left = [f for _, _, fs in os.walk(os.path.join(shared, "ckpt")) for f in fs if f.startswith(".resume_")]
good = all(ck[n].mom2.count == total and rel(ck[n].mom2.mom2, single[n].mom2.mom2) < 2e-6 for n in names) and not left
report("checkpointed pass (every 2 blocks)", good, f"leftover checkpoint files: {left}")

# 4. cross-attention K/V statistics
unet = rh.TinyUNet(768, seed=3).to(dev)
pipe = SimpleNamespace(text_encoder=model, unet=unet, tokenizer=None, device=dev)
kv = layer_stats.get_all_cross_attn_kv_layer_names(pipe)
one = layer_stats.layer_stats_cross_attn_kv(pipe, kv[0], stats_dir=os.path.join(base, "kv_single"), distributed=False,
                                            sample_size=2500, precision="float32", progress=None, num_workers=0)
st = layer_stats.layer_stats_cross_attn_kv(pipe, kv[0], stats_dir=os.path.join(shared, "kv"), sample_size=2500,
                                           precision="float32", progress=None, num_workers=0)
dist.barrier()
files = [f for _, _, fs in os.walk(os.path.join(shared, "kv")) for f in fs]
report("layer_stats_cross_attn_kv", st.mom2.count == one.mom2.count and rel(st.mom2.mom2, one.mom2.mom2) < 2e-6
       and len(files) == len(kv), f"{len(files)} files for {len(kv)} modules")

dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
