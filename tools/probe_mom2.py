"""GPU probe of the mom2 accumulator against a torch fp64 Gram of the same activations."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emcid_b200.mom2 import Mom2Accumulator  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
out = []


def ref_mom2(X, valid, W, b, act):
    Xv = X.reshape(-1, X.shape[-1])
    if valid is not None:
        Xv = Xv[valid.reshape(-1) != 0]
    z = Xv.double() @ W.double().T + b.double()
    if act == "quick_gelu":
        a = z * torch.sigmoid(1.702 * z)
    else:
        a = torch.nn.functional.gelu(z)
    z32 = torch.nn.functional.linear(Xv, W, b)
    a32 = z32 * torch.sigmoid(1.702 * z32) if act == "quick_gelu" else torch.nn.functional.gelu(z32)
    return a.T @ a, Xv.shape[0], a32.T @ a32


def run(name, d, h, T, act="quick_gelu", frac_valid=1.0, calls=1, slab=0, seed=0, **kw):
    g = torch.Generator(device="cuda").manual_seed(seed)
    W = torch.randn(d, h, device="cuda", generator=g) * (0.7 / h ** 0.5)
    b = torch.randn(d, device="cuda", generator=g) * 0.1
    acc = Mom2Accumulator("cuda:0", d, h, act, slab_tokens=slab, **kw)
    acc.set_weights(W, b)
    tot = torch.zeros(d, d, device="cuda", dtype=torch.float64)
    tot32 = torch.zeros(d, d, device="cuda", dtype=torch.float32)
    n = 0
    for c in range(calls):
        X = torch.randn(T, h, device="cuda", generator=g)
        valid = None
        if frac_valid < 1.0:
            valid = (torch.rand(T, device="cuda", generator=g) < frac_valid)
        acc.add(X, valid)
        m, k, m32 = ref_mom2(X, valid, W, b, act)
        tot += m
        tot32 += m32
        n += k
    mom2, count = acc.finalize()
    torch.cuda.synchronize()
    err = float((mom2.double() - tot).norm() / tot.norm())
    err32 = float((tot32.double() - tot).norm() / tot.norm())
    sym = float((mom2 - mom2.T).abs().max())
    line = dict(name=name, d=d, h=h, T=T, calls=calls, act=act, count=int(count), count_ref=n,
                rel_fro=err, torch_fp32_rel_fro=err32, asym=sym, **{k: v for k, v in kw.items()})
    print(json.dumps(line), flush=True)
    out.append(line)
    acc.close()


run("tiny", 256, 64, 300)
run("tiny_masked", 256, 64, 1000, frac_valid=0.6)
run("tiny_gelu", 256, 64, 777, act="gelu", frac_valid=0.8)
run("odd_dims", 200, 80, 500, frac_valid=0.5)
run("clipl_1slab", 3072, 768, 1536)
run("clipl_masked_multi", 3072, 768, 9856, frac_valid=0.55, calls=2)
run("clipl_chunks_2_4", 3072, 768, 9856, calls=2, fc1_chunk=2, syrk_chunk=4)
run("clipl_20calls", 3072, 768, 9856, calls=20)
run("bigg", 5120, 1280, 4096, act="gelu", frac_valid=0.9)

# ---- timing: steady-state tokens/s of the accumulate call alone (X resident, L2-sized slabs)
res = {"checks": out, "timings": []}
for (d, h, T, slab, act) in [(3072, 768, 9856, 0, "quick_gelu"), (3072, 768, 9856, 1024, "quick_gelu"),
                             (3072, 768, 9856, 2048, "quick_gelu"), (3072, 768, 12288, 0, "quick_gelu"),
                             (5120, 1280, 9856, 1024, "gelu")]:
    g = torch.Generator(device="cuda").manual_seed(1)
    W = torch.randn(d, h, device="cuda", generator=g) * (0.7 / h ** 0.5)
    b = torch.randn(d, device="cuda", generator=g) * 0.1
    acc = Mom2Accumulator("cuda:0", d, h, act, slab_tokens=slab)
    acc.set_weights(W, b)
    X = torch.randn(T, h, device="cuda", generator=g)
    for _ in range(3):
        acc.add(X)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    s.record()
    for _ in range(iters):
        acc.add(X)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    F = 2.0 * h * d + d * (d + 1.0)
    line = dict(d=d, h=h, T=T, slab=slab, ms=ms, tokens_per_s=T / ms * 1e3, algo_tflops=T * F / ms / 1e9,
                issued_tflops=3 * T * F / ms / 1e9)
    print(json.dumps(line), flush=True)
    res["timings"].append(line)
    acc.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/probe_mom2.json", "w"), indent=1)
print("PROBE DONE")
