"""GPU probe: 16-bit split planes (fp16 hi + bf16/fp16 lo, tcgen05.mma.kind::f16) against 3xTF32.
Checks that mixed operand formats are accepted by the hardware, the accuracy of the generic GEMM and of
the mom2 pass in each precision mode, and their throughput."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emcid_b200 import _lib  # noqa: E402
from emcid_b200.mom2 import Mom2Accumulator  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
res = {"gemm": [], "mom2": [], "timings": []}

# ---- generic GEMM accuracy
for (M, N, K, scale) in [(128, 256, 64, 1.0), (200, 260, 72, 1.0), (512, 768, 768, 1.0), (1024, 1024, 2048, 1.0),
                         (512, 768, 768, 0.02), (512, 768, 768, 300.0)]:
    g = torch.Generator(device=dev).manual_seed(M + N + K)
    A = torch.randn(M, K, device=dev, generator=g) * scale
    B = torch.randn(N, K, device=dev, generator=g)
    ref = A.double() @ B.double().T
    row = dict(M=M, N=N, K=K, scale=scale)
    for name, kw in [("tf32x3", {}), ("f16x3", dict(f16=True)), ("f16x3_chunk1", dict(f16=True, chunk=1)),
                     ("f16x3_chunk4", dict(f16=True, chunk=4))]:
        try:
            C = _lib.gemm3x_nt(A, B, **kw)
            torch.cuda.synchronize()
            row[name] = float((C.double() - ref).norm() / ref.norm())
        except Exception as e:  # noqa: BLE001
            row[name] = f"ERR {e}"
    row["torch_fp32"] = float(((A @ B.T).double() - ref).norm() / ref.norm())
    print(json.dumps(row), flush=True)
    res["gemm"].append(row)


def ref_mom2(X, valid, W, b, act):
    Xv = X.reshape(-1, X.shape[-1])
    if valid is not None:
        Xv = Xv[valid.reshape(-1) != 0]
    z = Xv.double() @ W.double().T + b.double()
    a = z * torch.sigmoid(1.702 * z) if act == "quick_gelu" else torch.nn.functional.gelu(z)
    return a.T @ a, Xv.shape[0]


def run(name, d, h, T, act="quick_gelu", frac_valid=1.0, calls=1, seed=0, wscale=None, **kw):
    g = torch.Generator(device=dev).manual_seed(seed)
    W = torch.randn(d, h, device=dev, generator=g) * (wscale if wscale else 0.7 / h ** 0.5)
    b = torch.randn(d, device=dev, generator=g) * 0.1
    row = dict(name=name, d=d, h=h, T=T, calls=calls, act=act)
    for prec in ("tf32x3", "f16x3"):
        gg = torch.Generator(device=dev).manual_seed(seed + 1)
        acc = Mom2Accumulator(dev, d, h, act, precision=prec, **kw)
        acc.set_weights(W, b)
        tot = torch.zeros(d, d, device=dev, dtype=torch.float64)
        n = 0
        for c in range(calls):
            X = torch.randn(T, h, device=dev, generator=gg)
            valid = (torch.rand(T, device=dev, generator=gg) < frac_valid) if frac_valid < 1.0 else None
            acc.add(X, valid)
            m, k = ref_mom2(X, valid, W, b, act)
            tot += m
            n += k
        mom2, count = acc.finalize()
        torch.cuda.synchronize()
        row[prec] = float((mom2.double() - tot).norm() / tot.norm())
        row[prec + "_count_ok"] = int(count) == n
        acc.close()
    print(json.dumps(row), flush=True)
    res["mom2"].append(row)


run("one_token", 256, 64, 1)
run("tiny", 256, 64, 300)
run("tiny_masked", 256, 64, 1000, frac_valid=0.6)
run("tiny_gelu", 256, 64, 777, act="gelu", frac_valid=0.8)
run("odd_dims", 200, 80, 500, frac_valid=0.5)
run("clipl_1slab", 3072, 768, 1536)
run("clipl_small_w", 3072, 768, 1536, wscale=0.02)
run("clipl_masked_multi", 3072, 768, 9856, frac_valid=0.55, calls=2)
run("clipl_20calls", 3072, 768, 9856, calls=20)
run("clipl_chunk_2_4", 3072, 768, 9856, calls=4, fc1_chunk=2, syrk_chunk=4)
run("bigg", 5120, 1280, 4096, act="gelu", frac_valid=0.9)

# ---- timing
for (d, h, T, slab, act) in [(3072, 768, 9856, 0, "quick_gelu"), (3072, 768, 12288, 2048, "quick_gelu"),
                             (3072, 768, 12288, 3072, "quick_gelu"), (3072, 768, 16384, 4096, "quick_gelu"),
                             (5120, 1280, 9856, 1024, "gelu")]:
    for prec, kw in [("tf32x3", {}), ("f16x3", {}), ("f16x3", dict(fc1_chunk=2, syrk_chunk=2)),
                     ("f16x3", dict(fc1_chunk=1, syrk_chunk=4))]:
        g = torch.Generator(device=dev).manual_seed(1)
        W = torch.randn(d, h, device=dev, generator=g) * (0.7 / h ** 0.5)
        b = torch.randn(d, device=dev, generator=g) * 0.1
        acc = Mom2Accumulator(dev, d, h, act, slab_tokens=slab, precision=prec, **kw)
        acc.set_weights(W, b)
        X = torch.randn(T, h, device=dev, generator=g)
        for _ in range(3):
            acc.add(X)
        torch.cuda.synchronize()
        acc.profile(True)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        s.record()
        for _ in range(iters):
            acc.add(X)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / iters
        p = acc.get_profile()
        F = 2.0 * h * d + d * (d + 1.0)
        line = dict(d=d, h=h, T=T, slab=slab, prec=prec, ms=ms, tokens_per_s=T / ms * 1e3, algo_tflops=T * F / ms / 1e9,
                    issued_tflops=3 * T * F / ms / 1e9,
                    fc1_tflops=p["fc1_rows"] * 2 * h * d / p["fc1_ms"] / 1e9 if p["fc1_ms"] else None,
                    syrk_tflops=p["syrk_rows"] * d * (d + 1.0) / p["syrk_ms"] / 1e9 if p["syrk_ms"] else None, **kw)
        print(json.dumps(line), flush=True)
        res["timings"].append(line)
        acc.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/probe_f16.json", "w"), indent=1)
print("hang_code", _lib.lib().emcid_hang_code())
print("PROBE DONE")
