"""GPU probe for the 3xTF32 tcgen05 GEMM: correctness ladder, accumulation-rounding behaviour and
first timings.  Run on a B200 via gpurun; prints one line per check and writes gpurun_out/probe_gemm.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emcid_b200 import _lib  # noqa: E402

out = {"checks": []}


def rel(a, b):
    return float((a.double() - b).norm() / b.norm())


def check(name, M, N, K, **kw):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g)
    ref = A.double() @ B.double().T
    C0 = None
    alpha, beta = kw.pop("alpha", 1.0), kw.pop("beta", 0.0)
    if beta != 0.0:
        C0 = torch.randn(M, N, device="cuda", generator=g)
        ref = alpha * ref + beta * C0.double()
        C0 = C0.clone()
    else:
        ref = alpha * ref
    C = _lib.gemm3x_nt(A, B, C0, alpha=alpha, beta=beta, **kw)
    torch.cuda.synchronize()
    if kw.get("lower"):
        mask = torch.tril(torch.ones(M, N, device="cuda", dtype=torch.bool))
        e = float(((C.double() - ref) * mask).norm() / (ref * mask).norm())
    else:
        e = rel(C, ref)
    fp32 = rel(A @ B.T, A.double() @ B.double().T)
    line = {"name": name, "M": M, "N": N, "K": K, "rel_err": e, "torch_fp32_rel_err": fp32,
            **{k: str(v) for k, v in kw.items()}}
    print(json.dumps(line), flush=True)
    out["checks"].append(line)
    return e


torch.backends.cuda.matmul.allow_tf32 = False
print("device", torch.cuda.get_device_name(0), flush=True)
_lib.check(_lib.lib().emcid_device_check(0))

check("one_tile_one_kblock", 128, 256, 32)
check("one_tile_k128", 128, 256, 128)
check("one_tile_n128", 128, 128, 64, n128=True)
check("multi_tile", 512, 768, 256)
check("ragged", 200, 260, 72)
check("ragged_n128", 130, 132, 40, n128=True)
check("beta", 256, 512, 96, alpha=0.5, beta=2.0)
check("big", 3072, 2048, 768)
check("lower", 1024, 1024, 512, lower=True)
check("lower_streamk", 1024, 1024, 2048, lower=True, streamk=True, alpha=1.0, beta=1.0)
check("streamk_full", 640, 512, 4096, streamk=True, alpha=1.0, beta=1.0)

# ---- accumulation rounding probe: positive terms, growing K.  Linear growth of the relative
# error with K means round-toward-zero accumulation in the tensor core; sqrt growth means RN.
probe = []
for chunk in (1, 2, 4, 8, 32):
    for K in (256, 1024, 4096, 16384):
        g = torch.Generator(device="cuda").manual_seed(K)
        A = torch.rand(128, K, device="cuda", generator=g) + 0.5
        B = torch.rand(256, K, device="cuda", generator=g) + 0.5
        ref = A.double() @ B.double().T
        C = _lib.gemm3x_nt(A, B, chunk=chunk)
        torch.cuda.synchronize()
        d = (C.double() - ref) / ref
        A2 = torch.randn(128, K, device="cuda", generator=g)
        B2 = torch.randn(256, K, device="cuda", generator=g)
        ref2 = A2.double() @ B2.double().T
        C2 = _lib.gemm3x_nt(A2, B2, chunk=chunk)
        probe.append({"chunk_kblocks": chunk, "K": K, "pos_mean_rel": float(d.mean()),
                      "pos_rms_rel": float(d.pow(2).mean().sqrt()),
                      "randn_rel_fro": rel(C2, ref2), "torch_fp32_randn_rel_fro": rel(A2 @ B2.T, ref2),
                      "torch_fp32_pos_mean_rel": float((((A @ B.T).double() - ref) / ref).mean())})
        print(json.dumps(probe[-1]), flush=True)
out["rounding_probe"] = probe


# ---- timings
def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


timings = []
lib = _lib.lib()
for (M, N, K, kw) in [
    (3072, 3072, 1536, dict(lower=True, streamk=True, alpha=1.0, beta=1.0, chunk=1)),
    (3072, 3072, 1536, dict(lower=True, streamk=True, alpha=1.0, beta=1.0, chunk=2)),
    (3072, 3072, 1536, dict(lower=True, streamk=True, alpha=1.0, beta=1.0, chunk=4)),
    (3072, 3072, 1536, dict(lower=True, streamk=True, alpha=1.0, beta=1.0, chunk=48)),
    (3072, 3072, 1536, dict(lower=True)),
    (3072, 3072, 1536, dict()),
    (3072, 1536, 768, dict()),
    (3072, 3072, 4096, dict(lower=True, streamk=True, alpha=1.0, beta=1.0)),
    (8192, 8192, 8192, dict(chunk=1)),
    (8192, 8192, 8192, dict(chunk=2)),
    (8192, 8192, 8192, dict(chunk=4)),
    (8192, 8192, 8192, dict(chunk=255)),
]:
    A = torch.randn(M, K, device="cuda")
    B = A if (M == N) else torch.randn(N, K, device="cuda")
    C = torch.zeros(M, N, device="cuda")
    ws_bytes = lib.emcid_gemm3x_workspace_bytes(M, N, K)
    ws = torch.empty(ws_bytes, device="cuda", dtype=torch.uint8)
    flags = (1 if kw.get("lower") else 0) | (2 if kw.get("streamk") else 0) | (kw.get("chunk", 0) << 8)
    alpha, beta = kw.get("alpha", 1.0), kw.get("beta", 0.0)

    def run():
        _lib.check(lib.emcid_gemm3x_nt(M, N, K, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), C.data_ptr(),
                                       C.stride(0), alpha, beta, flags, ws.data_ptr(), ws_bytes,
                                       torch.cuda.current_stream().cuda_stream))

    ms = timeit(run)
    flops = 2.0 * M * N * K * (0.5 if kw.get("lower") else 1.0)
    torch.backends.cuda.matmul.allow_tf32 = True
    ms_tf32 = timeit(lambda: torch.matmul(A, B.T))
    torch.backends.cuda.matmul.allow_tf32 = False
    ms_fp32 = timeit(lambda: torch.matmul(A, B.T), iters=3)
    line = {"M": M, "N": N, "K": K, **{k: str(v) for k, v in kw.items()}, "ms_incl_split": ms,
            "algo_tflops": flops / ms / 1e9, "issued_tflops": 3 * flops / ms / 1e9,
            "torch_tf32_ms": ms_tf32, "torch_tf32_tflops": 2.0 * M * N * K / ms_tf32 / 1e9,
            "torch_fp32_ms": ms_fp32}
    print(json.dumps(line), flush=True)
    timings.append(line)
out["timings"] = timings
out["hang_code"] = int(lib.emcid_hang_code())
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_gemm.json", "w"), indent=1)
print("PROBE DONE", flush=True)
