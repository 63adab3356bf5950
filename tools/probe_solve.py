"""GPU probe of the closed-form update against torch fp64 (the reference's own arithmetic)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emcid_b200 import solve as S  # noqa: E402

out = []


def make_problem(B, d, h, n, seed, cond_pow=6.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    Cs, Ks, Ss = [], [], []
    for b in range(B):
        # covariance-like SPD matrix: mean outer product + decaying spectrum
        T = 2 * d
        A = torch.randn(T, d, device="cuda", generator=g, dtype=torch.float64)
        scale = torch.logspace(0, -cond_pow / 2, d, device="cuda", dtype=torch.float64)
        A = A * scale + 0.2
        C = (A.T @ A / T).float()
        Cs.append(C)
        Ks.append((torch.randn(n, d, device="cuda", generator=g) * 0.5 + 0.2))
        Ss.append(torch.randn(n, h, device="cuda", generator=g))
    return torch.stack(Cs), torch.stack(Ks), torch.stack(Ss)


def reference(C32, Kt, St, lam, ew, left):
    s = (ew / 0.5) ** 0.5
    res = []
    for b in range(C32.shape[0]):
        Kd = Kt[b].T.double() * s
        Sd = St[b].T.double() * s
        M = lam * C32[b].double() + Kd @ Kd.T
        adj = torch.linalg.solve(M, Kd)
        resid = Sd / left[b]
        res.append((adj, resid, (resid @ adj.T), torch.linalg.cond(M).item() if C32.shape[1] <= 1024 else float("nan")))
    return res


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def run(name, B, d, h, n, lam=4000.0, ew=0.5, refine=2, seed=0, time_it=False):
    C32, Kt, St = make_problem(B, d, h, n, seed)
    left = list(range(B, 0, -1))
    ref = reference(C32, Kt, St, lam, ew, left)
    s = (ew / 0.5) ** 0.5
    adj, resid, dW = S.solve_layers(C32, Kt, St, lam, s, left, refine_steps=refine)
    torch.cuda.synchronize()
    line = dict(name=name, B=B, d=d, h=h, n=n, lam=lam, ew=ew, refine=refine,
                adj_rel=max(rel(adj[b], ref[b][0]) for b in range(B)),
                resid_rel=max(rel(resid[b], ref[b][1]) for b in range(B)),
                dW_rel=max(rel(dW[b], ref[b][2]) for b in range(B)), cond=ref[0][3])
    if time_it:
        for _ in range(2):
            S.solve_layers(C32, Kt, St, lam, s, left, refine_steps=refine, check=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            S.solve_layers(C32, Kt, St, lam, s, left, refine_steps=refine, check=False)
        e1.record()
        torch.cuda.synchronize()
        line["ms"] = e0.elapsed_time(e1) / 3
        e0.record()
        reference(C32, Kt, St, lam, ew, left)
        e1.record()
        torch.cuda.synchronize()
        line["torch_fp64_ms"] = e0.elapsed_time(e1)
    print(json.dumps(line), flush=True)
    out.append(line)


run("tiny", 1, 256, 64, 10)
run("tiny_b2_ew06", 2, 256, 64, 10, lam=10000.0, ew=0.6)
run("mid", 1, 1024, 256, 200)
run("mid_r0", 1, 1024, 256, 200, refine=0)
run("mid_r1", 1, 1024, 256, 200, refine=1)
run("clipl_1", 1, 3072, 768, 1000, time_it=True)
run("clipl_1_r1", 1, 3072, 768, 1000, refine=1, time_it=True)
run("clipl_5", 5, 3072, 768, 1000, time_it=True)
run("clipl_5_n100", 5, 3072, 768, 100, time_it=True)
run("bigg_1", 1, 5120, 1280, 1000, time_it=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_solve.json", "w"), indent=1)
print("PROBE DONE")
