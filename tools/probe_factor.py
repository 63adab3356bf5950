"""GPU probe of the cached-factor solve (emcid_factor_*) against torch fp64 LU and the direct solver:
accuracy, ms per edit, ms to build the factor.  usage: python tools/probe_factor.py [out.json]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emcid_b200 import solve as S  # noqa: E402


def make_problem(d, h, n, seed, cond_pow=6.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    T = 2 * d
    A = torch.randn(T, d, device="cuda", generator=g, dtype=torch.float64)
    A = A * torch.logspace(0, -cond_pow / 2, d, device="cuda", dtype=torch.float64) + 0.2
    C = (A.T @ A / T).float()
    return C, torch.randn(n, d, device="cuda", generator=g) * 0.5 + 0.2, torch.randn(n, h, device="cuda", generator=g)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


out = []
for d, h, ns in ((3072, 768, (10, 100, 300, 700, 1000)), (5120, 1280, (100, 1000))):
    lam, ew, left = 4000.0, 0.5, 3
    C, _, _ = make_problem(d, h, 1, seed=d)
    fac = S.CachedFactor(C, lam)
    t_create = timed(lambda: S.CachedFactor(C, lam).close(), reps=3)
    for n in ns:
        _, Kt, St = make_problem(d, h, n, seed=d + n)
        Kd = Kt.T.double()
        M = lam * C.double() + Kd @ Kd.T
        adj_ref = torch.linalg.solve(M, Kd)
        dW_ref = (St.T.double() / left) @ adj_ref.T
        adj, resid, dW = fac.solve(Kt, St, 1.0, left)
        adj_d, _, dW_d = S.solve_layers(C, Kt, St, lam, 1.0, [left])
        line = dict(d=d, h=h, n=n, adj_rel=rel(adj, adj_ref), dW_rel=rel(dW, dW_ref), direct_adj_rel=rel(adj_d[0], adj_ref),
                    direct_dW_rel=rel(dW_d[0], dW_ref), factor_create_ms=t_create,
                    cached_ms=timed(lambda: fac.solve(Kt, St, 1.0, left, check=False)),
                    direct_ms=timed(lambda: S.solve_layers(C, Kt, St, lam, 1.0, [left], check=False)),
                    torch_fp64_lu_ms=timed(lambda: torch.linalg.solve(lam * C.double() + Kd @ Kd.T, Kd), reps=3))
        print(json.dumps(line), flush=True)
        out.append(line)
    fac.close()
if len(sys.argv) > 1:
    os.makedirs(os.path.dirname(sys.argv[1]) or ".", exist_ok=True)
    json.dump(out, open(sys.argv[1], "w"), indent=1)
print("PROBE DONE")
