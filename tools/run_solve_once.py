"""One batched closed-form update (B layers, n concepts, CLIP-L sizes) for an ncu launch list / timing."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emcid_b200 import solve as S
B = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
d, h = 3072, 768
g = torch.Generator(device="cuda").manual_seed(0)
Cs = []
for b in range(B):
    T = 2 * d
    A = torch.randn(T, d, device="cuda", generator=g, dtype=torch.float64)
    A = A * torch.logspace(0, -3.0, d, device="cuda", dtype=torch.float64) + 0.2
    Cs.append((A.T @ A / T).float())
C32 = torch.stack(Cs)
Kt = torch.randn(B, n, d, device="cuda", generator=g) * 0.5 + 0.2
St = torch.randn(B, n, h, device="cuda", generator=g)
left = list(range(B, 0, -1))
S.solve_layers(C32, Kt, St, 4000.0, 1.0, left)   # warm-up (workspace, attributes)
torch.cuda.synchronize()
import time
per = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    adj, resid, dW = S.solve_layers(C32, Kt, St, 4000.0, 1.0, left, check=False)
    e1.record()
    torch.cuda.synchronize()
    per.append((round(e0.elapsed_time(e1), 2), round(1e3 * (time.perf_counter() - t0), 2)))
print(json.dumps({"B": B, "n": n, "ms": sum(p[0] for p in per) / reps, "per_rep_device_host_ms": per}))
