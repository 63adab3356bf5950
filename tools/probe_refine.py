"""Prints the refinement sweeps of the cached-factor and the direct solve (EMCID_SOLVE_DEBUG=1):
python tools/probe_refine.py [d] [n]"""
import os
import sys

os.environ["EMCID_SOLVE_DEBUG"] = "1"
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emcid_b200 import solve as S  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 5120
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
h = d // 4
g = torch.Generator(device="cuda").manual_seed(d)
A = torch.randn(2 * d, d, device="cuda", generator=g, dtype=torch.float64)
A = A * torch.logspace(0, -3, d, device="cuda", dtype=torch.float64) + 0.2
C = (A.T @ A / (2 * d)).float()
g = torch.Generator(device="cuda").manual_seed(d + n)
Kt = torch.randn(n, d, device="cuda", generator=g) * 0.5 + 0.2
St = torch.randn(n, h, device="cuda", generator=g)
Kd = Kt.T.double()
ref = torch.linalg.solve(4000.0 * C.double() + Kd @ Kd.T, Kd)
print("== cached", file=sys.stderr)
fac = S.CachedFactor(C, 4000.0)
adj, _, _ = fac.solve(Kt, St, 1.0, 3)
print("cached adj_rel", float((adj - ref).norm() / ref.norm()), file=sys.stderr)
print("== direct", file=sys.stderr)
adj, _, _ = S.solve_layers(C, Kt, St, 4000.0, 1.0, [3])
print("direct adj_rel", float((adj[0] - ref).norm() / ref.norm()), file=sys.stderr)
