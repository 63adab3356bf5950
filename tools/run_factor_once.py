"""One cached-factor solve (for ncu launch lists): python tools/run_factor_once.py [n] [d]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emcid_b200 import solve as S  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
d = int(sys.argv[2]) if len(sys.argv) > 2 else 3072
h = d // 4
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(2 * d, d, device="cuda", generator=g, dtype=torch.float64)
A = A * torch.logspace(0, -3, d, device="cuda", dtype=torch.float64) + 0.2
C = (A.T @ A / (2 * d)).float()
Kt = torch.randn(n, d, device="cuda", generator=g) * 0.5 + 0.2
St = torch.randn(n, h, device="cuda", generator=g)
fac = S.CachedFactor(C, 4000.0)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("factor_solve")
fac.solve(Kt, St, 1.0, 3)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("done")
