"""Per-phase clock64 breakdown of mid2_kernel (debug handle): mean / max cycles per phase over the commits."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hdgnn_b200.engine import Engine, DeviceBatch
from hdgnn_b200.synthetic import make_commits

Ne, Nc, B = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (200, 74, 100)))
cb = make_commits(B, Ne, Nc, seed=20260)
eng = Engine(Ne, Nc, variant=2, max_batch=B, flags=2)
db = DeviceBatch.from_numpy(cb.adj, cb.x, cb.hmap, cb.L, cb.Y, eng.tdev)
params = (0.1 * torch.randn(eng.n_params)).cuda()
for _ in range(3):
    eng.forward_backward(db, params)
torch.cuda.synchronize()
clk = eng.workspace("CLK", (B, 16), dtype=torch.int64).cpu().numpy()
names = ["A load", "B node fwd", "D pool fwd", "tables+E hunk fwd", "F head tables", "G1 head", "G2 delta sums", "H head bwd",
         "I hunk bwd", "J+K pool bwd", "L node bwd"]
d = np.diff(clk[:, :12], axis=1)
ident = cb.L == Ne
print(f"commits {B}, L==Ne for {int(ident.sum())}")
for i, n in enumerate(names):
    print(f"{n:20s} mean {d[:, i].mean():9.0f}  max {d[:, i].max():9.0f}   ident {d[ident, i].mean():9.0f}  general {d[~ident, i].mean() if (~ident).any() else 0:9.0f}")
tot = clk[:, 11] - clk[:, 0]
print(f"total mean {tot.mean():.0f} max {tot.max():.0f} cycles")
