"""Small training steps through every branch of the fused kernel, for compute-sanitizer (memcheck / racecheck):
  compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hdgnn_b200.engine import Engine, DeviceBatch, normalize_propagate, normalize_propagate_backward, map_conv_backward
from hdgnn_b200.synthetic import make_commits

rng = np.random.default_rng(0)
for variant in (2, 4, 1):
    for attr in ("classes", "continuous"):
        for (B, Ne, Nc) in ((3, 40, 12), (2, 70, 33)):
            cb = make_commits(B, Ne, Nc, seed=3 + Ne, p_edge=0.15, p_short=0.5)
            cb.L[0] = Ne
            if B > 1:
                cb.L[1] = Ne - 5
            if attr == "continuous":
                cb.x[:] = rng.normal(size=cb.x.shape)
            eng = Engine(Ne, Nc, variant=variant, max_batch=B)
            db = DeviceBatch.from_numpy(cb.adj, cb.x, cb.hmap, cb.L, cb.Y, eng.tdev)
            p = (0.1 * torch.randn(eng.n_params)).cuda(); m = torch.zeros_like(p); v = torch.zeros_like(p)
            step = torch.zeros(1, dtype=torch.int32, device="cuda"); loss3 = torch.zeros(3, device="cuda")
            probs = torch.zeros(B, 2, eng.Ncr, device="cuda")
            for _ in range(2):
                eng.train_step(db, p, m, v, step, loss3, probs=probs)
            eng.forward(db, p)
            torch.cuda.synchronize()
            assert torch.isfinite(p).all() and torch.isfinite(probs).all()
            print("ok", variant, attr, B, Ne, Nc, eng.last_launch_count(), float(loss3[0]))
            eng.close()
adj = (torch.rand(3, 50, 64, device="cuda") < 0.1).to(torch.uint8); adj[:, :, 50:] = 0
H = torch.rand(3, 50, 7, device="cuda"); W = torch.rand(7, 20, device="cuda")
out, _ = normalize_propagate(adj, H, W, flags=2)
normalize_propagate_backward(adj, H, torch.rand(3, 50, 20, device="cuda"), W=W, out=out, flags=2)
map_conv_backward(adj, torch.rand(3, 50, device="cuda"), torch.tensor([0.1, -0.2], device="cuda"))
torch.cuda.synchronize()
print("legacy ok")
