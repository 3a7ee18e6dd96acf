"""Device-resident training step time for a synthetic batch with chosen generator settings:
python tools/step_probe.py [p_short] [p_edge] [x_max] [Ne] [Nc] [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hdgnn_b200.engine import Engine, DeviceBatch, F_LABEL_BITS
from hdgnn_b200.synthetic import make_commits

a = sys.argv[1:]
p_short = float(a[0]) if len(a) > 0 else 0.2
p_edge = float(a[1]) if len(a) > 1 else 0.05
x_max = int(a[2]) if len(a) > 2 else 9
Ne, Nc, B = (int(v) for v in (a[3:6] if len(a) > 5 else (200, 74, 100)))
eng = Engine(Ne, Nc, variant=2, max_batch=B, flags=F_LABEL_BITS)
pool = []
for i in range(40):
    cb = make_commits(B, Ne, Nc, seed=100 + i, p_short=p_short, p_edge=p_edge, x_max=x_max)
    pool.append(DeviceBatch.from_numpy(cb.adj, cb.x, cb.hmap, cb.L, cb.Y, eng.tdev, bits=True))
p = (0.1 * torch.randn(eng.n_params)).cuda(); m = torch.zeros_like(p); v = torch.zeros_like(p)
step = torch.zeros(1, dtype=torch.int32, device="cuda"); loss3 = torch.zeros(3, device="cuda")
for k in range(20):
    eng.train_step(pool[k % 40], p, m, v, step, loss3)
torch.cuda.synchronize()
ts = []
for rep in range(7):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(200):
        eng.train_step(pool[k % 40], p, m, v, step, loss3)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 200 * 1e3)
print(f"p_short={p_short} p_edge={p_edge} x_max={x_max} Ne={Ne} Nc={Nc} B={B}: {np.median(ts):.1f} us/step  ({B / np.median(ts) * 1e6:.0f} commits/s), launches {eng.last_launch_count()}")
