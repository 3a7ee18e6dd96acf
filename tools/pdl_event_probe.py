"""Does a cudaEventRecord between two steps cost the programmatic launch edge optimizer -> next per-commit kernel?
Device-resident training steps (hdgnn_train_step), with and without an event record after every step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hdgnn_b200.engine import Engine, DeviceBatch, F_LABEL_BITS
from hdgnn_b200.synthetic import make_commits

B, Ne, Nc = 100, 200, 74
eng = Engine(Ne, Nc, variant=2, max_batch=B, flags=F_LABEL_BITS)
dbs = []
for s in range(8):
    cb = make_commits(B, Ne, Nc, seed=20260 + s)
    dbs.append(DeviceBatch.from_numpy(cb.adj, cb.x, cb.hmap, cb.L, cb.Y, eng.tdev, bits=True))
p = (0.1 * torch.randn(eng.n_params)).cuda(); m = torch.zeros_like(p); v = torch.zeros_like(p)
step = torch.zeros(1, dtype=torch.int32, device="cuda"); loss3 = torch.zeros(3, device="cuda")
evs = [torch.cuda.Event() for _ in range(4)]
side = torch.cuda.Stream()
hsrc = torch.zeros(920000, dtype=torch.uint8).pin_memory(); ddst = torch.zeros(2, 920000, dtype=torch.uint8, device="cuda")
for rec in (0, 1, 2, 3):
    for k in range(50):
        eng.train_step(dbs[k % 8], p, m, v, step, loss3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    N = 1000
    e0.record()
    for k in range(N):
        eng.train_step(dbs[k % 8], p, m, v, step, loss3)
        if rec == 1:
            evs[k % 4].record()
        if rec >= 2:                                   # a concurrent 0.92 MB H2D copy per step on a side stream (2: free-running,
            with torch.cuda.stream(side):              # 3: ordered behind the step before last, as the staging slots are)
                if rec == 3:
                    side.wait_event(evs[(k + 2) % 4])
                ddst[k & 1].copy_(hsrc, non_blocking=True)
            if rec == 3:
                evs[k % 4].record()
    e1.record(); torch.cuda.synchronize()
    print(f"cluster={os.environ.get('HDGNN_CLUSTER', '1')} mode={rec}: {1e3 * e0.elapsed_time(e1) / N:.1f} us/step")
