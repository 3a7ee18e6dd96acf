"""Eager launches vs CUDA-graph replay of the 4-launch training step (same batch), glide shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hdgnn_b200.engine import Engine, DeviceBatch, F_LABEL_BITS
from hdgnn_b200.model import truncated_normal_init
from hdgnn_b200.synthetic import make_commits
Ne, Nc, B = 200, 74, 100
eng = Engine(Ne, Nc, variant=2, max_batch=B, flags=F_LABEL_BITS)
pool = []
for i in range(8):
    cb = make_commits(B, Ne, Nc, seed=20260 + i)
    pool.append(DeviceBatch.from_numpy(cb.adj, cb.x, cb.hmap, cb.L, cb.Y, eng.tdev, bits=True))
params = truncated_normal_init(2, 3).cuda(); m = torch.zeros_like(params); v = torch.zeros_like(params)
step = torch.zeros(1, dtype=torch.int32, device="cuda"); loss3 = torch.zeros(3, device="cuda")
probs = torch.zeros(B, 2, Nc * (Nc - 1), device="cuda")
def run(fn, n=400):
    for k in range(20): fn(k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for k in range(n): fn(k)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
eager = run(lambda k: eng.train_step(pool[k % 8], params, m, v, step, loss3, probs=probs))
graphs = []
for db in pool:
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        eng.train_step(db, params, m, v, step, loss3, probs=probs)
    graphs.append(g)
# 8 steps per graph as well
g8 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g8):
    for db in pool:
        eng.train_step(db, params, m, v, step, loss3, probs=probs)
rep = run(lambda k: graphs[k % 8].replay())
rep8 = run(lambda k: g8.replay(), n=50) / 8
print(f"eager {eager:.1f} us/step, graph replay {rep:.1f} us/step, 8-step graph {rep8:.1f} us/step")
