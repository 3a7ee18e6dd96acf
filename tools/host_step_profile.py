"""Host-side cost of one end-to-end training step (graph2graph.train_step from pinned buffers): wall time of the enqueue loop with
the GPU idle (tiny batch: the kernels are shorter than the host work) and a cProfile of where it goes."""
import cProfile, pstats, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hdgnn_b200.model import graph2graph, HostBatch
from hdgnn_b200.synthetic import make_commits

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100
Ne, Nc = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (200, 74)
m = graph2graph(Ne=Ne, Nc=Nc, Mini_batch=B, variant=2, seed=1)
m.initialize()
hbs = [HostBatch(make_commits(B, Ne, Nc, seed=s), bits=m.engine.host_bits) for s in range(4)]
for k in range(50):
    m.train_step(hbs[k % 4])
torch.cuda.synchronize()
N = 3000
t0 = time.perf_counter()
for k in range(N):
    m.train_step(hbs[k % 4])
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"B={B}: enqueue {1e6 * (t1 - t0) / N:.1f} us/step, incl. drain {1e6 * (t2 - t0) / N:.1f} us/step")
pr = cProfile.Profile()
pr.enable()
for k in range(N):
    m.train_step(hbs[k % 4])
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(14)
print(s.getvalue()[:3500])
