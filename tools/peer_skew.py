"""Where the multi-GPU step loses time: per rank, the wait between pushing its gradient slice and holding every peer's
(reduce_adam_kernel, CTA 0; %globaltimer).  The rank that finishes its backward LAST waits only for the NVLink round trip,
the others additionally for the slowest rank (skew).  Run under torchrun with HDGNN_PEER_STAMPS=1:
  HDGNN_PEER_STAMPS=1 python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/peer_skew.py > out.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from hdgnn_b200.model import graph2graph
from hdgnn_b200.synthetic import make_commits

os.environ["HDGNN_PEER_STAMPS"] = "1"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
Ne, Nc, B, steps = 200, 74, 100, 600
model = graph2graph(None, Ne=Ne, Nc=Nc, Mini_batch=B, variant=2, device=local, seed=1, max_batch=B, collective="peer")
pool = [model.host_batch(make_commits(B, Ne, Nc, seed=500 + 100 * rank + i)) for i in range(8)]
for k in range(steps):
    model.train_step(pool[k % 8])
torch.cuda.synchronize()
st = model.engine.workspace("PEER_STAMPS", (4096, 2), dtype=torch.int64).cpu().numpy()
wait = (st[100:steps, 1] - st[100:steps, 0]).astype(np.float64) / 1e3          # us, steps 100.. (sequence numbers start at 1)
allw = [None] * world
dist.all_gather_object(allw, wait.tolist())
if rank == 0:
    W = np.array(allw)                      # (ranks, steps)
    lo, hi = W.min(0), W.max(0)
    print(json.dumps({"ranks": world, "steps": int(W.shape[1]), "unit": "us",
                      "wait_of_last_rank_median": float(np.median(lo)), "wait_of_first_rank_median": float(np.median(hi)),
                      "skew_median": float(np.median(hi - lo)), "per_rank_median": [float(np.median(w)) for w in W],
                      "what": "per step: min over ranks of (last pull - push) = NVLink store + poll latency seen by the rank that finished last; "
                              "max over ranks = the same plus the spread of the ranks' backward finish times"}))
dist.destroy_process_group()
