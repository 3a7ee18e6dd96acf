"""Worker of tests/test_gpu_peer.py: one process per GPU (torchrun env).  Trains a few steps on this rank's shard with the
gradient exchange fused into the last kernel (peer memory) and with the NCCL path, and writes what the test compares."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from hdgnn_b200.engine import Engine, DeviceBatch          # noqa: E402
from hdgnn_b200.model import graph2graph, HostBatch         # noqa: E402
from hdgnn_b200.synthetic import make_commits               # noqa: E402


def main():
    out_dir, variant, Ne, Nc, per, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    Bg = per * world
    res = {}
    for mode in ("peer", "nccl"):
        model = graph2graph(None, Ne=Ne, Nc=Nc, Mini_batch=per, variant=variant, device=local, seed=5, max_batch=Bg, collective=mode)
        assert model.peer == (mode == "peer")
        ces = []
        for t in range(steps):
            cb = make_commits(Bg, Ne, Nc, seed=100 + t)           # the global batch, generated identically on every rank
            hb = model.host_batch(cb.slice(rank * per, (rank + 1) * per))
            l3 = model.train_step(hb)
            torch.cuda.synchronize()
            ce = float(l3[0])
            if mode == "nccl":                                     # shares add up to the global mean
                tt = torch.tensor([ce], device=model.engine.tdev, dtype=torch.float64)
                dist.all_reduce(tt)
                ce = float(tt.item())
            ces.append(ce)
        res[mode + "_params"] = model.params.cpu().numpy()
        res[mode + "_m"] = model.m.cpu().numpy()
        res[mode + "_ce"] = np.array(ces)
        res[mode + "_launches"] = np.array(model.engine.last_launch_count())
        res[mode + "_reg"] = l3[1:3].clone().numpy()
        model.engine.close()
    if rank == 0:                                                  # the same global batches on ONE GPU
        torch.cuda.set_device(local)
        eng = Engine(Ne, Nc, variant=variant, max_batch=Bg, device=local)
        from hdgnn_b200.model import truncated_normal_init
        params = truncated_normal_init(variant, 5).cuda()
        m = torch.zeros_like(params); v = torch.zeros_like(params)
        step = torch.zeros(1, dtype=torch.int32, device="cuda"); loss3 = torch.zeros(3, device="cuda")
        ces = []
        for t in range(steps):
            cb = make_commits(Bg, Ne, Nc, seed=100 + t)
            db = DeviceBatch.from_numpy(cb.adj, cb.x, cb.hmap, cb.L, cb.Y, eng.tdev)
            eng.train_step(db, params, m, v, step, loss3)
            torch.cuda.synchronize()
            ces.append(float(loss3[0]))
        res["single_params"] = params.cpu().numpy(); res["single_ce"] = np.array(ces)
        eng.close()
    # graph2graph.train() sharded over the ranks (peer exchange, device evaluation counters) vs the same loop on one GPU
    if Ne <= 64:
        from hdgnn_b200.utils2 import write_dataset, read_compact, split_half
        from hdgnn_b200.engine import eval_counts
        N_all, MB, epochs = 8 * per * world, per * world, 2
        if rank == 0:
            write_dataset(make_commits(N_all, Ne, Nc, seed=77, p_edge=0.2, p_short=0.4, p_noise=0.2), "toy", 2, root=out_dir)
        dist.barrier()
        model = graph2graph(None, Ne=Ne, Nc=Nc, Mini_batch=MB, epoch=epochs, Step=2, Repo="toy", variant=variant, device=local,
                            seed=9, checkpoint_dir=os.path.join(out_dir, "ck"), collective="peer")
        hist = model.train(None, root=out_dir, log=lambda *_: None)
        res["train_loss"] = np.array([h["hedge_loss"] for h in hist]); res["train_acc"] = np.array([h["acc"] for h in hist])
        res["train_params"] = model.params.cpu().numpy()
        model.engine.close()
        if rank == 0:
            from hdgnn_b200.model import truncated_normal_init
            train, _ = split_half(read_compact("toy", 2, Ne, Nc, root=out_dir))
            eng = Engine(Ne, Nc, variant=variant, max_batch=MB, device=local)
            params = truncated_normal_init(variant, 9).cuda()
            m = torch.zeros_like(params); v = torch.zeros_like(params)
            step = torch.zeros(1, dtype=torch.int32, device="cuda"); loss3 = torch.zeros(3, device="cuda")
            losses, accs = [], []
            for ep in range(epochs):
                ls, hits = [], 0
                for j in range(train.B // MB):
                    b = train.slice(j * MB, (j + 1) * MB)
                    db = DeviceBatch.from_numpy(b.adj, b.x, train.hmap[:MB], train.L[:MB], b.Y, eng.tdev)      # quirk Q2
                    probs = torch.zeros(MB, 2, Nc * (Nc - 1), device="cuda")
                    eng.train_step(db, params, m, v, step, loss3, probs=probs)
                    c, _ = eval_counts(probs, torch.as_tensor(b.Y).cuda())
                    torch.cuda.synchronize()
                    ls.append(float(loss3[0])); hits += int(c[:, 0].sum())
                losses.append(np.mean(ls)); accs.append(hits / ((train.B // MB) * MB * Nc * (Nc - 1)))
            res["single_train_loss"] = np.array(losses); res["single_train_acc"] = np.array(accs)
            res["single_train_params"] = params.cpu().numpy()
            eng.close()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
