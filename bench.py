#!/usr/bin/env python
"""bench.py -- training commits/s of the HD-GNN hot path on B200 (BASELINE.json metric).

A step = forward + backward + regularisers + TF-Adam over one batch of B synthetic commits per
GPU (model_2.py:369-383).  `value` times the step with inputs resident in HBM; `e2e` times the
public call `graph2graph.train_step` from pinned HOST buffers (H2D of the five compact inputs
and D2H of the loss inside the timed region).  `--impl reference` times the CPU restatement of
the reference (oracle/) on the host cores instead.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload glide|cfg2|cfg3|cfg4] [--variant 2]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {   # BASELINE.json configs; B = commits per GPU per step
    "glide": dict(Ne=200, Nc=74, B=100, step=2, desc="glide-shaped synthetic, Ne=200 Nc=74 Ner=39800 Ncr=5402"),
    "cfg2": dict(Ne=250, Nc=114, B=100, step=3, desc="synthetic Ne=250 Nc=114 Ner=62250 Ncr=12882 Step=3"),
    "cfg3": dict(Ne=250, Nc=150, B=100, step=5, desc="synthetic Ne=250 Nc=150 Ner=62250 Ncr=22350 Step=5"),
    "cfg4": dict(Ne=512, Nc=256, B=512, step=0, desc="synthetic scale-up Ne=512 Nc=256, 512 commits/GPU/step"),
}
L2_BYTES = 126 * 2 ** 20
METRIC = "train commits/sec (fwd+bwd+Adam)"


def flops_fwd_per_commit(Ne, Nc, variant):
    """Canonical algorithmic FLOPs of the closed form, no one-hot matmuls, no collapse."""
    ner, ncr = Ne * (Ne - 1), Nc * (Nc - 1)
    f = ncr * 2230 + ner * 8
    if variant in (2, 4):
        f += ner * 1000 + Ne * 880
    if variant == 4:
        f += ner * 1986
    return f


def bytes_per_commit_train(Ne, Nc):
    """Compulsory HBM bytes of one training step per commit in the u8 encoding: inputs read by
    forward and again by backward (recompute) + probs written."""
    inp = Ne * Ne + 4 * Ne + 4 * Ne + 4 + Nc * Nc
    return 2 * inp + 8 * Nc * (Nc - 1)


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1]))
            except ValueError:
                continue
            for n, v in zip(names, c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_rate(Ne, Nc, variant, sample_commits, steps, warmup):
    """fwd + bwd (autograd) + TF-Adam of the closed-form CPU restatement, fp32, all host threads."""
    import torch
    from hdgnn_b200.synthetic import make_commits
    from oracle import hdgnn_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cb = make_commits(sample_commits, Ne, Nc, seed=20260)
    flat = O.init_params(variant, seed=1234, dtype=torch.float32)
    m = torch.zeros_like(flat); v = torch.zeros_like(flat)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        _, _, _, g, _ = O.train_loss_and_grad(variant, flat, cb.adj, cb.x, cb.hmap, cb.L, cb.Y)
        flat, m, v = O.tf_adam_step(flat, g, m, v, it + 1)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return sample_commits * len(times) / sum(times), cores, sum(times) / len(times)


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.ref_sample
    rate, cores, sec = cpu_oracle_rate(wl["Ne"], wl["Nc"], args.variant, sample, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "commits/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "variant": args.variant, "commits_per_step": sample},
        "cpu_baseline": {"value": rate, "unit": "commits/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} commits/step x {args.steps} steps of the same workload; closed-form "
                                   "restatement of model_2.py:86-118 + autograd + TF-Adam in PyTorch-CPU fp32 "
                                   "(TensorFlow is not installable here, so the reference itself cannot run)"},
        "e2e": {"value": rate, "unit": "commits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained"), "measured"
    return 6650.0, 1590.0, None, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="glide", choices=sorted(WORKLOADS))
    ap.add_argument("--variant", type=int, default=2)
    ap.add_argument("--batch", type=int, default=0, help="commits per GPU per step (default: workload's)")
    ap.add_argument("--ref-sample", type=int, default=25, help="commits per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--collective", default="auto", choices=["auto", "peer", "nccl"],
                    help="N>1 gradient exchange: fused into the last kernel over NVLink peer memory, or NCCL all-reduce")
    ap.add_argument("--rows-e", type=int, default=0)
    ap.add_argument("--rows-c", type=int, default=0)
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl["B"] = args.batch
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        # K and W are honoured; the per-step sample shrinks if K steps would not finish in ~3 minutes
        # (the oracle does ~100 commits/s on 16 cores at glide)
        while args.ref_sample > 5 and (args.steps + args.warmup) * args.ref_sample / 80.0 > 180.0:
            args.ref_sample = max(5, args.ref_sample // 2)
        return run_reference(args, wl)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"       # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from hdgnn_b200.engine import DeviceBatch
    from hdgnn_b200.model import graph2graph, HostBatch
    from hdgnn_b200.synthetic import make_commits

    Ne, Nc, B, variant = wl["Ne"], wl["Nc"], wl["B"], args.variant
    model = graph2graph(None, Ne=Ne, Nc=Nc, Mini_batch=B, Step=wl["step"], Repo="synthetic", variant=variant,
                        device=local, seed=1234, max_batch=B, collective=args.collective)
    eng = model.engine
    dev = eng.tdev
    # pool of distinct batches, larger than L2, generated globally (seed per pool slot and rank)
    batch_bytes = B * (Ne * eng.pe + Nc * eng.pc + 8 * Ne + 4)
    pool_n = max(4, int(np.ceil(1.25 * L2_BYTES / batch_bytes)) + 1)
    pool_n = min(pool_n, 256)
    n_distinct = min(pool_n, 32)        # distinct synthetic batches; further pool slots are copies at distinct addresses
    host_pool, dev_pool = [], []
    for i in range(pool_n):
        if i < n_distinct:
            cb = make_commits(B, Ne, Nc, seed=20260 + 1000 * rank + i)
        hb = model.host_batch(cb) if i < n_distinct else host_pool[i % n_distinct].clone()
        host_pool.append(hb)
        if i < n_distinct:
            dev_pool.append(DeviceBatch.from_numpy(cb.adj, cb.x, cb.hmap, cb.L, cb.Y, dev, bits=model.host_bits))
        else:
            d0 = dev_pool[i % n_distinct]
            dev_pool.append(DeviceBatch(d0.adj.clone(), d0.x.clone(), d0.hmap.clone(), d0.L.clone(), d0.Y.clone(), Ne, Nc))
    Bg = B * world
    probs = torch.empty(B, 2, eng.Ncr, dtype=torch.float32, device=dev)
    loss = torch.zeros(1, dtype=torch.float32, device=dev)

    loss3 = torch.zeros(3, dtype=torch.float32, device=dev)

    def device_step(k):
        db = dev_pool[k % pool_n]
        if world == 1:      # one fused call: forward, backward, gradient reduction + regularisers + Adam
            eng.train_step(db, model.params, model.m, model.v, model.step_counter, loss3, probs=probs)
            return eng.last_launch_count()
        if model.peer:      # same 5 launches: the all-reduce is fused into the reduce + Adam kernel (peer memory over NVLink)
            eng.train_step_peer(db, model.params, model.m, model.v, model.step_counter, loss3, probs=probs)
            return eng.last_launch_count()
        eng.forward_backward(db, model.params, B_global=Bg, grads=model.grads, probs=probs, loss=loss)
        n = eng.last_launch_count()
        dist.all_reduce(model.grads)
        eng.adam_step(model.params, model.grads, model.m, model.v, model.step_counter, reg_losses=model.reg)
        return n + eng.last_launch_count()

    def host_step(k):
        model.train_step(host_pool[k % pool_n])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sample_clocks=False):
        for k in range(warmup):
            fn(k)
        barrier()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
            time.sleep(0.15)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        launches = 0
        for k in range(steps):
            r = fn(warmup + k)
            launches += r or 0
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches, clocks

    ms_dev, launches, clocks = timed(device_step, args.steps, args.warmup, sample_clocks=True)
    model.initialize(model.params.clone())      # reset Adam state, keep weights
    ms_e2e, _, _ = timed(host_step, args.steps, args.warmup)
    value = Bg * args.steps / (ms_dev * 1e-3)
    e2e_value = Bg * args.steps / (ms_e2e * 1e-3)

    # per-kernel timing pass (events around every launch; perturbs the step, so never the headline)
    eng.profile(True)
    prof_steps = min(args.steps, 20)
    for k in range(prof_steps):
        device_step(k)
    torch.cuda.synchronize()
    recs = eng.profile_records()
    eng.profile(False)
    agg = {}
    for name, ms in recs:
        a = agg.setdefault(name, [0.0, 0])
        a[0] += ms; a[1] += 1
    tot = sum(a[0] for a in agg.values())
    kernels = {n: {"avg_us": 1e3 * a[0] / a[1], "share": a[0] / tot} for n, a in agg.items()}
    top = max(agg, key=lambda n: agg[n][0])
    top_us = kernels[top]["avg_us"]
    ner, ncr = Ne * (Ne - 1), Nc * (Nc - 1)
    alg = {   # algorithmic FLOPs per launch of each pair kernel (SURVEY 8(d) per-pair figures x pairs x B)
        "pairsum_fwd(ent)": B * ner * 1000, "pairsum_bwd(ent)": B * ner * 2000,
        "pairsum_fwd(edge)": B * ner * 1020, "pairsum_bwd(edge)": B * ner * 2040,
        "score_fwd(edge)": B * ner * 966, "score_bwd(edge)": B * ner * 3 * 966,
        "pairsum_fwd(hunk)": B * ncr * 1260, "pairsum_bwd(hunk)": B * ncr * 2520,
        "score(hunk)": B * ncr * 970 * 3,
        # fused path: canonical (un-collapsed) FLOPs of the reference ops each kernel covers
        "ent_fwd": B * ner * 1000, "ent_bwd": B * ner * 2000,
        "mid(train)": B * 3 * (ncr * 2230 + ner * 8 + Ne * 880), "mid(infer)": B * (ncr * 2230 + ner * 8 + Ne * 880),
    }
    hbm_peak, bf16_peak, bf16_sus, src = measured_peaks()
    # The contract's two bounds are quoted for the dominant kernel against the measured peaks; neither
    # binds this path (K = 20 contractions collapse algebraically): the bound that does is
    # the SM issue rate, reported under "issue" from the kernel's executed warp instructions (ncu
    # smsp__inst_executed.sum of the same launch shape, profiles/inst_counts.json) over the live time.
    roof = {"kernel": top, "bound": "tensor", "unit": "TFLOP/s", "avg_us": top_us,
            "achieved": alg.get(top, 0) / (top_us * 1e-6) / 1e12, "peak": bf16_peak,
            "peak_source": f"MEASURED_PEAKS.json bf16 burst ({src}); achieved = canonical (un-collapsed) FLOPs of "
                           "SURVEY 8(d) / live kernel time; the executed arithmetic is fp32 on the FMA/ALU pipes",
            "traffic": None, "kernels": kernels}
    roof["frac"] = roof["achieved"] / roof["peak"]
    try:
        ic = json.load(open(os.path.join(ROOT, "profiles", "inst_counts.json")))
        key = f"{args.workload}_B{B}_v{variant}"
        sm_clk = (clocks or {}).get("sm_mhz") or 1965.0
        issue = {}
        for name, rec in ic.get(key, {}).items():
            if name in kernels:
                slots = kernels[name]["avg_us"] * 1e-6 * sm_clk * 1e6 * 148 * 4
                issue[name] = {"warp_insts": rec["inst"], "issue_frac": rec["inst"] / slots,
                               "dram_bytes": rec.get("dram_bytes")}
        roof["issue"] = {"unit": "fraction of 148 SM x 4 issue slots x SM clock", "sm_mhz": sm_clk, "kernels": issue,
                         "source": ic.get("_source")}
        if top in issue:
            roof["traffic"] = issue[top]["dram_bytes"]
    except Exception:
        pass
    step_flops = 3 * flops_fwd_per_commit(Ne, Nc, variant)
    roof["step"] = {"algorithmic_tflops": value * step_flops / 1e12,
                    "algorithmic_hbm_gbs": value * bytes_per_commit_train(Ne, Nc) / 1e9,
                    "hbm_peak_gbs": hbm_peak, "hbm_frac": value * bytes_per_commit_train(Ne, Nc) / 1e9 / hbm_peak}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_sample = min(B, 100)
        rate, cores, sec = cpu_oracle_rate(Ne, Nc, variant, cpu_sample, 10, 1)
        cpu = {"value": rate, "unit": "commits/s", "cores": cores, "kind": "port",
               "sample": f"{cpu_sample} commits/step x 10 steps, closed-form PyTorch-CPU fp32 restatement "
                         "(oracle/hdgnn_oracle.py) incl. autograd backward and TF-Adam"}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "commits/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "variant": variant, "commits_per_gpu_per_step": B, "global_batch": Bg,
                       "parallelism": f"commit-sharded dp{world}, " + (
                           "no collective" if world == 1 else
                           "gradient all-reduce fused into the reduce+Adam kernel over NVLink peer memory (no NCCL call in the step)"
                           if model.peer else "one NCCL gradient all-reduce/step between backward and Adam"),
                       "l2": f"inputs rotate through {pool_n} batch buffers ({n_distinct} distinct batches) = {pool_n * batch_bytes / 2**20:.0f} MiB > 126 MiB L2"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "commits/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": host_pool[0].nbytes(), "d2h_bytes_per_step": 12,
                    "wire_format": "label grids as bitmaps (HDGNN_F_LABEL_BITS), x f32, hmap i32, L i32" if model.host_bits else "label grids as u8",
                    "api": "hdgnn_b200.model.graph2graph.train_step: pinned host buffers -> hdgnn_train_step_host (1 GPU) / "
                           "hdgnn_train_step_peer_host (N GPUs, peer exchange) or hdgnn_forward_backward_host + NCCL all-reduce + hdgnn_adam_step; "
                           "H2D of step k+1 overlaps the kernels of step k"},
            "gpu_launches": launches,
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
