"""HBM bandwidth of the device loader (hdgnn_compact_from_raw) and the device evaluation (hdgnn_eval_counts).
python tools/bench_io.py > gpurun_out/bench_io.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hdgnn_b200.engine import compact_from_raw_device, eval_counts, label_pitch, lib, _p
import ctypes as C

peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


out = {"hbm_peak_gbs": peak, "loader": [], "eval": []}
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
for (N, n, dt) in [(100, 200, torch.float64), (2048, 200, torch.float64), (512, 512, torch.float64), (1024, 512, torch.float32)]:
    raw = (torch.rand(N, n, n, device="cuda") < 0.05).to(dt)
    pitch = label_pitch(n)
    grid = torch.empty(N, n, pitch, dtype=torch.uint8, device="cuda")
    diag = torch.empty(N, n, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    f = lambda: lib.hdgnn_compact_from_raw(N, n, _p(raw), 1 if dt == torch.float64 else 0, _p(grid), pitch, _p(diag), _p(err), st())
    t = timeit(f)
    by = raw.numel() * raw.element_size() + grid.numel() + diag.numel() * 4
    out["loader"].append(dict(N=N, n=n, dtype=str(dt), us=t * 1e6, bytes=by, gbs=by / t / 1e9, frac=by / t / 1e9 / peak))
for (B, Nc, dens, auc) in [(100, 74, 0.05, False), (100, 74, 0.05, True), (4096, 74, 0.05, False), (100, 512, 0.05, False), (512, 256, 0.05, False),
                           (100, 150, 0.05, True)]:
    Ncr = Nc * (Nc - 1)
    probs = torch.softmax(torch.randn(B, 2, Ncr, device="cuda"), 1).contiguous()
    Y = (torch.rand(B, Nc, Nc, device="cuda") < dens).to(torch.uint8)
    counts = torch.empty(B, 8, dtype=torch.int64, device="cuda")
    aucb = torch.empty(B, 2, dtype=torch.int64, device="cuda") if auc else None
    t = timeit(lambda: lib.hdgnn_eval_counts(B, Nc, _p(probs), _p(Y), Nc, _p(counts), _p(aucb), 0, st()))
    by = probs.numel() * 4 + Y.numel()
    out["eval"].append(dict(B=B, Nc=Nc, auc=auc, us=t * 1e6, bytes=by, gbs=by / t / 1e9, frac=by / t / 1e9 / peak))
print(json.dumps(out, indent=1))
