"""Achieved HBM bandwidth of the legacy operator kernels (k_conv.cu): bytes = adjacency tile + H/x in + out."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hdgnn_b200.engine import normalize_propagate, map_conv, normalize_propagate_backward, map_conv_backward, label_pitch

peak = 6543.4
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
res = []
for (B, N, d, p) in [(4096, 200, 1, 0.05), (4096, 200, 20, 0.05), (4096, 200, 20, 0.5), (2048, 250, 20, 0.05)]:
    pitch = label_pitch(N)
    g = torch.Generator(device="cuda").manual_seed(1)
    adj = (torch.rand(B, N, pitch, device="cuda", generator=g) < p).to(torch.uint8)
    adj[:, :, N:] = 0
    x = torch.rand(B, N, device="cuda", generator=g)
    H = torch.rand(B, N, d, device="cuda", generator=g)
    theta = torch.tensor([0.1, -0.2], device="cuda")
    dOut = torch.rand(B, N, d, device="cuda", generator=g)
    for name, fn, nbytes in [("map_conv", lambda: map_conv(adj, x, theta), B * (N * pitch + 4 * N + 4)),
                             ("map_conv_backward", lambda: map_conv_backward(adj, x, theta), B * (N * pitch + 8 * N + 8)),
                             (f"normalize_propagate d={d}", lambda: normalize_propagate(adj, H), B * (N * pitch + 8 * N * d + 4 * N)),
                             (f"normalize_propagate_backward d={d}", lambda: normalize_propagate_backward(adj, H, dOut), B * (N * pitch + 12 * N * d))]:
        if name.startswith("map_conv") and d != 1:
            continue
        if not name.startswith("map_conv") and d == 1:
            continue
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gbs = nbytes / ms / 1e6
        res.append({"kernel": name, "B": B, "N": N, "density": p, "ms": ms, "GB/s": gbs, "frac_of_measured_hbm_peak": gbs / peak})
        print(res[-1])
json.dump(res, open(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "bench_conv.json"), "w"), indent=1)
