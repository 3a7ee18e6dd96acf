"""Per-phase clock64 breakdown of mid2_kernel (debug handle): mean / max cycles per phase over the commits."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hdgnn_b200.engine import Engine, DeviceBatch
from hdgnn_b200.synthetic import make_commits

Ne, Nc, B = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (200, 74, 100)))
variant = int(sys.argv[4]) if len(sys.argv) > 4 else 2
cb = make_commits(B, Ne, Nc, seed=20260)
eng = Engine(Ne, Nc, variant=variant, max_batch=B, flags=2)
db = DeviceBatch.from_numpy(cb.adj, cb.x, cb.hmap, cb.L, cb.Y, eng.tdev)
params = (0.1 * torch.randn(eng.n_params)).cuda()
for _ in range(3):
    eng.forward_backward(db, params)
torch.cuda.synchronize()
clk = eng.workspace("CLK", (B, 24), dtype=torch.int64).cpu().numpy()
names = ["A load", "B node fwd", "D pool fwd", "tables+E hunk fwd", "F head tables", "G1 head", "G2 delta sums", "H head bwd",
         "I hunk bwd", "J+K pool bwd", "L node bwd"]
d = np.diff(clk[:, :12], axis=1)
ident = cb.L == Ne
print(f"commits {B}, L==Ne for {int(ident.sum())}")
for i, n in enumerate(names):
    print(f"{n:20s} mean {d[:, i].mean():9.0f}  max {d[:, i].max():9.0f}   ident {d[ident, i].mean():9.0f}  general {d[~ident, i].mean() if (~ident).any() else 0:9.0f}")
g = ~ident
if g.any():
    print("general commits: J (9->12) %.0f | K gather (12->13) %.0f | K rest (13->10) %.0f | D: scan+TP gather (2->14) %.0f, rest (14->3) %.0f" % (
        (clk[g, 12] - clk[g, 9]).mean(), (clk[g, 13] - clk[g, 12]).mean(), (clk[g, 10] - clk[g, 13]).mean(),
        (clk[g, 14] - clk[g, 2]).mean(), (clk[g, 3] - clk[g, 14]).mean()))
print("ident D: X sum (2->13) %.0f | degrees (13->14) %.0f | fill (14->15) %.0f | seg reduce (15->3) %.0f" % (
    (clk[ident, 13] - clk[ident, 2]).mean(), (clk[ident, 14] - clk[ident, 13]).mean(), (clk[ident, 15] - clk[ident, 14]).mean(),
    (clk[ident, 3] - clk[ident, 15]).mean()))
if g.any():
    print("general D: seg reduce (15->3) %.0f" % (clk[g, 3] - clk[g, 15]).mean())
print("ident commits: J (9->12) %.0f | K (12->10) %.0f" % ((clk[ident, 12] - clk[ident, 9]).mean(), (clk[ident, 10] - clk[ident, 12]).mean()))
if clk[:, 16].any():
    print("inline entity stage: prologue incl. sort (0->1) %.0f | weights (1->16) %.0f | fwd items (16->17) %.0f | node fwd (17->2) %.0f | L (10->18) %.0f | SG scan (18->19) %.0f | bwd items+reduce (19->11) %.0f" % (
        (clk[:, 1] - clk[:, 0]).mean(), (clk[:, 16] - clk[:, 1]).mean(), (clk[:, 17] - clk[:, 16]).mean(), (clk[:, 2] - clk[:, 17]).mean(),
        (clk[:, 18] - clk[:, 10]).mean(), (clk[:, 19] - clk[:, 18]).mean(), (clk[:, 11] - clk[:, 19]).mean()))
if variant == 4:
    print("variant 4: node-branch entity fwd (16->20 incl. edge pair sums) %.0f | edge head tables (20->21) %.0f | soft-edge fwd sweep (21->22) %.0f | edge bwd sweeps (node ent bwd end ->23) ... total after L: (18->23) %.0f | edge node-level + tied layer (23->11) %.0f" % (
        (clk[:, 20] - clk[:, 16]).mean(), (clk[:, 21] - clk[:, 20]).mean(), (clk[:, 22] - clk[:, 21]).mean(), (clk[:, 23] - clk[:, 18]).mean(), (clk[:, 11] - clk[:, 23]).mean()))
elif clk[:, 16].any():
    print("  fwd items: dense part (16->20) %.0f | row walk (20->21) %.0f | column walk (21->22) %.0f" % (
        (clk[:, 20] - clk[:, 16]).mean(), (clk[:, 21] - clk[:, 20]).mean(), (clk[:, 22] - clk[:, 21]).mean()))
tot = clk[:, 11] - clk[:, 0]
print(f"total mean {tot.mean():.0f} max {tot.max():.0f} cycles")
