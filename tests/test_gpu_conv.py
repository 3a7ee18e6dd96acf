"""GPU parity of the legacy operator (model.py:335-403): normalize_adj fused with propagation and the Chebyshev
map_conv regulariser, against the oracle (literal S.T / matrix_diag transcription and its closed form)."""
import numpy as np
import pytest
import torch

from oracle import hdgnn_oracle as O

pytestmark = pytest.mark.gpu


def _adj(B, N, p, seed):
    rng = np.random.default_rng(seed)
    a = (rng.random((B, N, N)) < p).astype(np.uint8)
    a[:, np.arange(N), np.arange(N)] = 0
    return a


def _ref_propagate(adj, H, W, bias, eps, self_loop, relu, transpose):
    A = adj.astype(np.float64)
    N = A.shape[1]
    if self_loop:
        A = A + np.eye(N)
    d = (A.sum(2) + eps) ** -0.5
    M = np.transpose(A, (0, 2, 1)) if transpose else A
    Ahat = d[:, :, None] * M * d[:, None, :]
    Z = H.astype(np.float64) if W is None else H.astype(np.float64) @ W.astype(np.float64)
    out = Ahat @ Z + (0 if bias is None else bias.astype(np.float64))
    return (np.maximum(out, 0) if relu else out), d


@pytest.mark.parametrize("B,N,d_in,d_out,p", [(3, 9, 4, 4, 0.3), (4, 50, 7, 20, 0.1), (2, 200, 20, 20, 0.05), (2, 250, 1, 1, 0.05),
                                              (1, 300, 20, 32, 0.02)])
@pytest.mark.parametrize("flags", [0, 1, 2 | 4, 1 | 2, 8, 8 | 1 | 2 | 4, 16, 16 | 1 | 4])      # 8 = CUDA-core path, 16 = one-CTA-per-commit tcgen05 kernel, else the persistent tcgen05 kernel when N <= 256
def test_normalize_propagate(B, N, d_in, d_out, p, flags):
    from hdgnn_b200.engine import normalize_propagate
    rng = np.random.default_rng(N + flags)
    adj = _adj(B, N, p, N)
    H = rng.normal(size=(B, N, d_in)).astype(np.float32)
    W = None if d_in == d_out and flags in (0, 8) else rng.normal(scale=0.3, size=(d_in, d_out)).astype(np.float32)
    bias = rng.normal(size=d_out).astype(np.float32) if flags & 2 else None
    out, dinv = normalize_propagate(torch.tensor(adj).cuda(), torch.tensor(H).cuda(), None if W is None else torch.tensor(W).cuda(),
                                    None if bias is None else torch.tensor(bias).cuda(), eps=1e-3, flags=flags)
    torch.cuda.synchronize()
    ref, d = _ref_propagate(adj, H, W, bias, 1e-3, bool(flags & 1), bool(flags & 2), not (flags & 4))
    assert np.abs(dinv.cpu().numpy() - d).max() / d.max() < 1e-6
    assert np.abs(out.cpu().numpy() - ref).max() / max(np.abs(ref).max(), 1e-30) < 1e-5      # fp32 sums vs fp64


@pytest.mark.parametrize("B,N,p", [(3, 6, 0.4), (5, 40, 0.15), (4, 200, 0.05), (2, 333, 0.03)])
def test_map_conv_matches_oracle(B, N, p):
    from hdgnn_b200.engine import map_conv
    rng = np.random.default_rng(N)
    adj = _adj(B, N, p, N + 1)
    x = rng.integers(0, 10, size=(B, N)).astype(np.float32) / 3.0
    theta = torch.tensor([0.13, -0.21], dtype=torch.float64)
    ref = float(O.map_conv_closed(theta, torch.tensor(adj), torch.tensor(x, dtype=torch.float64)))
    if N <= 40:     # the literal transcription (S, T, matrix_diag, transposes) agrees with the closed form
        No = N
        Ra = torch.zeros(B, 2, No * (No - 1), dtype=torch.float64)
        q = 0
        for i in range(No):
            for j in range(No):
                if i != j:
                    Ra[:, 1, q] = torch.tensor(adj[:, i, j], dtype=torch.float64); q += 1
        Ra[:, 0] = 1 - Ra[:, 1]
        lit = float(O.map_conv_dense(theta, Ra, torch.tensor(x, dtype=torch.float64).reshape(B, 1, No)))
        assert abs(lit - ref) <= 1e-9 * abs(ref)
    loss, per = map_conv(torch.tensor(adj).cuda(), torch.tensor(x).cuda(), theta.float().cuda())
    torch.cuda.synchronize()
    assert abs(loss.item() - ref) <= 1e-4 * abs(ref), (loss.item(), ref)
    # textbook variant: A + I, un-transposed
    ref2 = float(O.map_conv_closed(theta, torch.tensor(np.transpose(adj, (0, 2, 1)).copy()), torch.tensor(x, dtype=torch.float64), self_loop=True))
    loss2, _ = map_conv(torch.tensor(adj).cuda(), torch.tensor(x).cuda(), theta.float().cuda(), flags=1 | 4)
    torch.cuda.synchronize()
    # with the self loop the degrees of A^T + I differ from those of A + I, so only compare on symmetric graphs
    sym = np.maximum(adj, np.transpose(adj, (0, 2, 1)))
    ref3 = float(O.map_conv_closed(theta, torch.tensor(sym), torch.tensor(x, dtype=torch.float64), self_loop=True))
    loss3, _ = map_conv(torch.tensor(sym).cuda(), torch.tensor(x).cuda(), theta.float().cuda(), flags=1 | 4)
    torch.cuda.synchronize()
    assert abs(loss3.item() - ref3) <= 1e-4 * abs(ref3)
    del ref2, loss2


def test_tile_too_large_is_reported():
    from hdgnn_b200.engine import normalize_propagate
    from hdgnn_b200._lib import HdgnnError, E_UNSUPPORTED
    adj = torch.zeros(1, 500, 512, dtype=torch.uint8, device="cuda")
    with pytest.raises(HdgnnError) as e:
        normalize_propagate(adj, torch.zeros(1, 500, 20, device="cuda"))
    assert e.value.code == E_UNSUPPORTED


@pytest.mark.parametrize("flags", [0, 4, 1 | 2])
def test_normalize_propagate_persistent_many_commits(flags):
    """More commits than SMs: every CTA of the persistent tcgen05 kernel walks several commits with the prefetch."""
    from hdgnn_b200.engine import normalize_propagate
    B, N, d = 450, 200, 20
    rng = np.random.default_rng(3)
    adj = _adj(B, N, 0.05, 9)
    H = rng.normal(size=(B, N, d)).astype(np.float32)
    bias = rng.normal(size=d).astype(np.float32) if flags & 2 else None
    out, dinv = normalize_propagate(torch.tensor(adj).cuda(), torch.tensor(H).cuda(), None,
                                    None if bias is None else torch.tensor(bias).cuda(), eps=1e-3, flags=flags)
    torch.cuda.synchronize()
    ref, dd = _ref_propagate(adj, H, None, bias, 1e-3, bool(flags & 1), bool(flags & 2), not (flags & 4))
    assert np.abs(dinv.cpu().numpy() - dd).max() / dd.max() < 1e-6
    assert np.abs(out.cpu().numpy() - ref).max() / np.abs(ref).max() < 1e-5


def test_legacy_operator_matches_reference_model_py_golden():
    """Row a16 pinned: the CUDA normalize+propagate and map_conv against the reference's own model.py:335-403 executed over
    the shim (tests/golden/legacy_toy.npz, written by oracle/gen_golden.py)."""
    import os
    from hdgnn_b200.engine import map_conv, normalize_propagate
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "legacy_toy.npz"))
    adj = torch.tensor(z["adj"]).cuda()
    mb, No = z["adj"].shape[:2]
    x = torch.tensor(z["O"].reshape(mb, No), dtype=torch.float32).cuda()
    theta = torch.tensor(z["theta"], dtype=torch.float32).cuda()
    loss, per = map_conv(adj, x, theta)
    torch.cuda.synchronize()
    assert abs(loss.item() - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    # A_hat of normalize_adj recovered from the reference's T_1 = (2/1.5)(I - A_hat) - I, applied to random features
    Ahat = np.eye(No) - 0.75 * (z["t_k"][:, 1] + np.eye(No))
    H = np.random.default_rng(5).normal(size=(mb, No, 20)).astype(np.float32)
    for flags in (0, 8, 16):
        out, _ = normalize_propagate(adj, torch.tensor(H).cuda(), flags=flags)
        torch.cuda.synchronize()
        ref = Ahat @ H.astype(np.float64)
        assert np.abs(out.cpu().numpy() - ref).max() / np.abs(ref).max() < 1e-5, flags


def test_legacy_operator_backward_matches_reference_gradients():
    """map_conv backward against the gradients autograd gives for the reference's own model.py:394-403 (golden dO, dtheta)."""
    import os
    from hdgnn_b200.engine import map_conv_backward
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "legacy_toy.npz"))
    mb, No = z["adj"].shape[:2]
    adj = torch.tensor(z["adj"]).cuda()
    x = torch.tensor(z["O"].reshape(mb, No), dtype=torch.float32).cuda()
    theta = torch.tensor(z["theta"], dtype=torch.float32).cuda()
    dx, dth = map_conv_backward(adj, x, theta)
    torch.cuda.synchronize()
    ref_dx, ref_dth = z["dO"].reshape(mb, No), z["dtheta"]
    assert np.abs(dx.cpu().numpy() - ref_dx).max() / np.abs(ref_dx).max() < 1e-5
    assert np.abs(dth.cpu().numpy() - ref_dth).max() / np.abs(ref_dth).max() < 1e-5


@pytest.mark.parametrize("B,N,p", [(3, 6, 0.4), (4, 200, 0.05), (2, 333, 0.03)])
@pytest.mark.parametrize("flags", [0, 1, 4, 1 | 4])
def test_map_conv_backward_matches_autograd(B, N, p, flags):
    """dx, dtheta against autograd of an fp64 restatement with the same switches (self loop, un-transposed form); flags = 0 is
    the reference's form, which test_legacy_operator_backward_matches_reference_gradients pins to model.py itself."""
    from hdgnn_b200.engine import map_conv_backward
    rng = np.random.default_rng(N + flags)
    adj = _adj(B, N, p, N + 2)
    x = rng.integers(0, 10, size=(B, N)).astype(np.float64) / 3.0
    th = torch.tensor([0.13, -0.21], dtype=torch.float64, requires_grad=True)
    xt = torch.tensor(x, requires_grad=True)
    A = torch.tensor(adj, dtype=torch.float64)
    if flags & 1:
        A = A + torch.eye(N, dtype=torch.float64)
    d = (A.sum(2) + float(np.float32(1e-3))) ** -0.5
    Ahat = d[:, :, None] * (A if flags & 4 else A.transpose(1, 2)) * d[:, None, :]
    t = torch.softmax(th, 0)
    Lx = (2.0 / 1.5) * (xt - (Ahat @ xt[..., None])[..., 0]) - xt
    loss = (((xt * (t[0] * xt + t[1] * Lx)).sum(1)) ** 2).mean()
    if flags == 0:
        assert abs(float(loss) - float(O.map_conv_closed(th, torch.tensor(adj), xt))) <= 1e-10 * abs(float(loss))
    gx, gt = torch.autograd.grad(loss, [xt, th])
    dx, dth = map_conv_backward(torch.tensor(adj).cuda(), torch.tensor(x, dtype=torch.float32).cuda(), th.detach().float().cuda(),
                                flags=flags, gscale=0.7)
    torch.cuda.synchronize()
    assert np.abs(dx.cpu().numpy() - 0.7 * gx.numpy()).max() / np.abs(gx.numpy()).max() < 5e-5
    assert np.abs(dth.cpu().numpy() - 0.7 * gt.numpy()).max() / np.abs(gt.numpy()).max() < 5e-5


@pytest.mark.parametrize("B,N,d_in,d_out,p", [(3, 9, 4, 4, 0.3), (4, 50, 7, 20, 0.1), (2, 200, 20, 20, 0.05), (2, 300, 20, 32, 0.02)])
@pytest.mark.parametrize("flags", [0, 2, 1 | 2, 4, 8 | 2, 16])
def test_normalize_propagate_backward_matches_autograd(B, N, d_in, d_out, p, flags):
    """dH, dW, dbias of act(A_hat (H W) + b) against autograd of the fp64 restatement; A_hat^T dPre runs on the forward's kernels."""
    from hdgnn_b200.engine import normalize_propagate, normalize_propagate_backward
    rng = np.random.default_rng(N + 3 * flags)
    adj = _adj(B, N, p, N + 5)
    H = rng.normal(size=(B, N, d_in))
    W = rng.normal(scale=0.3, size=(d_in, d_out))
    bias = rng.normal(size=d_out) if flags & 2 else np.zeros(d_out)
    dOut = rng.normal(size=(B, N, d_out))
    A = torch.tensor(adj, dtype=torch.float64)
    if flags & 1:
        A = A + torch.eye(N, dtype=torch.float64)
    d = (A.sum(2) + float(np.float32(1e-3))) ** -0.5
    M = A if flags & 4 else A.transpose(1, 2)
    Ahat = d[:, :, None] * M * d[:, None, :]
    Ht = torch.tensor(H, requires_grad=True); Wt = torch.tensor(W, requires_grad=True); bt = torch.tensor(bias, requires_grad=True)
    pre = Ahat @ (Ht @ Wt) + bt
    out_ref = torch.relu(pre) if flags & 2 else pre
    gH, gW, gb = torch.autograd.grad((out_ref * torch.tensor(dOut)).sum(), [Ht, Wt, bt])
    adj_d = torch.tensor(adj).cuda()
    Hd, Wd = torch.tensor(H, dtype=torch.float32).cuda(), torch.tensor(W, dtype=torch.float32).cuda()
    bd = torch.tensor(bias, dtype=torch.float32).cuda() if flags & 2 else None
    out, _ = normalize_propagate(adj_d, Hd, Wd, bd, eps=1e-3, flags=flags)
    dH, dW, db = normalize_propagate_backward(adj_d, Hd, torch.tensor(dOut, dtype=torch.float32).cuda(), W=Wd, out=out, eps=1e-3, flags=flags)
    torch.cuda.synchronize()
    rel = lambda a, r: float(np.abs(a.cpu().numpy() - r.numpy()).max() / max(np.abs(r.numpy()).max(), 1e-30))
    assert rel(dH, gH) < 2e-5 and rel(dW, gW) < 2e-5 and rel(db, gb) < 2e-5
