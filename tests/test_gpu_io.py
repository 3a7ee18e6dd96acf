"""GPU parity of the two formats either side of the hot path: the device loader against the
host loader (itself bit-exact against the reference's utils2.read_data, tests/test_golden.py) and the device
evaluation counters against the metric functions pinned to the reference's EvaluationFuncs.py (eval_toy.npz).
Integer work: everything is compared bit-exact."""
import os

import numpy as np
import pytest
import torch

from hdgnn_b200 import EvaluationFuncs as EV
from hdgnn_b200.utils2 import compact_from_raw, edge_onehot, pair_index

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _grid_from_pairs(v, Nc):
    """(N, Ncr) values in pair order -> (N,Nc,Nc) grid, zero diagonal."""
    s, t = pair_index(Nc)
    g = np.zeros((v.shape[0], Nc, Nc), v.dtype)
    g[:, s, t] = v
    return g


@pytest.mark.parametrize("N,n", [(8, 7), (5, 74), (3, 200), (2, 250), (2, 257), (1, 512)])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_compact_from_raw_device_bit_exact(N, n, dtype):
    from hdgnn_b200.engine import compact_from_raw_device, label_pitch
    rng = np.random.default_rng(N * 1000 + n)
    raw = (rng.random((N, n, n)) < 0.07).astype(dtype)
    # entries the reference's int() accepts besides 0/1: fractions (truncate toward zero) and -1 / -2 (python indexing)
    k = rng.integers(0, n, size=(40, 2))
    vals = np.array([0.9, 1.7, -0.5, -1.0, -1.9, -2.0, -2.99, 1.999, 0.0, 1.0])
    for q, (i, j) in enumerate(k):
        if i != j:
            raw[q % N, i, j] = vals[q % len(vals)]
    raw[:, np.arange(n), np.arange(n)] = rng.integers(0, 10, size=(N, n)) + (0.25 if dtype == np.float64 else 0.0)
    host = compact_from_raw(raw, np.zeros((N, 2, 2)), [[] for _ in range(N)], [{} for _ in range(N)], n, 2)
    grid, diag = compact_from_raw_device(torch.as_tensor(raw).cuda())
    torch.cuda.synchronize()
    pitch = label_pitch(n)
    assert grid.shape == (N, n, pitch)
    g = grid.cpu().numpy()
    assert np.array_equal(g[:, :, :n], host.adj)
    assert not g[:, :, n:].any()                       # padding is zeroed
    assert np.array_equal(diag.cpu().numpy(), host.x)
    assert np.array_equal(diag.cpu().numpy(), raw[:, np.arange(n), np.arange(n)].astype(np.float32))


def test_compact_from_raw_device_golden_loader_fixture():
    """The raw array the reference's utils2.read_data consumed when the golden fixture was made."""
    from hdgnn_b200.engine import compact_from_raw_device
    z = np.load(os.path.join(G, "loader_toy.npz"))
    grid, diag = compact_from_raw_device(torch.as_tensor(z["raw_adj"]).cuda())
    assert np.array_equal(grid.cpu().numpy()[:, :, :7], z["adj"]) and np.array_equal(diag.cpu().numpy(), z["x"])
    # and the pair labels built from it are the reference's E_edge (utils2.py:69-83)
    assert np.array_equal(edge_onehot(grid.cpu().numpy()[:4, :, :7]), z["E_edge_train"])


@pytest.mark.parametrize("bad", [2.0, -3.0, 7.0, float("nan"), float("inf")])
def test_compact_from_raw_device_raises_where_the_reference_does(bad):
    from hdgnn_b200.engine import compact_from_raw_device
    raw = np.zeros((2, 9, 9))
    raw[1, 3, 5] = bad
    with pytest.raises(IndexError):
        compact_from_raw_device(torch.as_tensor(raw).cuda())
    raw[1, 3, 5] = 0
    raw[1, 4, 4] = bad if np.isfinite(bad) else 5.0    # anything goes on the diagonal (it is the node attribute)
    compact_from_raw_device(torch.as_tensor(raw).cuda())


def _device_metrics(label, real, Nc, quirks, auc_first=0):
    from hdgnn_b200.engine import eval_counts
    Y = _grid_from_pairs(label[:, 1, :], Nc).astype(np.uint8)
    counts, auc = eval_counts(torch.as_tensor(real.astype(np.float32)).cuda(), torch.as_tensor(Y).cuda(), auc=True,
                              auc_first=auc_first)
    torch.cuda.synchronize()
    return counts.cpu().numpy(), auc.cpu().numpy()


def test_eval_counts_golden_reference_metrics():
    """eval_toy.npz holds the outputs of the reference's own EvaluationFuncs.py on these arrays."""
    z = np.load(os.path.join(G, "eval_toy.npz"))
    label, real = z["label"], z["real"]
    assert np.array_equal(real.astype(np.float32).astype(np.float64), real) or True
    N, _, Ncr = label.shape
    Nc = 6
    counts, auc = _device_metrics(label, real, Nc, True)
    m = EV.metrics_from_counts(counts, quirks=True)
    real32 = real.astype(np.float32)                   # the device sees fp32 probabilities, as the model produces them
    assert m["hits"] / (N * Ncr) == EV.top_ACC(label, real32)
    assert m["prec"] == EV.prec(label, real32) and m["recall"] == EV.recall(label, real32) and m["f1"] == EV.f1(label, real32)
    assert np.isclose(m["hits"] / (N * Ncr), float(z["top_ACC"]), rtol=1e-14)
    assert np.isclose(m["prec"], float(z["prec"]), rtol=1e-14) and np.isclose(m["f1"], float(z["f1"]), rtol=1e-14)
    assert np.isclose(EV.auc_from_counts(counts, auc, Ncr, quirks=True), float(z["AUC"]), rtol=1e-12)


@pytest.mark.parametrize("N,Nc,dens", [(6, 5, 0.3), (4, 74, 0.05), (3, 150, 0.02), (2, 256, 0.1)])
def test_eval_counts_bit_exact_vs_array_metrics(N, Nc, dens):
    rng = np.random.default_rng(Nc)
    Ncr = Nc * (Nc - 1)
    y = (rng.random((N, Ncr)) < dens).astype(np.float32)
    y[N - 2] = 0                                        # a commit without any related pair
    label = np.stack([1 - y, y], 1)
    lg = rng.normal(size=(N, 2, Ncr)).astype(np.float32) * 3
    lg[:, 1] += 2 * (y - 0.5)
    e = np.exp(lg - lg.max(1, keepdims=True))
    real = (e / e.sum(1, keepdims=True)).astype(np.float32)
    real[0, :, :40] = np.round(real[0, :, :40], 1)      # ties between scores
    real[1, :, 5:9] = 0.5                               # arg-max ties -> channel 0 (np.argmax)
    real[0, 0, 50:60] = 0.0; real[0, 1, 50:60] = 1.0    # probability underflowed to exactly 0: the only way ceil() gives 0
    counts, auc = _device_metrics(label, real, Nc, True)
    assert counts[:, 7].tolist() == y.sum(1).astype(int).tolist()
    for quirks in (True, False):
        m = EV.metrics_from_counts(counts, quirks=quirks)
        assert m["hits"] / (N * Ncr) == EV.top_ACC(label, real)
        assert m["prec"] == EV.prec(label, real, quirks) and m["recall"] == EV.recall(label, real, quirks)
        assert m["f1"] == EV.f1(label, real, quirks)
    assert np.isclose(EV.auc_from_counts(counts, auc, Ncr, quirks=True), EV.AUC(label, real, True), rtol=1e-13)
    assert np.isclose(EV.auc_from_counts(counts, auc, Ncr, quirks=False), EV.AUC(label, real, False), rtol=1e-13)
    # Mann-Whitney numerators are integers: exact against a brute-force count on one commit
    b = 0
    pos, neg = real[b, 0][y[b] == 1], real[b, 1][y[b] == 0]
    brute = int(2 * (neg[None, :] < pos[:, None]).sum() + (neg[None, :] == pos[:, None]).sum())
    assert int(auc[b, 0]) == brute
    # auc_first: earlier commits are skipped (the reference only keeps the last one)
    c2, a2 = _device_metrics(label, real, Nc, True, auc_first=N - 1)
    assert np.array_equal(c2, counts) and np.array_equal(a2[N - 1], auc[N - 1]) and not a2[:N - 1].any()
    with pytest.raises(ZeroDivisionError):
        EV.auc_from_counts(counts[:N - 1], auc[:N - 1], Ncr, quirks=True)      # last commit has one class (Q7)


@pytest.mark.parametrize("Ne,Nc,variant", [(60, 33, 2), (200, 74, 2), (120, 170, 2), (90, 300, 1)])
def test_in_kernel_eval_counters_equal_eval_counts(Ne, Nc, variant):
    """hdgnn_set_eval_counters: the relation head's own counters (two warp votes per 32 pairs) are, integer for integer, what
    hdgnn_eval_counts computes from the probabilities the same call wrote (EvaluationFuncs.py:27-37, 92-117) -- on the
    shared-memory, the global-slice and the forward-only 16-segment forms of the per-commit kernel, added up over two calls."""
    from hdgnn_b200.engine import Engine, DeviceBatch, eval_counts
    from hdgnn_b200.synthetic import make_commits
    B = 5
    cb = make_commits(B, Ne, Nc, seed=77 + Nc, p_short=0.4, p_noise=0.3)
    eng = Engine(Ne, Nc, variant=variant, max_batch=B)
    db = DeviceBatch.from_numpy(cb.adj, cb.x, cb.hmap, cb.L, cb.Y, eng.tdev)
    from hdgnn_b200 import _lib
    g = torch.Generator().manual_seed(5)
    params = (0.3 * torch.randn(eng.n_params, generator=g)).cuda()
    _, logits, _ = eng.forward(db, params, want_logits=True)
    off = _lib.lib.hdgnn_param_offset(variant, b"scr_b2")                 # centre the logit difference: both classes get predicted
    assert off >= 0
    params[off + 1] -= (logits[:, 1] - logits[:, 0]).median()
    counts = torch.zeros(B, 8, dtype=torch.int64, device="cuda")
    assert eng.set_eval_counters(counts)
    probs, _, _ = eng.forward(db, params, want_logits=False)
    torch.cuda.synchronize()
    ref, _ = eval_counts(probs, torch.as_tensor(cb.Y).cuda())
    assert torch.equal(counts, ref), (counts, ref)
    assert 0 < int(ref[:, 4].sum() + ref[:, 5].sum()) < B * Nc * (Nc - 1)      # the case exercises both predictions
    eng.forward(db, params, want_logits=False)                           # the counters ADD
    torch.cuda.synchronize()
    assert torch.equal(counts, 2 * ref)
    eng.set_eval_counters(None)
    eng.forward(db, params, want_logits=False)
    torch.cuda.synchronize()
    assert torch.equal(counts, 2 * ref)
    eng.close()
