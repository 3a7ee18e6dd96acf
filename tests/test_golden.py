"""The oracle and the host-side loader / metrics against golden vectors produced by EXECUTING the
reference's own files (oracle/gen_golden.py: utils2.read_data and EvaluationFuncs run as they are;
model_1..4.graph2graph run over oracle/tf1_shim.py because TensorFlow is not installable here)."""
import os

import numpy as np
import pytest
import torch

from hdgnn_b200 import EvaluationFuncs as EV
from hdgnn_b200.synthetic import CommitBatch
from hdgnn_b200.utils2 import dense_feeds, edge_onehot, pair_index
from oracle import hdgnn_oracle as O
from oracle import plan_numpy as PN

G = os.path.join(os.path.dirname(__file__), "golden")
Ne, Nc, N, MB = 7, 4, 8, 4


def _loader():
    z = np.load(os.path.join(G, "loader_toy.npz"))
    return z, CommitBatch(z["adj"], z["x"], z["hmap"], z["L"], z["Y"])


def test_loader_matches_reference_read_data_bit_exact():
    z, cb = _loader()
    node, E_edge, C_edge, Es, Et, Cs, Ct, Esc, Etc = dense_feeds(cb, Ne, Nc, lead=N + 1)
    h = N // 2
    assert np.array_equal(node[:h], z["E_node_train"]) and np.array_equal(node[h:N], z["E_node_test"])
    assert np.array_equal(E_edge[:h], z["E_edge_train"]) and np.array_equal(C_edge[:h], z["C_edge_train"])
    # the reference slices E_edge[h:N] / C_edge[h:N] for the test half (utils2.py:144-149)
    assert np.array_equal(E_edge[h:N], z["E_edge_test"]) and np.array_equal(C_edge[h:N], z["C_edge_test"])
    for mine, name in ((Es, "Es_data"), (Et, "Et_data"), (Cs, "Cs_label"), (Ct, "Ct_label"), (Esc, "Esc_data"), (Etc, "Etc_data")):
        assert mine.dtype == z[name].dtype == np.float32
        assert np.array_equal(mine, z[name]), name
    # the compact form really is what was on disk
    raw = z["raw_adj"]
    assert np.array_equal(cb.x, raw[:, np.arange(Ne), np.arange(Ne)].astype(np.float32))
    assert (cb.L < Ne).any() and (cb.L == Ne).any() and (cb.hmap < 0).any()


def test_pair_enumeration_closed_form():
    for n in (2, 5, 9):
        i, j = pair_index(n)
        p = i * (n - 1) + j - (j > i)
        assert np.array_equal(p, np.arange(n * (n - 1)))           # utils2.py:69-83 counter order


@pytest.mark.parametrize("variant", [1, 2, 3, 4])
def test_oracle_matches_reference_model_files(variant):
    z, cb = _loader()
    g = np.load(os.path.join(G, f"model_{variant}_toy.npz"))
    spec = O.param_spec(variant)
    assert [int(np.prod(s[2])) for s in spec] == list(g["var_sizes"])        # TF creation order == blob order
    for (_, tfname, shape), gname, gshape in zip(spec, g["var_names"], g["var_shapes"]):
        tf_leaf = str(gname).split("/")[-1].split(":")[0]
        assert tf_leaf.startswith(tfname.split("/")[-1]), (tfname, gname)       # e.g. o1_w2r and TF's uniquified o1_w2r_1
        assert int(np.prod(shape)) == int(np.prod(eval(str(gshape))))
    flat = torch.as_tensor(g["params"])
    b = cb.slice(0, MB)                                                      # first batch: its own maps == maps[:MB]
    for dense in (True, False):
        loss, ce, lmap, grad, out = O.train_loss_and_grad(variant, flat, b.adj, b.x, b.hmap, b.L, b.Y, dense=dense)
        assert np.allclose(out["probs"].numpy(), g["probs"], rtol=1e-10, atol=1e-13)
        assert np.allclose(out["logits"].numpy(), g["logits"], rtol=1e-10, atol=1e-13)
        assert np.isclose(float(ce), float(g["ce"]), rtol=1e-12)
        assert np.isclose(float(lmap), float(g["loss_map"]), rtol=1e-12)
        assert np.isclose(float(loss), float(g["train_loss"]), rtol=1e-12)
        assert np.allclose(grad.numpy(), g["grad"], rtol=1e-9, atol=1e-14)
        if "E_node2" in g.files:
            assert np.allclose(out["E_node2"].numpy(), g["E_node2"], rtol=1e-10, atol=1e-13)
        if "E_edge2" in g.files:
            assert np.allclose(out["E_edge2"].numpy(), g["E_edge2"], rtol=1e-10, atol=1e-13)
    _, lp = O.reg_loss(flat, variant)
    assert np.isclose(float(lp), float(g["loss_para"]), rtol=1e-12)
    plan = PN.train_step_plan(variant, flat, b.adj, b.x, b.hmap, b.L, b.Y)
    assert np.allclose(plan["probs"], g["probs"], rtol=1e-10, atol=1e-13)
    assert np.allclose(plan["grad"], g["grad"], rtol=1e-8, atol=1e-13)
    # two TF-Adam steps (model_2.py:337)
    p = g["params"].copy(); m = np.zeros_like(p); v = np.zeros_like(p)
    p, m, v = O.tf_adam_step(p, g["grad"], m, v, 1)
    assert np.allclose(p, g["params_step1"], rtol=1e-12, atol=1e-15)
    _, _, _, grad2, _ = O.train_loss_and_grad(variant, torch.as_tensor(p), b.adj, b.x, b.hmap, b.L, b.Y)
    p, m, v = O.tf_adam_step(p, grad2.numpy(), m, v, 2)
    assert np.allclose(p, g["params_step2"], rtol=1e-10, atol=1e-14)


def test_metrics_match_reference_evaluationfuncs():
    z = np.load(os.path.join(G, "eval_toy.npz"))
    label, real = z["label"], z["real"]
    assert np.isclose(EV.top_ACC(label, real), float(z["top_ACC"]), rtol=1e-14)
    assert np.isclose(EV.prec(label, real), float(z["prec"]), rtol=1e-14)
    assert np.isclose(EV.recall(label, real), float(z["recall"]), rtol=1e-14)
    assert np.isclose(EV.f1(label, real), float(z["f1"]), rtol=1e-14)
    assert np.isclose(EV.AUC(label, real), float(z["AUC"]), rtol=1e-12)
    # the conventional definitions exist too and differ (the reference scores channel 0 after ceil)
    assert EV.prec(label, real, quirks=False) != EV.prec(label, real)


def test_edge_onehot_channel_convention():
    lab = np.zeros((1, 3, 3), np.uint8); lab[0, 0, 2] = 1
    oh = edge_onehot(lab)
    assert oh.shape == (1, 2, 6) and oh[0, 1, 1] == 1 and oh[0, 0, 1] == 0 and oh[0, 0].sum() == 5


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [1, 2, 3, 4])
def test_cuda_matches_reference_model_files(variant):
    """The CUDA path against numbers produced by the reference's own model code (fp32 vs fp64:
    tolerance 1e-3 of the tensor's scale per the north star; observed ~1e-6)."""
    from hdgnn_b200.engine import Engine, DeviceBatch
    z, cb = _loader()
    g = np.load(os.path.join(G, f"model_{variant}_toy.npz"))
    b = cb.slice(0, MB)
    eng = Engine(Ne, Nc, variant=variant, max_batch=MB)
    db = DeviceBatch.from_numpy(b.adj, b.x, b.hmap, b.L, b.Y, eng.tdev)
    params = torch.as_tensor(g["params"], dtype=torch.float32).cuda()
    probs, logits, loss, grads = eng.forward_backward(db, params, want_logits=True)
    m = torch.zeros_like(params); v = torch.zeros_like(params)
    step = torch.zeros(1, dtype=torch.int32, device="cuda"); reg = torch.zeros(2, device="cuda")
    p1 = params.clone()
    eng.adam_step(p1, grads, m, v, step, reg_losses=reg)
    torch.cuda.synchronize()
    rel = lambda a, r: float(np.abs(np.asarray(a, np.float64) - r).max() / max(np.abs(r).max(), 1e-30))
    assert rel(probs.cpu().numpy(), g["probs"]) < 1e-4
    assert rel(logits.cpu().numpy(), g["logits"]) < 1e-4
    assert abs(loss.item() - float(g["ce"])) < 1e-5 * float(g["ce"])
    assert rel(reg.cpu().numpy(), np.array([float(g["loss_map"]), float(g["loss_para"])])) < 1e-5
    assert rel(p1.cpu().numpy(), g["params_step1"]) < 1e-5
    # predicted relation classes identical
    assert np.array_equal(np.argmax(probs.cpu().numpy(), 1), np.argmax(g["probs"], 1))
    eng.close()


def test_metrics_from_integer_counters_equal_array_metrics():
    """Host half of the device evaluation (hdgnn_eval_counts -> EvaluationFuncs.metrics_from_counts / auc_from_counts):
    the counters are rebuilt here with NumPy from the golden arrays and must give the reference's numbers."""
    z = np.load(os.path.join(G, "eval_toy.npz"))
    label, real = z["label"], z["real"]
    N, _, Ncr = label.shape
    y = label[:, 1, :] > 0.5
    am = real[:, 1, :] > real[:, 0, :]
    qt, qp = ~y, real[:, 0, :] > 0
    counts = np.stack([(am == y).sum(1), (qt & qp).sum(1), (~qt & qp).sum(1), (qt & ~qp).sum(1),
                       (y & am).sum(1), (~y & am).sum(1), (y & ~am).sum(1), y.sum(1)], 1).astype(np.int64)
    auc = np.zeros((N, 2), np.int64)
    for b in range(N):
        for col, pos_score in ((0, real[b, 0][y[b]]), (1, real[b, 1][y[b]])):
            neg = real[b, 1][~y[b]]
            auc[b, col] = 2 * (neg[None, :] < pos_score[:, None]).sum() + (neg[None, :] == pos_score[:, None]).sum()
    m = EV.metrics_from_counts(counts, quirks=True)
    assert np.isclose(m["hits"] / (N * Ncr), float(z["top_ACC"]), rtol=1e-14)
    assert np.isclose(m["prec"], float(z["prec"]), rtol=1e-14) and np.isclose(m["recall"], float(z["recall"]), rtol=1e-14)
    assert np.isclose(m["f1"], float(z["f1"]), rtol=1e-14)
    assert np.isclose(EV.auc_from_counts(counts, auc, Ncr, quirks=True), float(z["AUC"]), rtol=1e-12)
    mc = EV.metrics_from_counts(counts, quirks=False)
    assert mc["prec"] == EV.prec(label, real, False) and mc["f1"] == EV.f1(label, real, False)
    assert np.isclose(EV.auc_from_counts(counts, auc, Ncr, quirks=False), EV.AUC(label, real, False), rtol=1e-13)


def test_legacy_operator_oracle_matches_reference_model_py():
    """the oracle's map_conv (literal and closed forms) and normalize_adj against the reference's own
    model.py:335-403 executed over the shim (tests/golden/legacy_toy.npz), forward value and gradients."""
    z = np.load(os.path.join(G, "legacy_toy.npz"))
    Ra, O0, th, adj = (torch.as_tensor(z[k]) for k in ("Ra", "O", "theta", "adj"))
    mb, _, No = O0.shape
    lit = O.map_conv_dense(th.reshape(1, 2, 1, 1), Ra, O0)
    assert abs(float(lit) - float(z["loss"])) <= 1e-12 * abs(float(z["loss"]))
    x = O0.reshape(mb, No).clone().requires_grad_(True)
    t = th.clone().requires_grad_(True)
    closed = O.map_conv_closed(t, adj, x)
    assert abs(float(closed) - float(z["loss"])) <= 1e-10 * abs(float(z["loss"]))
    gx, gt = torch.autograd.grad(closed, [x, t])
    assert np.allclose(gx.numpy(), z["dO"].reshape(mb, No), rtol=1e-9, atol=1e-12)
    assert np.allclose(gt.numpy(), z["dtheta"], rtol=1e-9, atol=1e-12)
    # normalize_adj (model.py:360-367): T_1 = (2/1.5)(I - A_hat) - I  =>  A_hat = I - 0.75 (T_1 + I)
    Ahat_ref = np.eye(No) - 0.75 * (z["t_k"][:, 1] + np.eye(No))
    A = z["adj"].astype(np.float64)
    d = (A.sum(2) + float(np.float32(1e-3))) ** -0.5       # the reference's constant is a float32 array (model.py:362)
    Ahat = d[:, :, None] * np.transpose(A, (0, 2, 1)) * d[:, None, :]
    assert np.allclose(Ahat, Ahat_ref, rtol=1e-12, atol=1e-13)
    assert np.array_equal(z["t_k"][:, 0], np.broadcast_to(np.eye(No), (mb, No, No)))
