"""GPU tests of the reference-facing surface: graph2graph.train/test and main.py on a toy data set
written in the reference's on-disk formats, checked against the CPU oracle driven the same way."""
import os

import numpy as np
import pytest
import torch

from hdgnn_b200.synthetic import make_commits, CommitBatch
from hdgnn_b200.utils2 import write_dataset, read_compact, split_half, edge_onehot
from oracle import hdgnn_oracle as O

pytestmark = pytest.mark.gpu
Ne, Nc, N, MB = 12, 5, 16, 4


def _dataset(tmp_path):
    cb = make_commits(N, Ne, Nc, seed=8, p_edge=0.2, p_short=0.4, p_noise=0.2)
    write_dataset(cb, "toy", 2, root=str(tmp_path))
    return read_compact("toy", 2, Ne, Nc, root=str(tmp_path))


@pytest.mark.parametrize("variant", [2, 4])
def test_train_loop_matches_oracle_training(tmp_path, variant):
    from hdgnn_b200.model import graph2graph, truncated_normal_init
    cb = _dataset(tmp_path)
    train, test = split_half(cb)
    model = graph2graph(None, Ne=Ne, Nc=Nc, Mini_batch=MB, epoch=3, Step=2, Repo="toy", variant=variant, seed=11,
                        checkpoint_dir=str(tmp_path / "ck"))
    logs = []
    hist = model.train(None, root=str(tmp_path), log=logs.append)
    # the same loop on the CPU oracle: same init, batches of MB with the maps of the first MB commits (quirk Q2)
    flat = truncated_normal_init(variant, 11).double()
    m = torch.zeros_like(flat); v = torch.zeros_like(flat)
    t = 0
    for ep in range(3):
        ces = []
        for j in range(train.B // MB):
            b = train.slice(j * MB, (j + 1) * MB)
            _, ce, _, g, _ = O.train_loss_and_grad(variant, flat, b.adj, b.x, train.hmap[:MB], train.L[:MB], b.Y)
            t += 1
            flat, m, v = O.tf_adam_step(flat, g, m, v, t)
            ces.append(float(ce))
        assert abs(hist[ep]["hedge_loss"] - np.mean(ces)) < 2e-4 * np.mean(ces), (ep, hist[ep], np.mean(ces))
    got = model.params.cpu().double()
    assert float((got - flat).abs().max() / flat.abs().max()) < 1e-4
    assert len(logs) == 4 and logs[0].startswith("Epoch 1 acc: ") and " Hedge loss: " in logs[0] and " theta: " in logs[0]
    res = tmp_path / "outputSelf" / "toy" / f"model_{variant}" / "2" / "result_2.npy"
    assert res.exists() and len(res.read_text().splitlines()) == 3
    ck = tmp_path / "ck" / "toy" / f"model_{variant}" / "2"
    assert (ck / "checkpoint").exists() and (ck / "g2g.model-4.npz").exists()
    # test(): loads the checkpoint, writes the two .npy files, prints the metric lines
    logs2 = []
    out, probs = model.test(None, root=str(tmp_path), log=logs2.append)
    assert any(l.startswith(" [*] Load SUCCESS") for l in logs2)
    assert [l.split(":")[0] for l in logs2[-6:]] == ["topol_acc", "prec", "recall", "F1-score", "AUC-score", "test time"]
    d = tmp_path / "outputSelf" / "toy" / f"model_{variant}" / "2"
    pt = np.load(d / f"C_edge_t{Ne}.npy"); py = np.load(d / f"C_edge_y{Ne}.npy")
    assert pt.shape == (test.B, 2, Nc * (Nc - 1)) and np.array_equal(py, edge_onehot(test.Y))
    # the metric lines come from device counters: identical to the array functions pinned to EvaluationFuncs.py
    from hdgnn_b200 import EvaluationFuncs as EV
    assert out["topol_acc"] == EV.top_ACC(py, pt)
    assert out["prec"] == EV.prec(py, pt) and out["recall"] == EV.recall(py, pt) and out["f1"] == EV.f1(py, pt)
    try:
        assert np.isclose(out["auc"], EV.AUC(py, pt), rtol=1e-13)
    except ZeroDivisionError:
        assert np.isnan(out["auc"])
    # and the per-epoch accuracy of train() is a ratio of integers over the training half
    assert all(0.0 <= h["acc"] <= 1.0 and (h["acc"] * train.B * Nc * (Nc - 1)) % 1 < 1e-6 for h in hist)
    # inference parity on the test half (maps of the first MB TRAIN commits, quirk Q2)
    P = O.unflatten(flat, variant)
    for j in range(test.B // MB):
        b = test.slice(j * MB, (j + 1) * MB)
        ref = O.forward_closed(variant, P, b.adj, b.x, train.hmap[:MB], train.L[:MB], b.Y)["probs"].numpy()
        assert np.abs(pt[j * MB:(j + 1) * MB] - ref).max() < 1e-4
    model.engine.close()


def test_device_loader_equals_host_loader(tmp_path):
    """graph2graph._load runs the array half of the loader on the GPU (hdgnn_compact_from_raw)."""
    cb = make_commits(N, Ne, Nc, seed=8, p_edge=0.2, p_short=0.4, p_noise=0.2)
    write_dataset(cb, "toy", 2, root=str(tmp_path))
    host = read_compact("toy", 2, Ne, Nc, root=str(tmp_path), cache=False)
    dev = read_compact("toy", 2, Ne, Nc, root=str(tmp_path), cache=False, device=torch.device("cuda", 0))
    for name in ("adj", "x", "hmap", "L", "Y"):
        a, b = getattr(host, name), getattr(dev, name)
        assert a.dtype == b.dtype and np.array_equal(a, b), name
    assert np.array_equal(host.adj, cb.adj) and np.array_equal(host.Y, cb.Y)


def test_main_cli_train_and_test(tmp_path, capsys):
    import main as cli
    _dataset(tmp_path)
    argv = ["--Ne", str(Ne), "--Nc", str(Nc), "--Ner", str(Ne * (Ne - 1)), "--Ncr", str(Nc * (Nc - 1)), "--Mini_batch", str(MB),
            "--epoch", "2", "--Repo", "toy", "--Step", "2", "--root", str(tmp_path), "--seed", "3"]
    res = cli.main(argv + ["--Type", "train"])
    assert len(res) == 3 and all(len(r) == 2 for r in res)       # the reference loops over its 3 presets (main.py:21)
    res = cli.main(argv + ["--Type", "test"])
    out = capsys.readouterr().out
    assert out.count("topol_acc: ") == 3 and out.count(" [*] Load SUCCESS") == 3
    with pytest.raises(ValueError):
        cli.main(argv + ["--Ner", "5"])


def test_constructor_rejects_unsupported_dimensions():
    from hdgnn_b200.model import graph2graph
    with pytest.raises(ValueError):
        graph2graph(None, Ne=10, Nc=5, De_e=16)
    with pytest.raises(ValueError):
        graph2graph(None, Ne=10, Nc=5, Ner=91)
