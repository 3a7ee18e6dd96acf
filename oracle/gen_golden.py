"""Generate tests/golden/*.npz by EXECUTING the reference's own files (TEST INFRASTRUCTURE).

Runs only where /root/reference is mounted (this build container):
    python -m oracle.gen_golden
It (1) writes a small synthetic data set in the reference's on-disk formats, (2) runs the
reference's unmodified utils2.read_data on it, (3) builds the reference's unmodified
model_1..4.graph2graph over oracle/tf1_shim.py (TensorFlow itself is not installable here),
evaluates forward / losses / gradients / two Adam steps on the loader's output, and (4) runs the
reference's EvaluationFuncs on random inputs.  Everything is float64 except what the reference
itself stores as float32.  The committed .npz files are what the CPU and GPU tests pin against.
"""
from __future__ import annotations

import importlib
import io
import os
import sys
import tempfile
import types
import contextlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("HDGNN_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

Ne, Nc, N, MB, STEP, REPO = 7, 4, 8, 4, 2, "toy"


def main():
    from oracle import tf1_shim
    from hdgnn_b200.synthetic import make_commits
    from hdgnn_b200.utils2 import write_dataset, read_compact
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not found: golden vectors can only be regenerated where the reference is mounted")
    tf1_shim.install()
    sys.path.insert(0, REF)
    os.makedirs(OUT, exist_ok=True)
    work = tempfile.mkdtemp(prefix="hdgnn_golden_")
    os.chdir(work)
    cb = make_commits(N, Ne, Nc, seed=42, p_edge=0.3, p_null=0.2, p_short=0.5, p_noise=0.3)
    cb.L[1] = 5; cb.L[2] = Ne                       # make sure both L < Ne and L == Ne occur in the first batch
    write_dataset(cb, REPO, STEP, root=".")
    os.makedirs(f"./Intermediate_products/{REPO}", exist_ok=True)     # the reference needs it to pre-exist (Q14)
    Ner, Ncr = Ne * (Ne - 1), Nc * (Nc - 1)

    # ---- (2) the reference's loader ---------------------------------------------------------
    ref_utils2 = importlib.import_module("utils2")
    stub = types.SimpleNamespace(Repo=REPO, Ne=Ne, Nc=Nc, Ner=Ner, Ncr=Ncr, Dr=2)
    with contextlib.redirect_stderr(io.StringIO()):
        tup = ref_utils2.read_data(stub, STEP)
    names = ["E_node_train", "E_node_test", "E_edge_train", "E_edge_test", "C_edge_train", "C_edge_test",
             "Es_data", "Et_data", "Cs_label", "Ct_label", "Esc_data", "Etc_data"]
    loader = {n: np.asarray(a) for n, a in zip(names, tup)}
    # only the rows that carry information (the reference allocates 100 leading rows, utils2.py:50-61)
    slim = dict(loader)
    for n in ("Es_data", "Et_data", "Cs_label", "Ct_label", "Esc_data", "Etc_data"):
        assert not loader[n][N:].any() or n in ("Es_data", "Et_data", "Cs_label", "Ct_label")
        slim[n] = loader[n][:N + 1]                 # one row beyond N shows what the padding rows hold
    mine = read_compact(REPO, STEP, Ne, Nc, root=".", cache=False)
    np.savez_compressed(os.path.join(OUT, "loader_toy.npz"), adj=mine.adj, x=mine.x, hmap=mine.hmap, L=mine.L, Y=mine.Y,
                        raw_adj=np.load(f"./Adjset/{REPO}/Cutting_Adjs/CAdjs_{STEP}.npy"), **slim)

    # ---- (3) the reference's models over the shim -----------------------------------------------
    feed_np = dict(E_node=loader["E_node_train"][:MB], E_edge=loader["E_edge_train"][:MB], C_edge=loader["C_edge_train"][:MB],
                   Es=loader["Es_data"][:MB], Et=loader["Et_data"][:MB], Cs=loader["Cs_label"][:MB],
                   Ct=loader["Ct_label"][:MB], Esc=loader["Esc_data"][:MB], Etc=loader["Etc_data"][:MB])
    for variant in (1, 2, 3, 4):
        tf1_shim.reset_default_graph()
        tf1_shim.set_seed(100 + variant)
        mod = importlib.import_module(f"model_{variant}")
        sess = tf1_shim.Session()
        with contextlib.redirect_stdout(io.StringIO()):
            m = mod.graph2graph(sess, Ds=1, Ne=Ne, Nc=Nc, Ner=Ner, Ncr=Ncr, Dr=2, De_e=20, De_er=20, Mini_batch=MB,
                                checkpoint_dir="./ck/", epoch=1, Ds_inter=1, Dr_inter=2, Step=STEP, Repo=REPO)
        gv = tf1_shim.global_variables()
        # zero biases hide bias paths: perturb every variable a little (same perturbation goes to the oracle)
        rng = np.random.default_rng(7 + variant)
        for v in gv:
            v.initial = v.initial + 0.05 * tf1_shim.torch.as_tensor(rng.standard_normal(tuple(v.initial.shape)))
        sess.run(tf1_shim.global_variables_initializer())
        flat0 = np.concatenate([v.value.detach().numpy().reshape(-1) for v in gv])
        feed = {m.E_node_train: feed_np["E_node"], m.E_edge_train: feed_np["E_edge"], m.C_edge_train: feed_np["C_edge"],
                m.Es: feed_np["Es"], m.Et: feed_np["Et"], m.Cs: feed_np["Cs"], m.Ct: feed_np["Ct"],
                m.Esc: feed_np["Esc"], m.Etc: feed_np["Etc"]}
        train_loss = 10 * m.loss_Hedge_mse + 0.1 * m.loss_map + m.loss_para         # model_2.py:336
        opt = tf1_shim.train.AdamOptimizer(0.0003)
        trainer = opt.minimize(train_loss)
        fetch = [m.C_edge_output2, m.C_edge_output2_logits, m.loss_Hedge_mse, m.loss_map, m.loss_para, train_loss]
        extra = {}
        if hasattr(m, "E_node_train2"):
            extra["E_node2"] = sess.run(m.E_node_train2, feed)
        if hasattr(m, "E_edge_train2"):
            extra["E_edge2"] = sess.run(m.E_edge_train2, feed)
        probs, logits, ce, loss_map, loss_para, tl, _ = sess.run(fetch + [trainer], feed)
        grad = np.concatenate([g.numpy().reshape(-1) for g in opt.last_grads])
        flat1 = np.concatenate([v.value.detach().numpy().reshape(-1) for v in gv])
        sess.run(trainer, feed)
        flat2 = np.concatenate([v.value.detach().numpy().reshape(-1) for v in gv])
        np.savez_compressed(os.path.join(OUT, f"model_{variant}_toy.npz"),
                            var_names=np.array([v.name for v in gv]), var_sizes=np.array([v.value.numel() for v in gv]),
                            var_shapes=np.array([str(tuple(v.value.shape)) for v in gv]),
                            params=flat0, probs=probs, logits=logits, ce=ce, loss_map=loss_map, loss_para=loss_para,
                            train_loss=tl, grad=grad, params_step1=flat1, params_step2=flat2, **extra)
        print(f"model_{variant}: {len(gv)} variables, {flat0.size} parameters, CE {float(ce):.6f}")
        del sys.modules[f"model_{variant}"]

    # ---- (4) the reference's evaluation functions ---------------------------------------------------
    ev = importlib.import_module("EvaluationFuncs")
    rng = np.random.default_rng(3)
    lab1 = (rng.random((5, 30)) < 0.3).astype(np.float64)
    label = np.stack([1 - lab1, lab1], 1)
    z = rng.standard_normal((5, 2, 30)) + 1.5 * (label - 0.5)
    real = np.exp(z) / np.exp(z).sum(1, keepdims=True)
    real[0, 0, :4] = 0.0; real[0, 1, :4] = 1.0          # exact zeros: the only way ceil() gives a 0 prediction
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = dict(top_ACC=ev.top_ACC(label.copy(), real.copy()), prec=ev.prec(label, real), recall=ev.recall(label, real),
                   f1=ev.f1(label, real), AUC=ev.AUC(label, real))
    np.savez_compressed(os.path.join(OUT, "eval_toy.npz"), label=label, real=real, **{k: np.float64(v) for k, v in res.items()})
    print("eval:", res)

    # ---- (5) the legacy operator: model.py's own chebyshev_polynomials / map_conv over the shim -----------------
    # (model.py:335-403).  The methods are called unbound on a stub `self`; Ra is the soft edge tensor
    # the legacy network would produce (argmax makes it a hard adjacency, model.py:337), O the node output, theta the
    # Chebyshev coefficients.  O and theta are shim Variables so that autograd gives the reference's backward
    # (the argmax blocks any gradient to Ra).
    legacy = importlib.import_module("model")
    No, mb, k = 9, 4, 2
    Nr = No * (No - 1)
    rng = np.random.default_rng(11)
    lab = (rng.random((mb, Nr)) < 0.3).astype(np.float64)
    z = rng.standard_normal((mb, 2, Nr)) + 3.0 * (np.stack([1 - lab, lab], 1) - 0.5)
    Ra = np.exp(z) / np.exp(z).sum(1, keepdims=True)                  # soft one-hots, (mb, 2, Nr)
    O0 = rng.integers(0, 10, size=(mb, 1, No)).astype(np.float64) / 3.0
    th0 = np.array([0.13, -0.21]).reshape(1, k, 1, 1)
    tf1_shim.reset_default_graph()
    stub = object.__new__(legacy.graph2graph)                         # no __init__: build_model needs the dead loaders
    stub.mini_batch_num, stub.Nr, stub.No, stub.Ds = mb, Nr, No, 1
    Ov = tf1_shim.Variable(O0, name="O_legacy")
    thv = tf1_shim.Variable(th0, name="theta_legacy")
    t_k = stub.chebyshev_polynomials(tf1_shim.constant(Ra), k)
    loss_node = stub.map_conv(thv, tf1_shim.constant(Ra), Ov, k)
    tk_val = tf1_shim._eval(t_k, {}, {}).detach().numpy()
    val = tf1_shim._eval(loss_node, {}, {})
    gO, gth = tf1_shim.torch.autograd.grad(val, [Ov.value, thv.value])
    hard = np.argmax(Ra, 1)                                             # (mb, Nr) -> dense adjacency by the pair order p(i,j)
    adj = np.zeros((mb, No, No), dtype=np.uint8)
    q = 0
    for i in range(No):
        for j in range(No):
            if i != j:
                adj[:, i, j] = hard[:, q]; q += 1
    np.savez_compressed(os.path.join(OUT, "legacy_toy.npz"), Ra=Ra, O=O0, theta=th0.reshape(k), adj=adj, t_k=tk_val,
                        loss=np.float64(val.item()), dO=gO.numpy(), dtheta=gth.numpy().reshape(k))
    print(f"legacy map_conv: loss {val.item():.6f}")


if __name__ == "__main__":
    main()
