"""CPU restatement of the HD-GNN hot path (TEST INFRASTRUCTURE -- not product code).

Two independent restatements of the reference network, both in PyTorch on the CPU
(fp64 or fp32) with autograd for the backward pass:

* ``forward_dense``  -- a function-by-function transcription of the reference's TF1
  graph that really builds the one-hot incidence tensors ``Es/Et/Cs/Ct/Esc/Etc`` and
  multiplies by them (reference: model_2.py:141-324, model_4.py:206-304,
  model_1.py:72-80, model_3.py:82-102; inputs per utils2.py:29-137).
* ``forward_closed`` -- the same arithmetic in index form (gathers, row/column sums,
  segmented sums) on the compact inputs; no one-hot tensors.  This is also the CPU
  baseline timed by ``bench.py``.

The two are checked against each other and against golden vectors obtained by
running the reference's own model files over ``oracle/tf1_shim.py``
(tests/test_oracle.py).  Parity status: see ``oracle/__init__.py`` -- the TF runtime
itself is unavailable, so parity is unpinned at the TF-kernel boundary.

Parameter order = TF variable-creation order of each ``build_model``
(model_2.py:86-121 etc.); names below give ``<tf scope>/<tf name>``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

H = 20      # h_size, model_2.py:163
DS = 1      # --Ds   main.py:26
DR = 2      # --Dr   main.py:28
DE = 20     # --De_e / --De_er  main.py:31-32

# (short name, tf scope/name, shape)
_ENT = [
    ("ent_w1", "phi_E_O1/r1_w1o", (2 * DS + DR, H)),   # model_2.py:167
    ("ent_b1", "phi_E_O1/r1_b1o", (H,)),
    ("ent_w5", "phi_E_O1/r1_w5o", (H, DE)),            # model_2.py:172
    ("ent_b5", "phi_E_O1/r1_b5o", (DE,)),
    ("nod_w1", "phi_U_O1/o1_w1o", (DS + DE, H)),       # model_2.py:196
    ("nod_b1", "phi_U_O1/o1_b1o", (H,)),
    ("nod_w2", "phi_U_O1/o1_w2o", (H, DS)),            # model_2.py:200
    ("nod_b2", "phi_U_O1/o1_b2o", (DS,)),
]
_EDGE = [
    ("edg_w11", "phi_E_R1/r1_w1r1", (DS, H)),          # model_4.py:219
    ("edg_w12", "phi_E_R1/r1_w1r2", (DR, H)),          # model_4.py:220
    ("edg_b1", "phi_E_R1/r1_b1r", (H,)),
    ("edg_w2", "phi_E_R1/r1_w2r", (H, DE)),
    ("edg_b2", "phi_E_R1/r1_b2r", (DE,)),
    ("eup_w1", "phi_U_R1/o1_w1r", (DE + DR, H)),       # model_4.py:292
    ("eup_b1", "phi_U_R1/o1_b1r", (H,)),
    ("eup_w2", "phi_U_R1/o1_w2r", (H, DR)),
    ("eup_b2", "phi_U_R1/o1_b2r", (DR,)),
]
_HUNK = [
    ("hnk_w1", "mlp_hunk_B2/w1", (2 * (2 * DS + DR) + DR, H)),   # (10,20) model_2.py:257
    ("hnk_b1", "mlp_hunk_B2/b1", (H,)),
    ("hnk_w2", "mlp_hunk_B2/r1_w2r", (H, DE)),
    ("hnk_b2", "mlp_hunk_B2/b2", (DE,)),
    ("scr_w1", "phi_U_R1*/C_edge_w1", (DE + DR, H)),             # (22,20) model_2.py:311
    ("scr_b1", "phi_U_R1*/C_edge_b1", (H,)),
    ("scr_w2", "phi_U_R1*/o1_w2r", (H, DR)),
    ("scr_b2", "phi_U_R1*/o1_b2r", (DR,)),
]
_THETA = [
    ("theta1", "map_conv/map_theta1", (2,)),           # (1,k,1,1), model_2.py:329
    ("theta2", "map_conv/map_theta2", (2,)),
]
_BLOCKS = {1: _HUNK + _THETA, 2: _ENT + _HUNK + _THETA,
           3: _EDGE + _HUNK + _THETA, 4: _ENT + _EDGE + _HUNK + _THETA}


def param_spec(variant: int) -> List[Tuple[str, str, Tuple[int, ...]]]:
    return list(_BLOCKS[variant])


def param_count(variant: int) -> int:
    return sum(int(np.prod(s)) for _, _, s in _BLOCKS[variant])


def unflatten(flat: torch.Tensor, variant: int) -> Dict[str, torch.Tensor]:
    out, off = {}, 0
    for name, _, shape in _BLOCKS[variant]:
        n = int(np.prod(shape))
        out[name] = flat[off:off + n].reshape(shape)
        off += n
    assert off == flat.numel()
    return out


def init_params(variant: int, seed: int = 1234, dtype=torch.float32) -> torch.Tensor:
    """tf.truncated_normal(stddev=0.1) weights (resample beyond 2 sigma), zero biases
    (model_2.py:167-201).  The reference is unseeded; parity always injects weights."""
    g = torch.Generator().manual_seed(seed)
    parts = []
    for name, _, shape in _BLOCKS[variant]:
        n = int(np.prod(shape))
        if "_b" in name:          # biases are tf.zeros
            parts.append(torch.zeros(n, dtype=torch.float64))
            continue
        v = torch.randn(n, generator=g, dtype=torch.float64)
        bad = v.abs() > 2
        while bad.any():
            v[bad] = torch.randn(int(bad.sum()), generator=g, dtype=torch.float64)
            bad = v.abs() > 2
        parts.append(0.1 * v)
    return torch.cat(parts).to(dtype)


# ----------------------------------------------------------------------------
# index helpers (utils2.py:63-83, 85-106, 111-137)
# ----------------------------------------------------------------------------
def pair_index(n: int) -> Tuple[np.ndarray, np.ndarray]:
    """Row-major enumeration of ordered pairs (i, j), i != j (utils2.py:69-83)."""
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    keep = ii != jj
    return ii[keep].astype(np.int64), jj[keep].astype(np.int64)


def dense_inputs(adj, x, hmap, L, Y, dtype=torch.float64):
    """Compact inputs -> the nine dense feeds of model_2.py:54-82, built exactly as
    utils2.py:63-137 builds them (one-hot over pairs; Esc/Etc with the LOCAL L*(L-1)
    counter of utils2.py:123-137, quirk Q3)."""
    adj = np.asarray(adj); Y = np.asarray(Y); hmap = np.asarray(hmap); L = np.asarray(L)
    B, Ne, _ = adj.shape
    Nc = Y.shape[1]
    Ner, Ncr = Ne * (Ne - 1), Nc * (Nc - 1)
    ei, ej = pair_index(Ne)
    ci, cj = pair_index(Nc)
    Es = np.zeros((B, Ne, Ner)); Et = np.zeros((B, Ne, Ner))
    Es[:, ei, np.arange(Ner)] = 1.0
    Et[:, ej, np.arange(Ner)] = 1.0
    Cs = np.zeros((B, Nc, Ncr)); Ct = np.zeros((B, Nc, Ncr))
    Cs[:, ci, np.arange(Ncr)] = 1.0
    Ct[:, cj, np.arange(Ncr)] = 1.0
    E_edge = np.zeros((B, 2, Ner)); C_edge = np.zeros((B, 2, Ncr))
    bb = np.arange(B)[:, None]
    E_edge[bb, adj[:, ei, ej].astype(np.int64), np.arange(Ner)[None, :]] = 1.0
    C_edge[bb, Y[:, ci, cj].astype(np.int64), np.arange(Ncr)[None, :]] = 1.0
    Esc = np.zeros((B, Nc, Ner)); Etc = np.zeros((B, Nc, Ner))
    for b in range(B):
        Lb = int(L[b])
        cnt2 = 0
        for i in range(Lb):
            for j in range(Lb):
                if i != j:
                    hs, ht = int(hmap[b, i]), int(hmap[b, j])
                    if 0 <= hs < Nc:
                        Esc[b, hs, cnt2] = 1.0
                    if 0 <= ht < Nc:
                        Etc[b, ht, cnt2] = 1.0
                    cnt2 += 1
    E_node = np.asarray(x, dtype=np.float64).reshape(B, 1, Ne)
    t = lambda a: torch.as_tensor(a, dtype=dtype)
    return dict(E_node=t(E_node), Es=t(Es), Et=t(Et), E_edge=t(E_edge), Cs=t(Cs), Ct=t(Ct),
                C_edge=t(C_edge), Esc=t(Esc), Etc=t(Etc))


# ----------------------------------------------------------------------------
# dense transcription (one function per reference function)
# ----------------------------------------------------------------------------
def _T(a):
    return a.transpose(1, 2)


def marshalling_B1(O, Es, Et, Ra):                       # model_2.py:141-144
    return torch.cat([O @ Es, O @ Et, Ra], 1)


def marshalling_B2(B_1, Esc, Etc, Cs, Ct, C_edge):        # model_2.py:146-159
    B_t = _T(B_1)
    neighbors = Esc @ B_t + Etc @ B_t
    hs = _T(_T(Cs) @ neighbors)
    ht = _T(_T(Ct) @ neighbors)
    return torch.cat([hs, ht, C_edge], 1)


def mlp_entity_B1(B, P):                                  # model_2.py:161-179
    mb, _, Ner = B.shape
    Bt = _T(B).reshape(mb * Ner, 2 * DS + DR)
    h1 = torch.relu(Bt @ P["ent_w1"] + P["ent_b1"])
    h5 = h1 @ P["ent_w5"] + P["ent_b5"]
    return _T(h5.reshape(mb, Ner, DE))


def agg_entity_B1(E, Es, Et, O):                          # model_2.py:181-188
    E_bar = E @ _T(Es) + E @ _T(Et)
    return torch.cat([O, E_bar], 1)


def mlp2_entity_B1(C, P):                                 # model_2.py:190-205
    mb, _, Ne = C.shape
    Ct_ = _T(C).reshape(mb * Ne, DS + DE)
    h1 = torch.relu(Ct_ @ P["nod_w1"] + P["nod_b1"])
    h2 = torch.relu(h1 @ P["nod_w2"] + P["nod_b2"])
    return _T(h2.reshape(mb, Ne, DS))


def mlp_entityedge_B1(B, Es, Et, P):                      # model_4.py:206-243
    mb, _, Ner = B.shape
    Bt = _T(B).reshape(mb * Ner, 2 * DS + DR)
    w1 = torch.cat([P["edg_w11"], P["edg_w11"], P["edg_w12"]], 0)
    h1 = torch.relu(Bt @ w1 + P["edg_b1"])
    h2 = _T((h1 @ P["edg_w2"] + P["edg_b2"]).reshape(mb, Ner, DE))
    bar1 = h2 @ _T(Es)
    bar2 = h2 @ _T(Et)
    return bar1 @ Es + bar2 @ Et


def mlp2_entityedge_B1(C_R, P):                           # model_4.py:286-304
    mb, _, Ner = C_R.shape
    Ct_ = _T(C_R).reshape(mb * Ner, DE + DR)
    h1 = torch.relu(Ct_ @ P["eup_w1"] + P["eup_b1"])
    logits = _T((h1 @ P["eup_w2"] + P["eup_b2"]).reshape(mb, Ner, DR))
    return torch.softmax(logits, 1), logits


def mlp_hunk_B2(B2, Cs, Ct, P):                           # model_2.py:245-277
    mb, d, Ncr = B2.shape
    Bt = _T(B2).reshape(mb * Ncr, d)
    h1 = torch.relu(Bt @ P["hnk_w1"] + P["hnk_b1"])
    h2 = _T((h1 @ P["hnk_w2"] + P["hnk_b2"]).reshape(mb, Ncr, DE))
    bar1 = h2 @ _T(Cs)
    bar2 = h2 @ _T(Ct)
    return bar1 @ Cs + bar2 @ Ct


def mlp_hunkedge_B2(HRa, P):                              # model_2.py:304-324
    mb, d, Ncr = HRa.shape
    Ct_ = _T(HRa).reshape(mb * Ncr, d)
    h1 = torch.relu(Ct_ @ P["scr_w1"] + P["scr_b1"])
    logits = _T((h1 @ P["scr_w2"] + P["scr_b2"]).reshape(mb, Ncr, DR))
    return torch.softmax(logits, 1), logits


def forward_dense(variant: int, P: Dict[str, torch.Tensor], D: Dict[str, torch.Tensor]):
    """Returns dict(probs (B,2,Ncr), logits (B,2,Ncr), ce scalar, plus intermediates)."""
    out = {}
    B_1 = marshalling_B1(D["E_node"], D["Es"], D["Et"], D["E_edge"])           # model_2.py:86
    node2, edge2 = D["E_node"], D["E_edge"]
    if variant in (2, 4):                                                       # model_2.py:89-91
        e = mlp_entity_B1(B_1, P)
        a = agg_entity_B1(e, D["Es"], D["Et"], D["E_node"])
        node2 = mlp2_entity_B1(a, P)
        out["E_node2"] = node2
    if variant in (3, 4):                                                       # model_4.py:92-94
        eff = mlp_entityedge_B1(B_1, D["Es"], D["Et"], P)
        a_E_edge = torch.cat([D["E_edge"], eff], 1)
        e2, _ = mlp2_entityedge_B1(a_E_edge, P)
        out["E_edge2"] = e2
        if variant == 4:
            edge2 = e2                                                          # model_4.py:97
    if variant in (2, 4):
        B_2 = marshalling_B1(node2, D["Es"], D["Et"], edge2)                    # model_2.py:94
    else:
        B_2 = B_1                                                               # model_1.py:76, model_3.py:97
    B_3 = marshalling_B2(B_2, D["Esc"], D["Etc"], D["Cs"], D["Ct"], D["C_edge"])
    out["B_3"] = B_3
    C_out = mlp_hunk_B2(B_3, D["Cs"], D["Ct"], P)                               # model_2.py:103
    a_C = torch.cat([D["C_edge"], C_out], 1)                                    # model_2.py:104,279-281
    probs, logits = mlp_hunkedge_B2(a_C, P)                                     # model_2.py:105
    ce = -(D["C_edge"] * torch.log_softmax(logits, 1)).sum(1).mean()            # model_2.py:115-118
    out.update(probs=probs, logits=logits, ce=ce)
    return out


# ----------------------------------------------------------------------------
# closed (index) form
# ----------------------------------------------------------------------------
def _offdiag(n, dtype):
    return (1.0 - torch.eye(n, dtype=dtype))


def _pair_mlp_sums(rowf, colf, lab, w_row, w_col, w_lab, b1, w2, b2):
    """For an N x N off-diagonal grid: per-pair 2-layer MLP on
    [rowf_i, colf_j, 1-lab_ij, lab_ij]; returns (e (B,N,N,DE) masked, h1)."""
    pre = (rowf @ w_row)[:, :, None, :] + (colf @ w_col)[:, None, :, :] \
        + (1.0 - lab)[..., None] * w_lab[0] + lab[..., None] * w_lab[1] + b1
    h1 = torch.relu(pre)
    e = h1 @ w2 + b2
    mask = _offdiag(lab.shape[1], lab.dtype)[None, :, :, None]
    return e * mask


def forward_closed(variant: int, P: Dict[str, torch.Tensor], adj, x, hmap, L, Y):
    """adj (B,Ne,Ne) {0,1} zero diagonal; x (B,Ne); hmap (B,Ne) int (-1 = none);
    L (B,) int; Y (B,Nc,Nc) {0,1} zero diagonal.  All float tensors share P's dtype."""
    dt = P["hnk_w1"].dtype
    A = torch.as_tensor(np.asarray(adj), dtype=dt)
    Yl = torch.as_tensor(np.asarray(Y), dtype=dt)
    xe = torch.as_tensor(np.asarray(x), dtype=dt)
    hmap = np.asarray(hmap); L = np.asarray(L)
    B, Ne, _ = A.shape
    Nc = Yl.shape[1]
    out = {}
    node2 = xe
    edge1 = A                       # P(edge)=1 channel; channel 0 is 1-A
    edge0 = 1.0 - A
    if variant in (2, 4):
        w1 = P["ent_w1"]
        e = _pair_mlp_sums(xe[..., None], xe[..., None], A, w1[0:1], w1[1:2], w1[2:4],
                           P["ent_b1"], P["ent_w5"], P["ent_b5"])
        Ebar = e.sum(2) + e.sum(1)                                      # model_2.py:186
        z = torch.relu(torch.cat([xe[..., None], Ebar], -1) @ P["nod_w1"] + P["nod_b1"])
        node2 = torch.relu(z @ P["nod_w2"] + P["nod_b2"])[..., 0]       # (B,Ne)
        out["E_node2"] = node2[:, None, :]
    if variant in (3, 4):
        w12 = P["edg_w12"]
        eh = _pair_mlp_sums(xe[..., None], xe[..., None], A, P["edg_w11"], P["edg_w11"], w12,
                            P["edg_b1"], P["edg_w2"], P["edg_b2"])
        r = eh.sum(2); c = eh.sum(1)                                    # bar1, bar2
        eff = r[:, :, None, :] + c[:, None, :, :]                       # model_4.py:240
        g1 = P["eup_w1"]
        pre = (1.0 - A)[..., None] * g1[0] + A[..., None] * g1[1] + eff @ g1[2:] + P["eup_b1"]
        lg = torch.relu(pre) @ P["eup_w2"] + P["eup_b2"]                # (B,Ne,Ne,2)
        sm = torch.softmax(lg, -1)
        ei, ej = pair_index(Ne)
        out["E_edge2"] = sm[:, ei, ej, :].transpose(1, 2)               # (B,2,Ner)
        if variant == 4:
            edge0, edge1 = sm[..., 0], sm[..., 1]
    # B_2 on the Ne grid, flattened in pair order
    ei, ej = pair_index(Ne)
    B2 = torch.stack([node2[:, ei], node2[:, ej], edge0[:, ei, ej], edge1[:, ei, ej]], -1)  # (B,Ner,4)
    # pooling, utils2.py:111-137 + model_2.py:147-150
    nb = torch.zeros(B, Nc, 4, dtype=dt)
    rows = []
    for b in range(B):
        Lb = int(L[b])
        li, lj = pair_index(Lb)
        hs = hmap[b, li]; ht = hmap[b, lj]
        q = np.arange(Lb * (Lb - 1))
        vs = (hs >= 0) & (hs < Nc); vt = (ht >= 0) & (ht < Nc)
        acc = torch.zeros(Nc, 4, dtype=dt)
        acc = acc.index_add(0, torch.as_tensor(hs[vs], dtype=torch.long), B2[b, torch.as_tensor(q[vs])])
        acc = acc.index_add(0, torch.as_tensor(ht[vt], dtype=torch.long), B2[b, torch.as_tensor(q[vt])])
        rows.append(acc)
    nb = torch.stack(rows, 0)
    out["nb"] = nb
    v1 = P["hnk_w1"]
    g = _pair_mlp_sums(nb, nb, Yl, v1[0:4], v1[4:8], v1[8:10], P["hnk_b1"], P["hnk_w2"], P["hnk_b2"])
    r = g.sum(2); c = g.sum(1)
    eff = r[:, :, None, :] + c[:, None, :, :]                           # model_2.py:272-275
    s1 = P["scr_w1"]
    pre = (1.0 - Yl)[..., None] * s1[0] + Yl[..., None] * s1[1] + eff @ s1[2:] + P["scr_b1"]
    lg = torch.relu(pre) @ P["scr_w2"] + P["scr_b2"]                    # (B,Nc,Nc,2)
    ci, cj = pair_index(Nc)
    logits = lg[:, ci, cj, :].transpose(1, 2)                           # (B,2,Ncr)
    lab = Yl[:, ci, cj]
    onehot = torch.stack([1.0 - lab, lab], 1)
    ce = -(onehot * torch.log_softmax(logits, 1)).sum(1).mean()
    out.update(probs=torch.softmax(logits, 1), logits=logits, ce=ce)
    return out


# ----------------------------------------------------------------------------
# loss, gradient, TF1 Adam
# ----------------------------------------------------------------------------
def reg_loss(flat: torch.Tensor, variant: int):
    """0.1*loss_map + loss_para  (model_2.py:121-130, 326-336):
    loss_map = 0.01(|theta2| + |theta1|), loss_para = sum_v 0.001 * l2_loss(v) over ALL
    global variables (weights, biases, thetas), l2_loss = sum(v^2)/2."""
    P = unflatten(flat, variant)
    loss_map = 0.01 * (torch.sqrt((P["theta2"] ** 2).sum()) + torch.sqrt((P["theta1"] ** 2).sum()))
    loss_para = 0.001 * 0.5 * (flat ** 2).sum()
    return loss_map, loss_para


def train_loss_and_grad(variant, flat, adj, x, hmap, L, Y, dense=False):
    """train_loss = 10*CE + 0.1*loss_map + loss_para (model_2.py:336); returns
    (loss, ce, loss_map, grad_flat, forward outputs)."""
    flat = flat.detach().clone().requires_grad_(True)
    P = unflatten(flat, variant)
    if dense:
        out = forward_dense(variant, P, dense_inputs(adj, x, hmap, L, Y, dtype=flat.dtype))
    else:
        out = forward_closed(variant, P, adj, x, hmap, L, Y)
    loss_map, loss_para = reg_loss(flat, variant)
    loss = 10.0 * out["ce"] + 0.1 * loss_map + loss_para
    (grad,) = torch.autograd.grad(loss, flat)
    return loss.detach(), out["ce"].detach(), loss_map.detach(), grad, {k: v.detach() for k, v in out.items()}


def tf_adam_step(p, g, m, v, t, lr=3e-4, b1=0.9, b2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer (model_2.py:337): lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    m,v un-corrected; p -= lr_t * m / (sqrt(v) + eps).  Works on numpy or torch."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    p = p - lr_t * m / (v ** 0.5 + eps)
    return p, m, v


# ----------------------------------------------------------------------------
# legacy operator: normalize_adj + Chebyshev map_conv (model.py:335-403)
# ----------------------------------------------------------------------------
def map_conv_dense(theta, Ra, O, k=2):
    """Literal: Ra (mb,2,Nr) edge one-hots, O (mb,Ds=1,No), theta (1,k,1,1)."""
    mb, _, Nr = Ra.shape
    No = O.shape[2]
    dt = O.dtype
    ra = torch.argmax(Ra, 1).to(dt).reshape(mb, 1, Nr)                     # model.py:337-339
    S_rows = []
    for i in range(No):                                                    # model.py:341-348
        s = torch.zeros(mb, 1, Nr, dtype=dt)
        s[:, :, i * (No - 1):(i + 1) * (No - 1)] = 1
        S_rows.append(s * ra)
    S = torch.cat(S_rows, 1)
    T = torch.zeros(mb, Nr, No, dtype=dt)                                  # model.py:349-356
    for i in range(No):
        t = torch.zeros(mb, No - 1, No, dtype=dt)
        for j in range(0, i):
            t[:, j, j] = 1
        for j in range(i, No - 1):
            t[:, j, j + 1] = 1
        T[:, i * (No - 1):(i + 1) * (No - 1), :] = t
    adj = S @ T
    rowsum = adj.sum(2) + float(np.float32(0.001))                         # model.py:362 (the constant is a float32 array)
    d = rowsum ** -0.5
    dm = torch.diag_embed(d)
    a = (adj @ dm).transpose(1, 2)
    adj_n = a @ dm                                                         # model.py:366-367
    I = torch.eye(No, dtype=dt).expand(mb, No, No)
    lap = I - adj_n
    scaled = (2.0 / 1.5) * lap - I                                         # model.py:375-378
    t_k = torch.stack([I, scaled], 1)                                      # k == 2: loop :387 empty
    assert k == 2
    th = torch.softmax(theta.reshape(1, k, 1, 1), 1)                       # model.py:397
    O_copy = O.transpose(1, 2).reshape(mb, 1, No, DS).expand(mb, k, No, DS)
    conv = (th * t_k) @ O_copy
    loss = O @ conv.sum(1)                                                 # (mb,Ds,Ds)
    return (loss ** 2).mean()


def map_conv_closed(theta, adj, x, self_loop=False, eps=float(np.float32(1e-3)), lam_max=1.5):
    """Index form: adj (mb,No,No) zero diagonal, x (mb,No).  A_hat = D^-1/2 A^T D^-1/2,
    D = rowsum(A)+eps (reference: no self loop, model.py:360-367)."""
    dt = x.dtype
    A = adj.to(dt)
    No = A.shape[1]
    if self_loop:
        A = A + torch.eye(No, dtype=dt)
    d = (A.sum(2) + eps) ** -0.5
    Ahat = d[:, :, None] * A.transpose(1, 2) * d[:, None, :]
    th = torch.softmax(theta.reshape(-1), 0)
    Ax = (Ahat @ x[..., None])[..., 0]
    Lx = (2.0 / lam_max) * (x - Ax) - x
    y = th[0] * x + th[1] * Lx
    return (((x * y).sum(1)) ** 2).mean()
