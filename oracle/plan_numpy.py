"""Kernel-plan emulation (TEST INFRASTRUCTURE).

NumPy fp64 restatement of the *decomposition the CUDA kernels use* (DESIGN.md section
"Kernels"): every per-pair layer is  relu(P_i + Q_j + l_ij * D)  followed either by
row/column sums (the second, linear layer is applied to the sums) or by the 20->2 head;
the backward pass is written out by hand in the same form.  It exists to (1) prove the
hand-derived backward against autograd of ``hdgnn_oracle`` on the CPU and (2) give the
GPU tests named intermediates (RS1, nb, RS3, ...) to compare workspace buffers against.

Reference lines: forward model_2.py:141-324 / model_4.py:206-304, loss model_2.py:115-130,
336; pooling indices utils2.py:111-137.
"""
from __future__ import annotations

import numpy as np

from . import hdgnn_oracle as O

H = 20


def _np(P):
    return {k: v.detach().double().numpy() for k, v in P.items()}


def relu(a):
    return np.maximum(a, 0.0)


def pairsum_fwd(P, Q, D, lab):
    """P,Q (B,N,H); D (H,); lab (B,N,N) -> RS (B,N,H), CS (B,N,H) over i != j."""
    B, N, _ = P.shape
    t = P[:, :, None, :] + Q[:, None, :, :] + lab[..., None] * D
    h = relu(t) * (1.0 - np.eye(N))[None, :, :, None]
    return h.sum(2), h.sum(1)


def pairsum_bwd(P, Q, D, lab, GR, GC):
    """v_ij = (GR_i + GC_j) * [pre_ij > 0]; returns RSd, CSd, LS (label-1 sum, (B,H))."""
    N = P.shape[1]
    t = P[:, :, None, :] + Q[:, None, :, :] + lab[..., None] * D
    v = (GR[:, :, None, :] + GC[:, None, :, :]) * (t > 0) * (1.0 - np.eye(N))[None, :, :, None]
    return v.sum(2), v.sum(1), (v * lab[..., None]).sum((1, 2))


def unflat(q, n):
    i = q // (n - 1)
    r = q % (n - 1)
    return i, r + (r >= i)


def score_pass1(PR, PC, Dg, W2, b2, lab):
    t = PR[:, :, None, :] + PC[:, None, :, :] + lab[..., None] * Dg
    h = relu(t)
    lg = h @ W2 + b2                                    # (B,N,N,2)
    d = lg[..., 1] - lg[..., 0]
    e = np.exp(-np.abs(d))
    p1 = np.where(d >= 0, 1.0 / (1.0 + e), e / (1.0 + e))
    p0 = np.where(d >= 0, e / (1.0 + e), 1.0 / (1.0 + e))
    return t, h, lg, p0, p1


def score_pass2(t, h, delta, lab):
    """delta (B,N,N) with zero diagonal.  Returns RSm, CSm, LSm, HS (un-scaled by gamma)."""
    m = (t > 0) * delta[..., None]
    return m.sum(2), m.sum(1), (m * lab[..., None]).sum((1, 2)), (h * delta[..., None]).sum((1, 2))


def head_node_fwd(RS, CS, W2, b2, G1, gb1, N):
    """r,c = linear layer on sums; PR/PC tables of the 22->20 layer (model_2.py:272-275, 311-315)."""
    r = RS @ W2 + (N - 1) * b2
    c = CS @ W2 + (N - 1) * b2
    PR = r @ G1[2:] + gb1 + G1[0]
    PC = c @ G1[2:]
    return r, c, PR, PC


def head_node_bwd(RSm, CSm, LSm, HS, dsum, RS3, CS3, r, c, W2, G1, G2, N):
    """Backward of head_node_fwd + second layer.  Returns grads dict and GR, GC tables."""
    gam = G2[:, 1] - G2[:, 0]
    RS4, CS4, LS4 = RSm * gam, CSm * gam, LSm * gam
    g = {}
    g["w2_head"] = np.stack([-HS, HS], -1).sum(0)                        # (H,2)
    g["b2_head"] = np.stack([-dsum, dsum], -1).sum(0)
    db1 = RS4.sum(1)                                                     # (B,H)
    g["b1_head"] = db1.sum(0)
    dG1 = np.zeros_like(G1)
    dG1[1] = LS4.sum(0)
    dG1[0] = db1.sum(0) - LS4.sum(0)
    dG1[2:] = np.einsum("bnm,bnk->mk", r, RS4) + np.einsum("bnm,bnk->mk", c, CS4)
    g["w1_head"] = dG1
    dr = RS4 @ G1[2:].T
    dc = CS4 @ G1[2:].T
    g["w2_pair"] = np.einsum("bna,bnm->am", RS3, dr) + np.einsum("bna,bnm->am", CS3, dc)
    g["b2_pair"] = (N - 1) * (dr.sum((0, 1)) + dc.sum((0, 1)))
    return g, dr @ W2.T, dc @ W2.T


def pool_fwd(B2grid, hmap, L, Nc):
    """B2grid (B,Ne,Ne,4) -> nb (B,Nc,4) with the local L(L-1) enumeration (Q3)."""
    B, Ne = B2grid.shape[:2]
    nb = np.zeros((B, Nc, 4))
    for b in range(B):
        Lb = int(L[b])
        li, lj = O.pair_index(Lb)
        q = np.arange(Lb * (Lb - 1))
        gi, gj = unflat(q, Ne)
        val = B2grid[b, gi, gj]
        for side in (li, lj):
            hh = hmap[b, side]
            ok = (hh >= 0) & (hh < Nc)
            np.add.at(nb[b], hh[ok], val[ok])
    return nb


def pool_bwd(dnb, hmap, L, Ne):
    """-> dB2grid (B,Ne,Ne,4)."""
    B, Nc, _ = dnb.shape
    out = np.zeros((B, Ne, Ne, 4))
    for b in range(B):
        Lb = int(L[b])
        li, lj = O.pair_index(Lb)
        q = np.arange(Lb * (Lb - 1))
        gi, gj = unflat(q, Ne)
        w = np.zeros((q.size, 4))
        for side in (li, lj):
            hh = hmap[b, side]
            ok = (hh >= 0) & (hh < Nc)
            w[ok] += dnb[b, hh[ok]]
        out[b, gi, gj] = w
    return out


def train_step_plan(variant, flat, adj, x, hmap, L, Y):
    """Returns dict(logits, probs, ce, grad (flat, incl. regularisers), intermediates)."""
    P = _np(O.unflatten(flat, variant))
    A = np.asarray(adj, dtype=np.float64)
    Yl = np.asarray(Y, dtype=np.float64)
    xe = np.asarray(x, dtype=np.float64)
    hmap = np.asarray(hmap); L = np.asarray(L)
    B, Ne, _ = A.shape
    Nc = Yl.shape[1]
    Ncr = Nc * (Nc - 1)
    I = {}
    ent = variant in (2, 4)
    edg = variant == 4           # variant 3's edge branch is dead w.r.t. the loss (model_3.py:91-97)
    x2 = xe
    if ent:
        w1 = P["ent_w1"]
        Pe = xe[..., None] * w1[0] + P["ent_b1"] + w1[2]
        Qe = xe[..., None] * w1[1]
        De = w1[3] - w1[2]
        RS1, CS1 = pairsum_fwd(Pe, Qe, De, A)
        Ebar = (RS1 + CS1) @ P["ent_w5"] + 2 * (Ne - 1) * P["ent_b5"]
        zpre = xe[..., None] * P["nod_w1"][0] + Ebar @ P["nod_w1"][1:] + P["nod_b1"]
        z = relu(zpre)
        upre = z @ P["nod_w2"][:, 0] + P["nod_b2"][0]
        x2 = relu(upre)
        I.update(RS1=RS1, CS1=CS1, x2=x2)
    a0, a1 = 1.0 - A, A
    if edg:
        Pg = xe[..., None] * P["edg_w11"][0] + P["edg_b1"] + P["edg_w12"][0]
        Qg = xe[..., None] * P["edg_w11"][0]
        Dg = P["edg_w12"][1] - P["edg_w12"][0]
        RSe, CSe = pairsum_fwd(Pg, Qg, Dg, A)
        re, ce_, PRe, PCe = head_node_fwd(RSe, CSe, P["edg_w2"], P["edg_b2"], P["eup_w1"], P["eup_b1"], Ne)
        Dge = P["eup_w1"][1] - P["eup_w1"][0]
        te, he, lge, a0, a1 = score_pass1(PRe, PCe, Dge, P["eup_w2"], P["eup_b2"], A)
        I.update(RSe=RSe, CSe=CSe, a1=a1)
    B2 = np.stack([np.broadcast_to(x2[:, :, None], A.shape), np.broadcast_to(x2[:, None, :], A.shape), a0, a1], -1)
    nb = pool_fwd(B2, hmap, L, Nc)
    v1 = P["hnk_w1"]
    PH = nb @ v1[0:4] + P["hnk_b1"] + v1[8]
    QH = nb @ v1[4:8]
    DH = v1[9] - v1[8]
    RS3, CS3 = pairsum_fwd(PH, QH, DH, Yl)
    r, c, PR, PC = head_node_fwd(RS3, CS3, P["hnk_w2"], P["hnk_b2"], P["scr_w1"], P["scr_b1"], Nc)
    Ds = P["scr_w1"][1] - P["scr_w1"][0]
    t4, h4, lg, p0, p1 = score_pass1(PR, PC, Ds, P["scr_w2"], P["scr_b2"], Yl)
    ci, cj = O.pair_index(Nc)
    logits = np.stack([lg[:, ci, cj, 0], lg[:, ci, cj, 1]], 1)
    probs = np.stack([p0[:, ci, cj], p1[:, ci, cj]], 1)
    d = lg[..., 1] - lg[..., 0]
    z_ = np.where(Yl > 0.5, -d, d)
    cegrid = (np.maximum(z_, 0) + np.log1p(np.exp(-np.abs(z_)))) * (1 - np.eye(Nc))
    ce = cegrid.sum() / (B * Ncr)
    I.update(nb=nb, RS3=RS3, CS3=CS3, PR=PR, PC=PC)

    # ------------------------------ backward ------------------------------
    scale = 10.0 / (B * Ncr)
    delta = scale * (p1 - Yl) * (1 - np.eye(Nc))
    RSm, CSm, LSm, HS = score_pass2(t4, h4, delta, Yl)
    G = {}
    g, GRh, GCh = head_node_bwd(RSm, CSm, LSm, HS, delta.sum((1, 2)), RS3, CS3, r, c,
                                P["hnk_w2"], P["scr_w1"], P["scr_w2"], Nc)
    G["scr_w2"], G["scr_b2"], G["scr_b1"], G["scr_w1"] = g["w2_head"], g["b2_head"], g["b1_head"], g["w1_head"]
    G["hnk_w2"], G["hnk_b2"] = g["w2_pair"], g["b2_pair"]
    RS3d, CS3d, LS3 = pairsum_bwd(PH, QH, DH, Yl, GRh, GCh)
    dv1 = np.zeros_like(v1)
    db = RS3d.sum((0, 1))
    dv1[9] = LS3.sum(0); dv1[8] = db - LS3.sum(0)
    dv1[0:4] = np.einsum("bna,bnk->ak", nb, RS3d)
    dv1[4:8] = np.einsum("bna,bnk->ak", nb, CS3d)
    G["hnk_w1"], G["hnk_b1"] = dv1, db
    dnb = RS3d @ v1[0:4].T + CS3d @ v1[4:8].T
    I.update(dnb=dnb, RS3d=RS3d)
    if ent or edg:
        dB2 = pool_bwd(dnb, hmap, L, Ne)
    if ent:
        dx2 = dB2[..., 0].sum(2) + dB2[..., 1].sum(1)
        du = dx2 * (upre > 0)
        G["nod_w2"] = np.einsum("bn,bnk->k", du, z)[:, None]
        G["nod_b2"] = np.array([du.sum()])
        dzp = du[..., None] * P["nod_w2"][:, 0] * (zpre > 0)
        dw = np.zeros_like(P["nod_w1"])
        dw[0] = np.einsum("bn,bnk->k", xe, dzp)
        dw[1:] = np.einsum("bnm,bnk->mk", Ebar, dzp)
        G["nod_w1"], G["nod_b1"] = dw, dzp.sum((0, 1))
        dEbar = dzp @ P["nod_w1"][1:].T
        G["ent_w5"] = np.einsum("bna,bnm->am", RS1 + CS1, dEbar)
        G["ent_b5"] = 2 * (Ne - 1) * dEbar.sum((0, 1))
        gE = dEbar @ P["ent_w5"].T
        RS1d, CS1d, LS1 = pairsum_bwd(Pe, Qe, De, A, gE, gE)
        db1 = RS1d.sum((0, 1))
        dw1 = np.zeros_like(w1)
        dw1[3] = LS1.sum(0); dw1[2] = db1 - LS1.sum(0)
        dw1[0] = np.einsum("bn,bnk->k", xe, RS1d)
        dw1[1] = np.einsum("bn,bnk->k", xe, CS1d)
        G["ent_w1"], G["ent_b1"] = dw1, db1
        I.update(dx2=dx2, gE=gE, RS1d=RS1d)
    if edg:
        da0, da1 = dB2[..., 2], dB2[..., 3]
        de = a1 * a0 * (da1 - da0) * (1 - np.eye(Ne))
        RSm, CSm, LSm, HS = score_pass2(te, he, de, A)
        g, GRe, GCe = head_node_bwd(RSm, CSm, LSm, HS, de.sum((1, 2)), RSe, CSe, re, ce_,
                                    P["edg_w2"], P["eup_w1"], P["eup_w2"], Ne)
        G["eup_w2"], G["eup_b2"], G["eup_b1"], G["eup_w1"] = g["w2_head"], g["b2_head"], g["b1_head"], g["w1_head"]
        G["edg_w2"], G["edg_b2"] = g["w2_pair"], g["b2_pair"]
        RSed, CSed, LSe = pairsum_bwd(Pg, Qg, Dg, A, GRe, GCe)
        dbe = RSed.sum((0, 1))
        G["edg_b1"] = dbe
        G["edg_w12"] = np.stack([dbe - LSe.sum(0), LSe.sum(0)], 0)
        G["edg_w11"] = (np.einsum("bn,bnk->k", xe, RSed) + np.einsum("bn,bnk->k", xe, CSed))[None, :]
    # regularisers (model_2.py:121-130, 326-336)
    parts = []
    for name, _, shape in O.param_spec(variant):
        parts.append(np.asarray(G.get(name, np.zeros(shape))).reshape(-1))
    grad = np.concatenate(parts)
    pf = flat.detach().double().numpy()
    grad = grad + 0.001 * pf
    off = 0
    for name, _, shape in O.param_spec(variant):
        n = int(np.prod(shape))
        if name.startswith("theta"):
            th = pf[off:off + n]
            grad[off:off + n] += 0.1 * 0.01 * th / np.sqrt((th ** 2).sum())
        off += n
    return dict(logits=logits, probs=probs, ce=ce, grad=grad, I=I)
