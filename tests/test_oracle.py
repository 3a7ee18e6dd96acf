"""CPU checks of the oracle itself: dense one-hot transcription == closed index form, and the
hand-derived kernel plan (forward + backward) == autograd of the oracle."""
import numpy as np
import pytest
import torch

from hdgnn_b200.synthetic import make_commits
from oracle import hdgnn_oracle as O
from oracle import plan_numpy as PN


def _params(variant, seed=7, bias=True):
    flat = O.init_params(variant, seed=seed, dtype=torch.float64)
    if bias:   # the reference starts biases at 0; perturb them so their gradients/paths are exercised
        g = torch.Generator().manual_seed(seed + 1)
        flat = flat + 0.05 * torch.randn(flat.numel(), generator=g, dtype=torch.float64)
    return flat


@pytest.mark.parametrize("variant", [1, 2, 3, 4])
def test_dense_equals_closed(variant):
    cb = make_commits(3, 7, 4, seed=3, p_edge=0.3, p_short=0.5, p_noise=0.3)
    flat = _params(variant)
    P = O.unflatten(flat, variant)
    d = O.forward_dense(variant, P, O.dense_inputs(cb.adj, cb.x, cb.hmap, cb.L, cb.Y))
    c = O.forward_closed(variant, P, cb.adj, cb.x, cb.hmap, cb.L, cb.Y)
    for k in ("logits", "probs", "ce"):
        assert torch.allclose(d[k], c[k], rtol=1e-11, atol=1e-12), k
    if variant in (2, 4):
        assert torch.allclose(d["E_node2"], c["E_node2"], rtol=1e-11, atol=1e-12)
    if variant in (3, 4):
        assert torch.allclose(d["E_edge2"], c["E_edge2"], rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("variant", [1, 2, 3, 4])
def test_dense_grad_equals_closed_grad(variant):
    cb = make_commits(2, 6, 4, seed=5, p_edge=0.3, p_short=0.5, p_noise=0.3)
    flat = _params(variant)
    ld, _, _, gd, _ = O.train_loss_and_grad(variant, flat, cb.adj, cb.x, cb.hmap, cb.L, cb.Y, dense=True)
    lc, _, _, gc, _ = O.train_loss_and_grad(variant, flat, cb.adj, cb.x, cb.hmap, cb.L, cb.Y, dense=False)
    assert torch.allclose(ld, lc, rtol=1e-12)
    assert torch.allclose(gd, gc, rtol=1e-9, atol=1e-13)


@pytest.mark.parametrize("variant", [1, 2, 3, 4])
@pytest.mark.parametrize("shape", [(3, 9, 5), (2, 33, 12)])
def test_plan_equals_autograd(variant, shape):
    B, Ne, Nc = shape
    cb = make_commits(B, Ne, Nc, seed=11, p_edge=0.2, p_short=0.5, p_noise=0.2)
    flat = _params(variant)
    loss, ce, _, grad, out = O.train_loss_and_grad(variant, flat, cb.adj, cb.x, cb.hmap, cb.L, cb.Y)
    plan = PN.train_step_plan(variant, flat, cb.adj, cb.x, cb.hmap, cb.L, cb.Y)
    assert np.allclose(plan["logits"], out["logits"].numpy(), rtol=1e-10, atol=1e-12)
    assert np.allclose(plan["probs"], out["probs"].numpy(), rtol=1e-10, atol=1e-12)
    assert np.isclose(plan["ce"], float(ce), rtol=1e-12)
    g = grad.numpy()
    assert np.allclose(plan["grad"], g, rtol=1e-8, atol=1e-12 + 1e-9 * np.abs(g).max())


def test_tf_adam_closed_form():
    rng = np.random.default_rng(0)
    p = rng.normal(size=50); g = rng.normal(size=50)
    m = np.zeros(50); v = np.zeros(50)
    p1, m1, v1 = O.tf_adam_step(p, g, m, v, 1)
    # first step of TF Adam: m = .1 g, v = .001 g^2, lr_t = lr*sqrt(.001)/.1
    lr_t = 3e-4 * np.sqrt(1 - 0.999) / (1 - 0.9)
    assert np.allclose(p1, p - lr_t * 0.1 * g / (np.sqrt(0.001 * g * g) + 1e-8))


def test_map_conv_closed_equals_dense():
    B, No = 3, 6
    rng = np.random.default_rng(1)
    adj = (rng.random((B, No, No)) < 0.4).astype(np.float64)
    adj[:, np.arange(No), np.arange(No)] = 0
    x = torch.as_tensor(rng.normal(size=(B, No)))
    ei, ej = O.pair_index(No)
    lab = adj[:, ei, ej]
    Ra = torch.as_tensor(np.stack([1 - lab, lab], 1))
    theta = torch.as_tensor(rng.normal(size=(1, 2, 1, 1)))
    d = O.map_conv_dense(theta, Ra, x[:, None, :])
    c = O.map_conv_closed(theta, torch.as_tensor(adj), x)
    assert torch.allclose(d, c, rtol=1e-11)
