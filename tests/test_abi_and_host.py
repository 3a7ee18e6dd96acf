"""CPU checks of the boundary and the host logic: the C-ABI library loads and exports every symbol
include/hdgnn.h declares, argument validation that needs no GPU, the on-disk format round trip,
and the commit-sharded data-parallel arithmetic over gloo (world_size 2)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from hdgnn_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "hdgnn.h")).read()
    declared = set(re.findall(r"\b(hdgnn_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(_lib.lib, name), f"{name} declared in include/hdgnn.h but not exported by libhdgnn.so"
    assert set(_lib.EXPORTS) == declared


def test_param_layout_matches_reference_variable_order():
    from hdgnn_b200 import _lib
    from oracle import hdgnn_oracle as O
    for variant in (1, 2, 3, 4):
        assert _lib.lib.hdgnn_param_count(variant) == O.param_count(variant)
        off = 0
        for name, _, shape in O.param_spec(variant):
            assert _lib.lib.hdgnn_param_offset(variant, name.encode()) == off, (variant, name)
            off += int(np.prod(shape))
    assert _lib.lib.hdgnn_param_count(0) < 0 and _lib.lib.hdgnn_param_count(5) < 0
    assert _lib.lib.hdgnn_param_offset(1, b"ent_w1") == -1          # model_1 has no entity block
    assert _lib.lib.hdgnn_label_pitch(200) == 208 and _lib.lib.hdgnn_label_pitch(74) == 80


def test_create_validates_arguments_without_a_gpu():
    from hdgnn_b200 import _lib
    h = C.c_void_p()
    for cfg in (_lib.Config(200, 74, 9, 10, 0, 0, 0, 0), _lib.Config(1, 74, 2, 10, 0, 0, 0, 0),
                _lib.Config(200, 74, 2, 0, 0, 0, 0, 0), _lib.Config(600, 74, 2, 10, 0, 0, 0, 0)):
        rc = _lib.lib.hdgnn_create(C.byref(cfg), C.byref(h))
        assert rc < 0 and not h.value
        assert _lib.lib.hdgnn_last_error(None)
    if not torch.cuda.is_available():
        cfg = _lib.Config(200, 74, 2, 10, 0, 0, 0, 0)
        assert _lib.lib.hdgnn_create(C.byref(cfg), C.byref(h)) == _lib.E_CUDA      # fails loudly, no fallback


def test_engine_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from hdgnn_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(20, 10)


def test_dataset_round_trip_through_reference_formats(tmp_path):
    from hdgnn_b200.synthetic import make_commits
    from hdgnn_b200.utils2 import write_dataset, read_compact, read_data, split_half
    cb = make_commits(10, 12, 5, seed=4, p_short=0.5)
    cb.hmap[np.arange(12)[None, :] >= cb.L[:, None]] = -1          # lines beyond L do not exist on disk
    write_dataset(cb, "toy", 3, root=str(tmp_path))
    back = read_compact("toy", 3, 12, 5, root=str(tmp_path))
    for a, b in zip((cb.adj, cb.x, cb.hmap, cb.L, cb.Y), (back.adj, back.x, back.hmap, back.L, back.Y)):
        assert np.array_equal(a, b)
    again = read_compact("toy", 3, 12, 5, root=str(tmp_path))           # served from the compact cache
    assert np.array_equal(again.adj, cb.adj)
    tr, te = split_half(back)
    assert tr.B == 5 and te.B == 5
    stub = type("S", (), dict(Repo="toy", Ne=12, Nc=5))()
    tup = read_data(stub, 3, root=str(tmp_path))
    assert len(tup) == 12 and tup[6].shape == (10, 12, 132) and tup[10].shape == (10, 5, 132)
    with pytest.raises(ValueError):
        read_compact("toy", 3, 13, 5, root=str(tmp_path), cache=False)


def test_loader_rejects_non_binary_adjacency(tmp_path):
    from hdgnn_b200.synthetic import make_commits
    from hdgnn_b200.utils2 import compact_from_raw
    cb = make_commits(2, 6, 3, seed=1)
    raw = cb.adj.astype(np.float64); raw[0, 1, 2] = 3.0
    with pytest.raises(IndexError):
        compact_from_raw(raw, cb.Y, [["null"] * 6] * 2, [{}] * 2, 6, 3)


WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
from hdgnn_b200.synthetic import make_commits
from oracle import hdgnn_oracle as O
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% sys.argv[1], rank=int(sys.argv[2]), world_size=2)
rank, world = dist.get_rank(), 2
variant, B, Ne, Nc = 2, 6, 9, 4
cb = make_commits(B, Ne, Nc, seed=5)                 # generated globally, then sharded
flat = O.init_params(variant, seed=3, dtype=torch.float64)
per = B // world
sh = cb.slice(rank * per, (rank + 1) * per)
# what hdgnn_forward_backward computes on a rank: d(10*CE)/dp with CE = sum over LOCAL pairs / (B_global*Ncr)
p = flat.clone().requires_grad_(True)
out = O.forward_closed(variant, O.unflatten(p, variant), sh.adj, sh.x, sh.hmap, sh.L, sh.Y)
ce_local = out["ce"] * per / B
(g,) = torch.autograd.grad(10.0 * ce_local, p)
dist.all_reduce(g)                                    # the single collective of the step
ce = ce_local.detach().clone(); dist.all_reduce(ce)
# regularisers are replica-identical and added after the reduce (hdgnn_adam_step)
lm, lp = O.reg_loss(flat, variant)
q = flat.clone().requires_grad_(True)
lm2, lp2 = O.reg_loss(q, variant)
(gr,) = torch.autograd.grad(0.1 * lm2 + lp2, q)
full = O.train_loss_and_grad(variant, flat, cb.adj, cb.x, cb.hmap, cb.L, cb.Y)
assert torch.allclose(g + gr, full[3], rtol=1e-10, atol=1e-14), (g + gr - full[3]).abs().max()
assert torch.allclose(ce, full[1], rtol=1e-12)
pn, _, _ = O.tf_adam_step(flat, g + gr, torch.zeros_like(flat), torch.zeros_like(flat), 1)
gather = [torch.zeros_like(pn) for _ in range(world)]
dist.all_gather(gather, pn)
assert torch.equal(gather[0], gather[1])              # every rank applies the identical update
dist.destroy_process_group()
print("ok", rank)
"""


def test_commit_sharded_data_parallel_arithmetic_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % ROOT)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_host_batch_wire_block_layout():
    """HostBatch(bits=True) keeps the five inputs of a step in ONE block, every array starting at the next multiple of 16 bytes in
    the order adj, Y, x, hmap, L -- the layout the *_host entry points copy with a single DMA (include/hdgnn.h, HDGNN_F_LABEL_BITS)
    -- with the same contents as the separately allocated form."""
    from hdgnn_b200 import _lib
    from hdgnn_b200.model import HostBatch
    from hdgnn_b200.synthetic import make_commits
    from hdgnn_b200.engine import pack_label_bits
    for B, Ne, Nc in ((5, 200, 74), (3, 33, 12), (1, 250, 150)):
        cb = make_commits(B, Ne, Nc, seed=B)
        hb = HostBatch(cb, bits=True)
        WPe, WPc = _lib.lib.hdgnn_bit_words(Ne), _lib.lib.hdgnn_bit_words(Nc)
        sizes = [B * Ne * WPe * 4, B * Nc * WPc * 4, B * Ne * 4, B * Ne * 4, B * 4]
        offs, o = [], 0
        for n in sizes:
            offs.append(o)
            o += (n + 15) & ~15
        base = hb.adj.data_ptr()
        assert [t.data_ptr() - base for t in (hb.adj, hb.Yw, hb.x, hb.hmap, hb.L)] == offs
        assert hb.adj.shape == (B, Ne, WPe) and hb.Yw.shape == (B, Nc, WPc) and hb.nbytes() == sum(sizes)
        assert np.array_equal(hb.adj.numpy().view(np.uint32), pack_label_bits(cb.adj)) and np.array_equal(hb.Yw.numpy().view(np.uint32), pack_label_bits(cb.Y))
        assert np.array_equal(hb.x.numpy(), cb.x) and np.array_equal(hb.hmap.numpy(), cb.hmap) and np.array_equal(hb.L.numpy(), cb.L)
        c = hb.clone()
        assert c.adj.data_ptr() != base and all(torch.equal(a, b) for a, b in zip(hb.tensors(), c.tensors()))
        plain = HostBatch(cb, bits=False)
        assert plain.adj.shape == (B, Ne, Ne) and plain.Yw is plain.Y
