"""CPU-side checks of the boundary: the C-ABI library loads here (no GPU) and exports every symbol
include/dir_b200.h declares; argument validation runs before any launch."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "dir_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dir_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(pkg):
    from dir_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    lib = _lib.lib()
    syms = _header_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), "libdir_b200.so does not export %s" % s
        assert s in _lib.SIGNATURES, "no ctypes signature for %s" % s
    assert sorted(_lib.SIGNATURES) == syms
    assert lib.dir_version() >= 100


def test_argument_validation_without_gpu(pkg):
    from dir_b200 import _lib
    lib = _lib.lib()
    # K = 12 is unsupported; no launch happens before validation
    rc = lib.dir_embed_fm_fwd(16, 12, None, 1, None, 16, None, 16, None, 10, 4, 3, 12, None, None, None, 16,
                              None, None, None)
    assert rc == -22 and b"K must be" in lib.dir_last_error()
    rc = lib.dir_cross_fwd(16, 16, 16, 4, 2000, 2, 16, None, None)
    assert rc == -22
    with pytest.raises(ValueError):
        _lib.check(rc, "dir_cross_fwd")
    assert lib.dir_embed_bwd_workspace_bytes(0, 16) == 0
    assert lib.dir_embed_bwd_workspace_bytes(39 * 1024, 16) > 39 * 1024 * 12


def test_layers_refuse_cpu_tensors(pkg):
    import torch
    layer_cls = pkg.EmbeddingFM
    with pytest.raises(ValueError):
        layer_cls(0, 16, [])
    with pytest.raises(ValueError):
        layer_cls(2, 16, [4, 4], optimizer="ftrl", device="cpu")
    layer = layer_cls(2, 16, [4, 4], device="cpu")       # construction is host-only plumbing
    with pytest.raises(ValueError, match="no CPU path"):
        layer(torch.zeros((3, 2), dtype=torch.int64))


def test_synth_workloads(pkg):
    import numpy as np
    s = pkg.synth
    w2 = s.cfg("cfg2")
    assert w2.field_size == 39 and w2.embedding_size == 16 and abs(w2.n_rows - 10_000_000) < 100
    w4 = s.cfg("cfg4")
    assert w4.field_size == 39 and w4.n_rows == 880_000_000 + 13
    w5 = s.cfg("cfg5", batch=4096)
    idx, val, lab = s.make_inputs(w5)
    assert idx.shape == (4096, 39) and idx.dtype == np.int64 and (idx[:, 26:] == 0).all()
    share = (idx[:, 0] == 0).mean()
    assert 0.08 < share < 0.18          # Zipf(1.1): hottest row ~ 12.8 % of lookups
    assert ((val[:, :26] == 1).all() and (val[:, 26:] < 1).all())
    idx2, _, _ = s.make_inputs(w5)
    assert np.array_equal(idx, idx2)    # seeded


def test_argument_validation_of_the_widened_entry_points(pkg):
    """Every new entry point validates before it launches (so this runs without a GPU): -EINVAL and a message."""
    import ctypes
    from dir_b200 import _lib
    lib = _lib.lib()
    P = 16                                    # any non-NULL, 16-byte aligned "pointer": nothing dereferences it
    # bags: combiner out of range, then K unsupported
    rc = lib.dir_embed_bag_fm_fwd(P, 16, None, 1, None, P, P, None, 4, P, None, 10, 2, 2, 16, 7, P, None, None, P,
                                  None, None, None, None, None)
    assert rc == -22 and b"combiner" in lib.dir_last_error()
    rc = lib.dir_embed_bag_fm_fwd(P, 16, None, 1, None, P, P, None, 4, P, None, 10, 2, 2, 12, 1, P, None, None, P,
                                  None, None, None, None, None)
    assert rc == -22 and b"K must be" in lib.dir_last_error()
    assert lib.dir_embed_bag_fm_fwd(None, 16, None, 1, None, None, None, None, 0, None, None, 10, 0, 2, 16, 1, None,
                                    None, None, None, None, None, None, None, None) == 0           # empty batch
    # the linear scope's optimizer: Ftrl without its z slot, an unknown optimizer code
    bad = _lib.LinearOpt(_lib.OPT_FTRL, 0.2, 0.0, 0.0, None)
    rc = lib.dir_embed_bwd_reduce_update(P, P, 32, P, P, 1, P, None, P, P, P, P, None, 4, 2, 16, 10, None, 0, None, 0,
                                         1, 0.05, None, ctypes.byref(bad), P, 1 << 20, None, None)
    assert rc == -22 and b"Ftrl needs z" in lib.dir_last_error()
    bad = _lib.LinearOpt(9, 0.2, 0.0, 0.0, None)
    rc = lib.dir_embed_bwd_reduce_update(P, P, 32, P, P, 1, P, None, P, P, P, P, None, 4, 2, 16, 10, None, 0, None, 0,
                                         1, 0.05, None, ctypes.byref(bad), P, 1 << 20, None, None)
    assert rc == -22 and b"unknown linear optimizer" in lib.dir_last_error()
    # input layer, column feed
    assert lib.dir_input_layer_fwd(None, 0, None, 0, P, 8, P, P, P, 4, 0, P, None) == -22
    assert lib.dir_input_layer_bwd(None, None, 0, 8, 8, None, None) == 0
    # the device-driven sharded exchange: the layout helper agrees with itself, every entry point refuses a layout
    # that was never filled (no launch happens before validation)
    lay = _lib.PeerLayout()
    assert lib.dir_peer_layout_init(8, 3, 16, 13, 1000, 5000, ctypes.byref(lay)) == 0
    offs = [lay.off_hdr, lay.off_ids, lay.off_rows, lay.off_w, lay.off_g, lay.off_g1, lay.off_dense, lay.total_bytes]
    assert offs == sorted(offs) and all(o % 256 == 0 for o in offs)
    assert lay.off_ids - lay.off_hdr >= 8 * 4 * 8 and lay.off_rows - lay.off_ids >= 8 * 1000 * 4
    assert lay.off_w - lay.off_rows >= (5000 + 13) * 16 * 4 and lay.off_g1 - lay.off_g >= 8 * 1000 * 16 * 4
    assert lay.total_bytes - lay.off_dense >= 8 * 13 * 20 * 4
    assert lib.dir_peer_layout_init(8, 8, 16, 0, 10, 10, ctypes.byref(lay)) == -22          # rank out of range
    assert lib.dir_peer_layout_init(2, 0, 12, 0, 10, 10, ctypes.byref(lay)) == -22 and b"K must be" in lib.dir_last_error()
    assert lib.dir_peer_layout_init(2, 0, 16, 2, 10, 10, ctypes.byref(lay)) == 0            # peer_base / local still NULL
    ref = ctypes.byref(lay)
    assert lib.dir_shard_ids_push(ref, P, P, 4, P, P, None) == -22 and b"peer_base" in lib.dir_last_error()
    assert lib.dir_shard_slots(ref, P, 10, P, P, None, None) == -22
    assert lib.dir_shard_gather_send(ref, P, 32, None, 1, P, 32, None, 0, None) == -22
    assert lib.dir_shard_g1_push(ref, P, P, 4, None) == -22
    assert lib.dir_shard_owner_update(ref, P, P, P, 32, None, None, 1, 10, P, 1, 0.05, None, None, None, None, None,
                                      None, None) == -22
    assert lib.dir_shard_dense_apply(ref, P, P, 32, None, None, 1, 0.05, None, None, P, P, 32, None, None, None, 1, P,
                                     None, None, None) == -22
    assert lib.dir_shard_dense_emit(ref, P, 32, P, None, P, None, P, P, None, P, 4, 3, P, 1 << 20, None) == -22
    with pytest.raises(ValueError):
        _lib.check(lib.dir_embed_bwd_reduce_emit_to(ref, None, None, P, P, None, P, P, 4, 2, 100, None, 0, P, P,
                                                    1 << 20, None), "emit_to")
    assert lib.dir_shard_dense_inv(None, None, None, 0, 4, 3, 0, None, None, None) == 0      # nothing to do
    assert lib.dir_shard_unique(P, P, 8, 10, 2, P, 3, 4, P, P, P, P, P, 1 << 20, None) == -22      # 8 is not B * 3
    assert lib.dir_table_init_counter(None, 16, 4, 16, 2, 0, 8, 1, 0.25, None) == -22
    assert lib.dir_shard_dense_workspace_bytes(16) > 296 * 64 * 20 * 8


def test_integration_md_stub_matches_the_library(pkg):
    """The ctypes stub printed in INTEGRATION.md (what a maintainer of the reference would paste) must bind:
    execute it against the built library and compare every argtypes list it declares with _lib.SIGNATURES."""
    import ctypes
    from dir_b200 import _lib
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = text.split("```python\n# dir_b200_binding.py", 1)[1].split("```", 1)[0]
    block = "# dir_b200_binding.py" + block
    block = block.replace('C.CDLL("details-in-recommendation_b200/libdir_b200.so")', "C.CDLL(%r)" % _lib.LIB_PATH)
    ns = {}
    exec(compile(block, "INTEGRATION.md", "exec"), ns)
    lib = ns["lib"]
    declared = 0
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(lib, name)
        if fn.argtypes is None:
            continue
        declared += 1
        assert len(fn.argtypes) == len(args), "%s: INTEGRATION.md lists %d arguments, the library takes %d" % (
            name, len(fn.argtypes), len(args))
        for a, b in zip(fn.argtypes, args):
            assert ctypes.sizeof(a) == ctypes.sizeof(b), "%s: argument width differs" % name
    assert declared >= 8
    assert ctypes.sizeof(ns["LinearOpt"]) == ctypes.sizeof(_lib.LinearOpt)
