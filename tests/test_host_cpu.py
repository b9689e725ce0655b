"""Host-side logic that needs no GPU: checkpoint export / import in the reference's per-column shape,
state_dict round trips, the roofline byte model against SURVEY.md's worked numbers, the InputLayer column map."""
import numpy as np
import pytest
import torch


def test_export_import_columns_round_trip(pkg):
    rows, K = [5, 1, 7], 8
    a = pkg.EmbeddingFM(3, K, rows, optimizer="adagrad", linear_optimizer="ftrl", device="cpu")
    with torch.no_grad():
        a.w1.normal_()
        a.accum.uniform_(0.1, 1.0)
        a.w1_accum.uniform_(0.1, 1.0)
        a.lin_z.normal_()
    names = ["C1_embedding", "I1", "C2_embedding"]
    ck = a.export_columns(names, with_slots=True)
    assert ck["C1_embedding/embedding_weights"].shape == (5, K) and ck["I1/weights"].shape == (1, 1)
    assert "C2_embedding/embedding_weights/Adagrad" in ck and "C2_embedding/weights/Ftrl" in ck and "I1/weights/Ftrl_1" in ck
    assert torch.equal(ck["C2_embedding/embedding_weights"], a.table[6:13])
    b = pkg.EmbeddingFM(3, K, rows, optimizer="adagrad", linear_optimizer="ftrl", device="cpu")
    b.import_columns(ck, names)
    for x, y in ((a.table, b.table), (a.w1, b.w1), (a.accum, b.accum), (a.w1_accum, b.w1_accum), (a.lin_z, b.lin_z)):
        assert torch.equal(x, y)
    # weights only (a TF checkpoint without optimizer slots): accumulators keep their initial value
    c = pkg.EmbeddingFM(3, K, rows, optimizer="adagrad", device="cpu")
    c.import_columns(a.export_columns(names), names)
    assert torch.equal(c.table, a.table) and float(c.accum.min()) == float(c.accum.max()) == pytest.approx(0.1)
    # partial warm start and error behaviour
    d = pkg.EmbeddingFM(3, K, rows, device="cpu")
    before = d.table.clone()
    d.import_columns({k: v for k, v in ck.items() if k.startswith("I1/") and "Ftrl" not in k}, names, strict=False)
    assert torch.equal(d.table[5:6], a.table[5:6]) and torch.equal(d.table[:5], before[:5])
    with pytest.raises(KeyError):
        d.import_columns({}, names)
    with pytest.raises(ValueError):
        d.import_columns({"C1_embedding/embedding_weights": torch.zeros(4, K), "C1_embedding/weights": torch.zeros(5, 1)},
                         names, strict=False)
    with pytest.raises(ValueError):
        pkg.EmbeddingFM(3, K, 13, device="cpu").export_columns()


def test_state_dict_round_trip(pkg):
    a = pkg.DeepFM(3, 8, [5, 1, 7], dnn_hidden_units=(16, 8), linear_optimizer="ftrl", device="cpu")
    b = pkg.DeepFM(3, 8, [5, 1, 7], dnn_hidden_units=(16, 8), linear_optimizer="ftrl", device="cpu")
    b.load_state_dict(a.state_dict())
    assert torch.equal(a.embedding.rows, b.embedding.rows) and torch.equal(a.embedding.lin_z, b.embedding.lin_z)
    assert torch.equal(a.dnn.final.weight, b.dnn.final.weight)


def test_roofline_bytes_match_the_survey(pkg):
    """SURVEY.md section 8d worked numbers (K = 16, F = 39, weighted): 3 128 B / sample forward (+2 496 with E),
    4 + 39*76 + u*39*272 backward, 4 992 / 7 488 B per sample for the 6-layer cross on d = 624."""
    R = pkg.roofline
    B, F, K = 65536, 39, 16
    assert R.embed_fwd_bytes(B, F, K, True, False) == B * 3128
    assert R.embed_fwd_bytes(B, F, K, True, True) == B * (3128 + 2496)
    U = 1_566_675
    assert R.embed_bwd_bytes(B, F, K, U, True, True, "adagrad") == 4 * B + B * F * 64 + B * F * 76 + U * 272
    assert R.embed_bwd_bytes(B, F, K, U, True, False, "sgd") == 4 * B + B * F * 76 + U * 136
    assert R.cross_fwd_bytes(B, 624, 6) == B * 4992 + 2 * 6 * 624 * 4
    assert R.cross_bwd_bytes(B, 624, 6) == B * 7488 + 4 * 6 * 624 * 4
    peak, src = R.measured_peaks()
    assert peak > 1000 and src in ("measured", "fallback")


def test_input_layer_column_map(pkg):
    cols = [("z", "embedding", 4), ("a", "indicator", 3), ("m", "embedding", 2), ("b", "numeric", 1)]
    layer = pkg.InputLayer(cols, device="cpu")
    # name order: a(3) b(1) m(2) z(4)
    assert layer.output_dim == 10
    assert layer.col_kind.tolist() == [1, 1, 1, 0, 2, 2, 2, 2, 2, 2]
    assert layer.col_arg.tolist()[:3] == [0, 1, 2]
    assert layer.col_src.tolist()[4:] == [4, 5, 0, 1, 2, 3]        # m is the second embedding column: components 4, 5
    assert layer.emb_col.tolist() == [6, 7, 8, 9, 4, 5]
    with pytest.raises(ValueError, match="no CPU path"):
        layer(torch.zeros((2, 1)), torch.zeros((2, 1), dtype=torch.int64), torch.zeros((2, 6)))


def test_torch_reference_of_the_full_size_tests_matches_the_oracle():
    """tests/test_gpu_zz_fullsize.py checks the kernels at BASELINE sizes against `torch_embedding_reference`
    (the numpy oracle is too slow there); here that reference is itself pinned to the oracle on a small case."""
    from oracle import deepctr_oracle as O
    from tests._util import make_case, torch_embedding_reference
    case = make_case(21, 150, [9, 1, 40, 3, 1], 8, weighted=True, prune=True, skew=2.0)
    rng, B, F, K = case["rng"], case["B"], case["F"], case["K"]
    g_first, g_fm = rng.standard_normal(B).astype(np.float32), rng.standard_normal(B).astype(np.float32)
    u = rng.standard_normal((B, F, K)).astype(np.float32)
    ref = torch_embedding_reference(torch.as_tensor(case["table"]), torch.as_tensor(case["w1"]), 0.25,
                                    torch.as_tensor(case["off"]), torch.as_tensor(case["idx"]), torch.as_tensor(case["val"]),
                                    torch.as_tensor(g_first), torch.as_tensor(g_fm), torch.as_tensor(u))
    t64, w64 = case["table"].astype(np.float64), case["w1"].astype(np.float64)
    e, keep = O.embedding_lookup(t64, case["off"], case["idx"], case["val"], "sum", np.float64)
    assert np.allclose(ref["e"].numpy(), e, rtol=0, atol=0)
    assert np.allclose(ref["fm"].numpy(), O.fm_second_order(e)[:, 0], rtol=1e-12, atol=1e-12)
    assert np.allclose(ref["first"].numpy(), O.first_order(w64, 0.25, case["off"], case["idx"], case["val"], np.float64)[:, 0],
                       rtol=1e-12, atol=1e-12)
    rows, G, g1, _, Gabs, g1abs = O.embedding_backward(t64, case["off"], case["idx"], case["val"], g_first, g_fm, u,
                                                        "sum", np.float64, return_abs=True)
    assert np.array_equal(np.flatnonzero(ref["touched"].numpy()), rows)
    assert np.allclose(ref["G"].numpy()[rows], G, rtol=1e-11, atol=1e-12)
    assert np.allclose(ref["g1"].numpy()[rows], g1, rtol=1e-11, atol=1e-12)
    assert np.allclose(ref["g1abs"].numpy()[rows], g1abs, rtol=1e-11, atol=1e-12)
    assert np.all(ref["Gfloor"].numpy()[rows] >= Gabs * (1 - 1e-12))        # the cancellation-aware floor dominates
    untouched = np.setdiff1d(np.arange(case["N"]), rows)
    assert np.all(ref["G"].numpy()[untouched] == 0)


def test_shard_plan_at_terabyte_scale(pkg):
    """cfg4 (BASELINE.json configs[3]): ~880 M rows over 2 / 4 / 8 ranks.  The composite sort key
    owner * cap + local row must stay below 2^32 - 1 (it travels as uint32), and owner / local / global
    must round-trip at the far end of the row range."""
    w = pkg.synth.cfg("cfg4")
    assert w.n_rows == 880_000_013 and w.field_size == 39
    rng = np.random.default_rng(1)
    for G in (2, 4, 8):
        for rank in (0, G - 1):
            plan = pkg.ShardPlan(list(w.rows_per_field), G, rank)
            assert plan.cap * G < 2 ** 32 - 1 and plan.cap * G >= w.n_rows
            assert sum(pkg.ShardPlan(list(w.rows_per_field), G, r).n_local for r in range(G)) == w.n_rows
            rows = torch.as_tensor(np.concatenate([rng.integers(0, w.n_rows, size=1000),
                                                   [0, w.n_rows - 1, w.n_rows - G, 291_999_999, 292_000_000]]))
            owner, local = plan.owner(rows), plan.local_row(rows)
            assert int(owner.max()) < G and int(local.max()) < plan.cap
            assert torch.equal(plan.global_row(local, owner), rows)
            key = owner * plan.cap + local
            assert int(key.max()) < plan.cap * G            # < the pruned key G * cap, itself < 2^32 - 1
    # one rank more than the uint32 key space allows is refused
    with pytest.raises(ValueError):
        pkg.ShardPlan([2 ** 32], 2, 0)
    # memory per rank: rows + accumulators interleaved (128 B per row at K = 16) must fit 180 GB from G = 2 on
    for G in (2, 4, 8):
        assert pkg.ShardPlan(list(w.rows_per_field), G, 0).cap * 128 < 100e9


def test_documents_name_files_that_exist():
    """DESIGN.md / INTEGRATION.md / README.md cite files of this repository (tests, sources, profiles): none may dangle."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg_dir = os.path.join(root, "details-in-recommendation_b200")
    missing = []
    for doc in ("DESIGN.md", "INTEGRATION.md", "README.md"):
        text = open(os.path.join(root, doc)).read()
        for m in set(re.findall(r"`([A-Za-z0-9_./\-]+\.(?:py|cu|cuh|h|md|json|txt|csv|npz))`", text)):
            if m.startswith(("models/", "dataset/", "examples/")) or "/root/" in m or "*" in m:
                continue                                            # reference citations
            if m in ("MEASURED_PEAKS.json", "COPYCHECK.json") or m.startswith(("BENCH_", "SCALE_")):
                continue                                            # driver-written, not part of the tree
            cands = [os.path.join(root, m), os.path.join(root, "tests", m), os.path.join(pkg_dir, m),
                     os.path.join(pkg_dir, "csrc", m), os.path.join(root, "oracle", m), os.path.join(root, "tests", "golden", m),
                     os.path.join(root, "include", m), os.path.join(root, "tools", m), os.path.join(root, "profiles", m)]
            if not any(os.path.exists(c) for c in cands):
                missing.append((doc, m))
    assert not missing, missing
