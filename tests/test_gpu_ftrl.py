"""The linear scope's own optimizer: first-order weights under Ftrl
(the reference's default linear_optimizer, models/DeepFM/deepFM.py:58, 236-241) while the embedding tables
stay on Adagrad / SGD, fused into the same backward pass.  Two consecutive steps against the fp64 oracle."""
import numpy as np
import pytest
import torch

from oracle import deepctr_oracle as O
from tests._util import REL, make_case, rel_err, to_dev

pytestmark = pytest.mark.gpu

SHAPES = [(64, [50, 1, 9, 1000, 3, 17, 1], 16), (257, [100] * 26 + [1] * 13, 16), (700, [5, 1, 3], 8),
          (3000, [40, 2000, 1], 32)]


def _oracle_steps(case, steps, table_opt, lin_opt, lr, lin_lr, l1, l2):
    t, w = case["table"].astype(np.float64), case["w1"].astype(np.float64)
    acc, n1, z1 = np.full_like(t, 0.1), np.full_like(w, 0.1), np.zeros_like(w)
    touched = np.zeros(case["N"], bool)
    for g_first, g_fm, u in steps:
        rows, G, g1, _ = O.embedding_backward(t, case["off"], case["idx"], case["val"], g_first, g_fm, u, "sum",
                                              np.float64)
        touched[rows] = True
        (O.sparse_adagrad(t, acc, rows, G, lr) if table_opt == "adagrad" else O.sparse_sgd(t, rows, G, lr))
        if lin_opt == "ftrl":
            O.sparse_ftrl(w, n1, z1, rows, g1, lin_lr, l1, l2)
        elif lin_opt == "adagrad":
            O.sparse_adagrad(w, n1, rows, g1, lin_lr)
        else:
            O.sparse_sgd(w, rows, g1, lin_lr)
    return t, w, n1, z1, touched


@pytest.mark.parametrize("B,rows,K", SHAPES)
@pytest.mark.parametrize("table_opt,lin_opt,l1,l2", [("adagrad", "ftrl", 0.0, 0.0), ("adagrad", "ftrl", 0.02, 0.1),
                                                      ("sgd", "ftrl", 0.0, 0.05), ("sgd", "adagrad", 0.0, 0.0),
                                                      ("adagrad", "sgd", 0.0, 0.0)])
@pytest.mark.parametrize("sharded", [False, True])
def test_linear_scope_optimizer(pkg, cuda, B, rows, K, table_opt, lin_opt, l1, l2, sharded):
    case = make_case(17, B, rows, K, weighted=True, prune=True, skew=2.0)
    rng, F = case["rng"], case["F"]
    lr, lin_lr = 0.05, 0.2
    cls = pkg.ShardedEmbeddingFM if sharded else pkg.EmbeddingFM
    layer = cls(F, K, [int(r) for r in rows], optimizer=table_opt, lr=lr, linear_optimizer=lin_opt, linear_lr=lin_lr,
                l1_regularization_strength=l1, l2_regularization_strength=l2).train()
    layer.load_tables(case["table"], case["w1"])
    idx, val = to_dev(case["idx"]), to_dev(case["val"])
    steps = []
    for _ in range(2):
        g_first = rng.standard_normal(B).astype(np.float32)
        g_fm = (rng.standard_normal(B) * 0.1).astype(np.float32)
        u = (rng.standard_normal((B, F, K)) * 0.1).astype(np.float32)
        steps.append((g_first, g_fm, u))
        first, fm, emb = layer(idx, val)
        torch.autograd.backward((first, fm, emb), (to_dev(g_first)[:, None], to_dev(g_fm)[:, None],
                                                   to_dev(u.reshape(B, -1))))
    torch.cuda.synchronize()
    t64, w64, n64, z64, touched = _oracle_steps(case, steps, table_opt, lin_opt, lr, lin_lr, l1, l2)
    N = case["N"]
    got_w = layer.w1.cpu().numpy()[:N]
    scale = np.abs(case["w1"]).max() + lin_lr
    assert rel_err(got_w, w64, scale) <= REL
    assert np.array_equal(got_w[~touched], case["w1"][~touched])
    assert rel_err(layer.table.cpu().numpy()[:N], t64, np.abs(case["table"]).max()) <= REL
    if lin_opt == "ftrl":
        got_z = layer.lin_z.cpu().numpy()[:N, 0]
        assert rel_err(got_z, z64, np.abs(z64).max() + 1.0) <= REL
        assert np.all(got_z[~touched] == 0)
        if l1 > 0:       # the proximal step produces exact zeros exactly where the oracle does
            assert np.array_equal(got_w[touched] == 0, w64[touched] == 0) or \
                np.abs(np.abs(z64[touched]) - l1).min() < 1e-5      # unless a |z| sits on the threshold
    if lin_opt in ("ftrl", "adagrad"):
        got_n = layer.w1_accum.cpu().numpy()[:N]
        assert rel_err(got_n, n64, 0.1 + np.abs(n64).max()) <= REL
        assert np.all(got_n[~touched] == np.float32(0.1))


def test_linear_optimizer_arguments(pkg, cuda):
    with pytest.raises(ValueError):
        pkg.EmbeddingFM(2, 8, [4, 4], linear_optimizer="adam")
    with pytest.raises(ValueError):
        pkg.EmbeddingFM(2, 8, [4, 4], linear_optimizer="ftrl", l1_regularization_strength=-1.0)
    layer = pkg.EmbeddingFM(2, 8, [4, 4], optimizer="sgd", linear_optimizer="ftrl")
    assert layer.lin_z is not None and layer.w1_accum is not None and layer.accum is None
    assert float(layer.w1_accum[0]) == pytest.approx(0.1)
    plain = pkg.EmbeddingFM(2, 8, [4, 4], optimizer="sgd")
    assert plain.lin_z is None and plain.w1_accum is None


@pytest.mark.parametrize("l1,l2", [(0.0, 0.0), (0.01, 0.01), (0.3, 0.0), (0.0, 0.5)])
def test_proximal_adagrad_tables(pkg, cuda, l1, l2):
    """dnn_optimizer = tf.train.ProximalAdagradOptimizer(lr, l1, l2) (models/ESMM/train.py:137-139) on the tables and,
    with no separate linear optimizer, on the first-order weights: two steps against the oracle's rule."""
    from oracle import deepctr_oracle as O
    from tests._util import REL, make_case, rel_err, to_dev
    B, rows, K, lr = 300, [40, 1, 700, 5, 1], 8, 0.05
    case = make_case(23, B, rows, K, weighted=True, prune=True, skew=2.0)
    rng, F = case["rng"], case["F"]
    layer = pkg.EmbeddingFM(F, K, rows, optimizer="proximal_adagrad", lr=lr, optimizer_l1=l1, optimizer_l2=l2).train()
    layer.load_tables(case["table"], case["w1"])
    t, w = case["table"].astype(np.float64), case["w1"].astype(np.float64)
    acc, acc1 = np.full_like(t, 0.1), np.full_like(w, 0.1)
    idx, val = to_dev(case["idx"]), to_dev(case["val"])
    for step in range(2):
        g_first = rng.standard_normal(B).astype(np.float32)
        g_fm = (rng.standard_normal(B) * 0.1).astype(np.float32)
        u = (rng.standard_normal((B, F, K)) * 0.1).astype(np.float32)
        first, fm, emb = layer(idx, val)
        torch.autograd.backward((first, fm, emb), (to_dev(g_first)[:, None], to_dev(g_fm)[:, None], to_dev(u.reshape(B, -1))))
        urows, G, g1, _ = O.embedding_backward(t, case["off"], case["idx"], case["val"], g_first, g_fm, u, "sum", np.float64)
        O.sparse_proximal_adagrad(t, acc, urows, G, lr, l1, l2)
        O.sparse_proximal_adagrad(w, acc1, urows, g1, lr, l1, l2)
    torch.cuda.synchronize()
    assert rel_err(layer.table.cpu().numpy(), t, np.abs(case["table"]).max()) <= 2 * REL
    assert rel_err(layer.w1.cpu().numpy(), w, np.abs(case["w1"]).max() + 1e-3) <= 2 * REL
    assert rel_err(layer.accum.cpu().numpy(), acc, 0.1 + np.abs(acc).max()) <= 2 * REL
    if l1 >= 0.3:
        assert (layer.table.cpu().numpy()[urows] == 0).any(), "a strong l1 must zero some components exactly"
    with pytest.raises(ValueError):      # (the sharded layer takes the rule too: tests/test_gpu_sharded.py)
        pkg.ShardedEmbeddingFM(F, K, rows, optimizer="proximal_adagrad", optimizer_l1=-1.0)
    with pytest.raises(ValueError):
        pkg.EmbeddingBagFM(F, K, rows, optimizer="proximal_adagrad")
