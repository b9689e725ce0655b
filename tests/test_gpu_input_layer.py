"""tf.feature_column.input_layer in the reference's own DCN convention: numeric pass-through,
one-hot indicator columns and embedding columns side by side in column-name order -- the census layout of
models/DeepCrossNetwork/train.py:88-100 (d = 51) and a Criteo-shaped one (26 x 16 + 13 = 429)."""
import numpy as np
import pytest
import torch

from oracle import deepctr_oracle as O
from tests._util import to_dev

pytestmark = pytest.mark.gpu

CENSUS = [("age", "numeric", 1), ("education_num", "numeric", 1), ("capital_gain", "numeric", 1),
          ("capital_loss", "numeric", 1), ("hours_per_week", "numeric", 1), ("workclass_indicator", "indicator", 9),
          ("education_indicator", "indicator", 16), ("marital_status_indicator", "indicator", 7),
          ("relationship_indicator", "indicator", 6), ("occupation_embedding", "embedding", 8)]
CRITEO = [("I%d" % i, "numeric", 1) for i in range(13)] + [("C%d_embedding" % i, "embedding", 16) for i in range(26)]
MIXED = [("z", "embedding", 4), ("a", "indicator", 3), ("m", "embedding", 8), ("b", "numeric", 1)]


def _inputs(rng, columns, B):
    n_num = sum(k == "numeric" for _, k, _ in columns)
    ind = [s for _, k, s in columns if k == "indicator"]
    emb_w = sum(s for _, k, s in columns if k == "embedding")
    numeric = rng.standard_normal((B, n_num)).astype(np.float32) if n_num else None
    ids = np.stack([rng.integers(-1, s + 1, size=B) for s in ind], 1).astype(np.int64) if ind else None   # -1 / s: OOV
    emb = rng.standard_normal((B, emb_w)).astype(np.float32) if emb_w else None
    return numeric, ids, emb


@pytest.mark.parametrize("columns,B", [(CENSUS, 257), (CRITEO, 100), (MIXED, 1), (CENSUS[:5], 33), (CENSUS[5:9], 64)])
def test_input_layer_forward_backward(pkg, cuda, columns, B):
    rng = np.random.default_rng(3)
    numeric, ids, emb = _inputs(rng, columns, B)
    layer = pkg.InputLayer(columns)
    want, where = O.input_layer(columns, numeric, ids, emb)
    t_emb = to_dev(emb).requires_grad_(True) if emb is not None else None
    x0 = layer(to_dev(numeric), to_dev(ids), t_emb)
    assert x0.shape == want.shape == (B, layer.output_dim)
    assert np.array_equal(x0.detach().cpu().numpy(), want), "input_layer is pure data movement: bit-exact"
    if emb is not None:
        dy = rng.standard_normal(want.shape).astype(np.float32)
        x0.backward(to_dev(dy))
        torch.cuda.synchronize()
        c, got = 0, t_emb.grad.cpu().numpy()
        for name, kind, size in columns:            # listing order of the embedding columns
            if kind == "embedding":
                a, b = where[name]
                assert np.array_equal(got[:, c:c + size], dy[:, a:b])
                c += size


def test_census_layout_is_the_reference_one(pkg, cuda):
    layer = pkg.InputLayer(CENSUS)
    assert layer.output_dim == 51                         # 5 + 9 + 16 + 7 + 6 + 8, SURVEY row A7
    names = sorted(n for n, _, _ in CENSUS)
    assert names[0] == "age" and names[-1] == "workclass_indicator"
    with pytest.raises(ValueError):
        pkg.InputLayer([])
    with pytest.raises(ValueError):
        pkg.InputLayer([("a", "numeric", 1), ("a", "indicator", 3)])
    with pytest.raises(ValueError):
        pkg.InputLayer([("a", "bucketized", 3)])
    with pytest.raises(ValueError):
        layer(torch.zeros((4, 4), device="cuda"), torch.zeros((4, 4), dtype=torch.int64, device="cuda"),
              torch.zeros((4, 8), device="cuda"))


def test_dcn_on_the_reference_input_layer(pkg, cuda):
    """EmbeddingFM -> InputLayer -> CrossNetwork end to end: the table receives the gradient that flows
    back through the cross stack and the column shuffle."""
    rng = np.random.default_rng(4)
    B, K, rows = 64, 8, [100]
    table = (rng.standard_normal((100, K)) * 0.3).astype(np.float32)
    emb_layer = pkg.EmbeddingFM(1, K, rows, optimizer="sgd", lr=1.0, first_order=False).train()
    emb_layer.load_tables(table, None)
    inp = pkg.InputLayer(CENSUS)
    cross = pkg.CrossNetwork(51, 2).train()
    numeric, ids, _ = _inputs(rng, CENSUS, B)
    occ = rng.integers(0, 100, size=(B, 1)).astype(np.int64)
    _, _, e = emb_layer(to_dev(occ))
    xL = cross(inp(to_dev(numeric), to_dev(ids), e))
    dy = rng.standard_normal((B, 51)).astype(np.float32)
    xL.backward(to_dev(dy))
    torch.cuda.synchronize()
    x0, where = O.input_layer(CENSUS, numeric, ids, table[occ[:, 0]])
    w, b = cross.cross_w.detach().cpu().numpy().astype(np.float64), cross.cross_b.detach().cpu().numpy().astype(np.float64)
    dx0, _, _ = O.cross_backward(x0.astype(np.float64), w, b, dy.astype(np.float64))
    a, bnd = where["occupation_embedding"]
    want = table.astype(np.float64).copy()
    np.subtract.at(want, occ[:, 0], dx0[:, a:bnd])       # SGD, lr = 1: the delta is the summed gradient
    got = emb_layer.table.cpu().numpy()
    assert np.abs(got - want).max() <= 1e-5 * max(1.0, np.abs(dx0).max() * 8)
