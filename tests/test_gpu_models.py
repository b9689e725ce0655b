"""Full DeepFM / DCN graphs (details-in-recommendation_b200/models.py) on the GPU against the fp64 oracle
(oracle/full_models.py): logits of a forward pass and every parameter after one training step."""
import numpy as np
import pytest
import torch

from oracle import full_models as FM
from oracle import tf_semantics as tfs
from tests._util import REL, make_case, rel_err, to_dev

pytestmark = pytest.mark.gpu
ROWS = [50, 1, 9, 1000, 3, 17, 1]


def _copy(p):
    return {k: ([a.copy() for a in v] if isinstance(v, list) else np.copy(v)) for k, v in p.items()}


def _load_tower(tower, W, b, final):
    lins = list(tower.hidden) + ([tower.final] if final else [])
    assert len(lins) == len(W)
    with torch.no_grad():
        for lin, w, bb in zip(lins, W, b):
            lin.weight.copy_(to_dev(w.T.astype(np.float32)))      # TF kernels are [in, out]
            lin.bias.copy_(to_dev(bb.astype(np.float32)))


def _tower_err(tower, W, b, W0, final):
    lins = list(tower.hidden) + ([tower.final] if final else [])
    worst = 0.0
    for lin, w, bb, w0 in zip(lins, W, b, W0):
        worst = max(worst, rel_err(lin.weight.detach().cpu().numpy().T, w, np.abs(w0).max()),
                    rel_err(lin.bias.detach().cpu().numpy(), bb, np.abs(w0).max()))
    return worst


@pytest.mark.parametrize("sharded", [False, True])
def test_deepfm_logits_and_one_step(pkg, cuda, sharded):
    B, K, hidden, lr = 256, 16, (64, 32), 0.05
    case = make_case(11, B, ROWS, K, weighted=True, prune=True)
    rng, F = case["rng"], case["F"]
    labels = (rng.random(B) < 0.3).astype(np.float32)
    W, b = FM.make_tower(rng, F * K, hidden, 1, np.float32)
    p32 = dict(table=case["table"], w1=case["w1"], bias=np.float32(0.03), W=W, b=b)
    p = {k: ([a.astype(np.float64) for a in v] if isinstance(v, list) else np.asarray(v, np.float64)) for k, v in p32.items()}
    p["bias"] = np.float64(p32["bias"])
    model = pkg.DeepFM(F, K, ROWS, dnn_hidden_units=hidden, dnn_learning_rate=lr, sharded=sharded).train()
    model.embedding.load_tables(case["table"], case["w1"])
    with torch.no_grad():
        model.embedding.bias.fill_(float(p32["bias"]))
    _load_tower(model.dnn, W, b, True)
    idx, val, y = to_dev(case["idx"]), to_dev(case["val"]), to_dev(labels)

    logits64, c = FM.deepfm_forward(p, case["off"], case["idx"], case["val"])
    with torch.no_grad():
        got = model(idx, val)[:, 0].cpu().numpy()
    floor = 0.5 * (c["e"] ** 2).sum((1, 2)) + np.abs(c["first"]) + 1.0     # magnitudes being summed into a logit
    assert rel_err(got, logits64, floor) <= REL

    p0, st = _copy(p), FM.adagrad_state(p)
    out = FM.deepfm_train_step(p, st, case["off"], case["idx"], case["val"], labels.astype(np.float64), lr)
    opt = model.dense_optimizer()
    loss = model.train_step(opt, idx, val, y)
    torch.cuda.synchronize()
    assert abs(float(loss) - out["loss"]) <= 1e-4 * abs(out["loss"])
    emb = model.embedding
    if sharded:                                   # world size 1: the shard is the whole table
        got_t, got_w = emb.table.cpu().numpy()[:case["N"]], emb.w1.cpu().numpy()[:case["N"]]
    else:
        got_t, got_w = emb.table.cpu().numpy(), emb.w1.cpu().numpy()
    assert rel_err(got_t, p["table"], np.abs(p0["table"]).max()) <= REL
    assert rel_err(got_w, p["w1"], np.abs(p0["w1"]).max() + 1e-3) <= REL
    untouched = np.ones(case["N"], bool)
    untouched[out["rows"]] = False
    assert np.array_equal(got_t[untouched], case["table"][untouched])
    assert abs(float(emb.bias.detach()) - p["bias"]) <= 1e-5
    assert _tower_err(model.dnn, p["W"], p["b"], p0["W"], True) <= REL


@pytest.mark.parametrize("clip", [100.0, 0.05])
def test_dcn_logits_and_one_step(pkg, cuda, clip):
    B, K, hidden, L, lr = 192, 16, (64, 32), 3, 0.05
    case = make_case(12, B, ROWS, K, weighted=True, prune=True)
    rng, F = case["rng"], case["F"]
    d = F * K
    labels = (rng.random(B) < 0.3).astype(np.float32)
    W, b = FM.make_tower(rng, d, hidden, 0, np.float32)
    p32 = dict(table=case["table"], cross_w=tfs.truncated_normal(rng, (L, d), 0.1), cross_b=tfs.truncated_normal(rng, (L, d), 0.1),
               W=W, b=b, Wl=FM.glorot_uniform(rng, d + hidden[-1], 1, np.float32), bl=np.asarray([0.02], np.float32))
    p = {k: ([a.astype(np.float64) for a in v] if isinstance(v, list) else np.asarray(v, np.float64)) for k, v in p32.items()}
    model = pkg.DCN(F, K, ROWS, cross_layer_num=L, hidden_units=hidden, learning_rate=lr, clip_norm=clip).train()
    model.embedding.load_tables(case["table"], None)
    _load_tower(model.deep, W, b, False)
    with torch.no_grad():
        model.cross.cross_w.copy_(to_dev(p32["cross_w"]))
        model.cross.cross_b.copy_(to_dev(p32["cross_b"]))
        model.logits.weight.copy_(to_dev(p32["Wl"].T))
        model.logits.bias.copy_(to_dev(p32["bl"]))
    idx, val, y = to_dev(case["idx"]), to_dev(case["val"]), to_dev(labels)

    logits64, c = FM.dcn_forward(p, case["off"], case["idx"], case["val"])
    with torch.no_grad():
        got = model(idx, val)[:, 0].cpu().numpy()
    floor = np.abs(c["m"] * p["Wl"][:, 0][None, :]).sum(1) + 1e-3
    assert rel_err(got, logits64, floor) <= REL

    p0, st = _copy(p), FM.adagrad_state(p)
    out = FM.dcn_train_step(p, st, case["off"], case["idx"], case["val"], labels.astype(np.float64), lr,
                            clip_norm=clip, clip_tables=True)
    if clip < 1.0:      # the clip must actually bind on at least one column's table gradient
        off = list(case["off"]) + [case["N"]]
        norms = [np.linalg.norm(out["G_unclipped"][(out["rows"] >= off[f]) & (out["rows"] < off[f + 1])]) for f in range(F)]
        assert max(norms) > clip, "clip too loose for this test to mean anything"
    opt = model.dense_optimizer()
    loss = model.train_step(opt, idx, val, y)
    torch.cuda.synchronize()
    assert abs(float(loss) - out["loss"]) <= 1e-5 * abs(out["loss"]) + 1e-7
    got_t = model.embedding.table.cpu().numpy()
    assert rel_err(got_t, p["table"], np.abs(p0["table"]).max()) <= REL
    assert rel_err(model.cross.cross_w.detach().cpu().numpy(), p["cross_w"], np.abs(p0["cross_w"]).max()) <= REL
    assert rel_err(model.cross.cross_b.detach().cpu().numpy(), p["cross_b"], np.abs(p0["cross_b"]).max()) <= REL
    assert _tower_err(model.deep, p["W"], p["b"], p0["W"], False) <= REL
    assert rel_err(model.logits.weight.detach().cpu().numpy().T, p["Wl"], np.abs(p0["Wl"]).max()) <= REL
    assert rel_err(model.logits.bias.detach().cpu().numpy(), p["bl"], 1.0) <= REL


def test_models_validate_arguments(pkg, cuda):
    with pytest.raises(ValueError):
        pkg.DeepFM(2, 8, [4, 4], dnn_hidden_units=())
    with pytest.raises(ValueError):
        pkg.DeepFM(2, 8, [4, 4], loss_reduction="none")
    with pytest.raises(ValueError):
        pkg.DCN(0, 8, [])
