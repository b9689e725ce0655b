"""Column feed (dir_expand_features / ColumnFeeder): the widened [B,F] pair is bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("index_dtype", [torch.int32, torch.int64])
@pytest.mark.parametrize("B,sparse,dense", [(1, [0], [1]), (257, [0, 2, 3, 5], [1, 4]), (1000, list(range(26)), list(range(26, 39))),
                                            (64, [], [0, 1, 2]), (64, [2, 0, 1], [])])
def test_expand_features_matches_numpy(pkg, cuda, B, sparse, dense, index_dtype):
    from dir_b200 import _lib
    F = len(sparse) + len(dense)
    rng = np.random.default_rng(5)
    sp = rng.integers(0, 2 ** 31 - 1, size=(B, max(len(sparse), 1)))
    de = rng.random((B, max(len(dense), 1)), dtype=np.float32)
    src = np.zeros(F, np.int32)
    want_i, want_v = np.zeros((B, F), np.int64), np.ones((B, F), np.float32)
    for j, f in enumerate(sparse):
        src[f] = j
        want_i[:, f] = sp[:, j]
    for j, f in enumerate(dense):
        src[f] = -(j + 1)
        want_v[:, f] = de[:, j]
    t_sp = torch.as_tensor(sp).to(index_dtype).cuda()
    t_de, t_src = torch.as_tensor(de).cuda(), torch.as_tensor(src).cuda()
    idx = torch.full((B, F), -7, dtype=torch.int64, device="cuda")
    val = torch.full((B, F), -7.0, dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib().dir_expand_features(
        t_sp.data_ptr(), 4 if index_dtype == torch.int32 else 8, t_de.data_ptr(), t_src.data_ptr(), B, F,
        len(sparse), len(dense), idx.data_ptr(), val.data_ptr(), torch.cuda.current_stream().cuda_stream), "expand")
    torch.cuda.synchronize()
    assert np.array_equal(idx.cpu().numpy(), want_i)
    assert np.array_equal(val.cpu().numpy(), want_v)


def test_expand_features_rejects_bad_arguments(pkg, cuda):
    from dir_b200 import _lib
    lib = _lib.lib()
    assert lib.dir_expand_features(None, 2, None, None, 4, 3, 1, 1, None, None, None) == -22
    assert lib.dir_expand_features(None, 4, None, None, 4, 3, 3, 1, None, None, None) == -22
    assert lib.dir_expand_features(None, 4, None, None, 0, 3, 1, 1, None, None, None) == 0     # empty batch


def test_column_feeder_round_trip(pkg, cuda):
    """ColumnFeeder lands the same [B,F] inputs HostFeeder does, from a third of the bytes."""
    w = pkg.synth.cfg("cfg1", batch=512)
    idx, val, y = pkg.synth.make_inputs(w)
    sp_f = [f for f, n in enumerate(w.rows_per_field) if n > 1]
    de_f = [f for f, n in enumerate(w.rows_per_field) if n == 1]
    slots = [[torch.zeros(idx.shape, dtype=torch.int64, device="cuda"), torch.zeros(val.shape, device="cuda"),
              torch.zeros(y.shape, device="cuda")] for _ in range(2)]
    feeder = pkg.ColumnFeeder(sp_f, de_f, *slots)
    host = [torch.as_tensor(idx[:, sp_f]).to(torch.int32).contiguous().pin_memory(),
            torch.as_tensor(val[:, de_f]).contiguous().pin_memory(), torch.as_tensor(y).pin_memory()]
    assert feeder.bytes_per_batch(host) * 2.9 < idx.nbytes + val.nbytes + y.nbytes
    for slot in (0, 1, 0):
        feeder.prefetch(slot, host)
        got = feeder.wait(slot)
        torch.cuda.synchronize()
        assert np.array_equal(got[0].cpu().numpy(), idx)
        assert np.array_equal(got[1].cpu().numpy(), val)
        assert np.array_equal(got[2].cpu().numpy(), y)
        feeder.release(slot)
    with pytest.raises(ValueError):
        feeder.prefetch(0, [host[0].to(torch.int64), host[1], host[2]])
    with pytest.raises(ValueError):
        pkg.ColumnFeeder([0, 1], [1], *slots)
