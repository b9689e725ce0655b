"""GPU parity of the DCN cross stack (forward + backward) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import deepctr_oracle as O
from tests._util import REL, rel_err, to_dev

pytestmark = pytest.mark.gpu

CASES = [(1, 1, 1), (5, 3, 2), (33, 51, 2), (64, 128, 3), (100, 429, 6), (129, 624, 6),
         (40, 1024, 4), (17, 1000, 1), (300, 52, 8), (9, 7, 32),
         # enough samples for every warp's ring of the register-accumulating backward to wrap several times
         (12007, 624, 6), (20011, 128, 3), (9001, 64, 8), (7000, 24, 5), (6001, 768, 4)]


def _case(B, d, L, seed=21):
    rng = np.random.default_rng(seed)
    x0 = (rng.standard_normal((B, d)) * 0.5).astype(np.float32)
    w = np.clip(rng.standard_normal((L, d)) * 0.1, -0.2, 0.2).astype(np.float32)
    b = np.clip(rng.standard_normal((L, d)) * 0.1, -0.2, 0.2).astype(np.float32)
    dy = rng.standard_normal((B, d)).astype(np.float32)
    return x0, w, b, dy


@pytest.mark.parametrize("B,d,L", CASES)
def test_cross_forward_backward_parity(pkg, cuda, B, d, L):
    x0, w, b, dy = _case(B, d, L)
    net = pkg.CrossNetwork(d, L).train()
    with torch.no_grad():
        net.cross_w.copy_(to_dev(w))
        net.cross_b.copy_(to_dev(b))
    tx0 = to_dev(x0).requires_grad_(True)
    xL = net(tx0)
    xL.backward(to_dev(dy))
    torch.cuda.synchronize()
    f64 = [a.astype(np.float64) for a in (x0, w, b, dy)]
    xL64, s64 = O.cross_forward(*f64[:3])
    dx0_64, dw64, db64 = O.cross_backward(*f64)
    # floors: the magnitude of the terms being summed
    assert rel_err(xL.detach().cpu().numpy(), xL64, np.abs(xL64).max()) <= REL
    assert rel_err(tx0.grad.cpu().numpy(), dx0_64, np.abs(dx0_64).max()) <= REL
    assert rel_err(net.cross_w.grad.cpu().numpy(), dw64, np.abs(dw64).max()) <= REL
    assert rel_err(net.cross_b.grad.cpu().numpy(), db64, np.abs(db64).max()) <= REL


def test_cross_backward_recompute_equals_saved(pkg, cuda):
    """s may be passed from the forward or recomputed (NULL): same gradients to 1e-6."""
    from dir_b200 import _lib
    B, d, L = 77, 624, 6
    x0, w, b, dy = (to_dev(a) for a in _case(B, d, L))
    lib = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    xL, s = torch.empty_like(x0), torch.empty((B, L), device="cuda")
    _lib.check(lib.dir_cross_fwd(x0.data_ptr(), w.data_ptr(), b.data_ptr(), B, d, L, xL.data_ptr(),
                                 s.data_ptr(), st), "fwd")
    ws = torch.empty(lib.dir_cross_bwd_workspace_bytes(B, d, L), dtype=torch.uint8, device="cuda")
    outs = []
    for sp in (s.data_ptr(), None):
        dx0, dw, db = torch.empty_like(x0), torch.empty_like(w), torch.empty_like(b)
        _lib.check(lib.dir_cross_bwd(x0.data_ptr(), w.data_ptr(), b.data_ptr(), dy.data_ptr(), sp, B, d, L,
                                     dx0.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(),
                                     ws.numel(), st), "bwd")
        outs.append((dx0.clone(), dw.clone(), db.clone()))
    for a, c in zip(*outs):
        assert torch.allclose(a, c, rtol=1e-5, atol=1e-6 * float(a.abs().max()))


def test_cross_deterministic_and_kat(pkg, cuda):
    B, d, L = 4096, 624, 6
    x0, w, b, dy = _case(B, d, L, seed=22)
    net = pkg.CrossNetwork(d, L).train()
    grads = []
    for _ in range(2):
        with torch.no_grad():
            net.cross_w.copy_(to_dev(w))
            net.cross_b.copy_(to_dev(b))
        net.zero_grad()
        tx0 = to_dev(x0).requires_grad_(True)
        net(tx0).backward(to_dev(dy))
        grads.append((tx0.grad.clone(), net.cross_w.grad.clone(), net.cross_b.grad.clone()))
    for a, c in zip(*grads):
        assert torch.equal(a, c)
    # KAT-4: w = 0 -> x_L = x0 + sum_l b_l.  The kernel adds sum_l b_l (layer order) to x0 in one step, the
    # reference adds b_l layer by layer: equal up to the rounding of L fp32 additions.
    with torch.no_grad():
        net.cross_w.zero_()
        out = net(to_dev(x0)).cpu().numpy()
    want = x0.copy()
    for l in range(L):
        want = (x0 * np.float32(0) + b[l]) + want
    beta = np.zeros(d, np.float32)
    for l in range(L):
        beta = beta + b[l]
    assert np.array_equal(out, x0 + beta[None, :])
    assert np.abs(out - want).max() <= L * np.finfo(np.float32).eps * np.abs(want).max()


def test_cross_errors(pkg, cuda):
    with pytest.raises(ValueError):
        pkg.CrossNetwork(2000, 2)
    net = pkg.CrossNetwork(8, 2)
    with pytest.raises(ValueError):
        net(torch.zeros((3, 9), device="cuda"))
    assert net(torch.zeros((0, 8), device="cuda")).shape == (0, 8)
