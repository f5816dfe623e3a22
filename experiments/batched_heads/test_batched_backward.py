"""python -m pytest experiments/batched_heads/test_batched_backward.py -q   (CPU)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from tensoir_b200 import heads                                  # noqa: E402
from batched_backward import heads_backward_batched             # noqa: E402


def _pad(t, dim):
    return torch.nn.functional.pad(t, (0, 0, 0, 4 - t.shape[0])) if dim == 0 else \
        torch.nn.functional.pad(t, (0, 4 - t.shape[1]))


def test_batched_equals_per_head():
    """radiance (sigmoid, 3 out), BRDF (sigmoid, 4), normal (tanh, 3), BRDF at the jittered points."""
    torch.manual_seed(11)
    n, F, aC3, hid = 700, 27, 144, 128
    in_dim = F + 3 + 2 * F * 2 + 2 * 3 * 2
    basis = torch.randn(F, aC3) * 0.1
    specs = [(0, 3), (0, 4), (1, 3), (0, 4)]                 # (act, out_dim)
    per, stacked = [], {k: [] for k in ("g", "out", "xl", "inp", "h1", "h2", "w0", "w1", "w2")}
    for act, od in specs:
        w0, w1, w2 = torch.randn(hid, in_dim) * 0.1, torch.randn(hid, hid) * 0.1, torch.randn(od, hid) * 0.1
        xl, inp = torch.randn(n, aC3), torch.randn(n, in_dim)
        h1, h2 = torch.relu(torch.randn(n, hid)), torch.relu(torch.randn(n, hid))
        out = torch.tanh(torch.randn(n, od)) if act else torch.sigmoid(torch.randn(n, od))
        g = torch.randn(n, od)
        per.append(heads._head_backward(act, "none", g, out, xl, inp, h1, h2, w0, w1, w2, basis, None, None, None))
        for k, v in (("g", _pad(g, 1)), ("out", _pad(out, 1)), ("xl", xl), ("inp", inp), ("h1", h1), ("h2", h2),
                     ("w0", w0), ("w1", w1), ("w2", _pad(w2, 0))):
            stacked[k].append(v)
    st = {k: torch.stack(v) for k, v in stacked.items()}
    is_tanh = torch.tensor([bool(a) for a, _ in specs])
    gw0, gb0, gw1, gb1, gw2, gb2, gbasis, gxl = heads_backward_batched(
        is_tanh, st["g"], st["out"], st["xl"], st["inp"], st["h1"], st["h2"], st["w0"], st["w1"], st["w2"], basis)
    tol = dict(rtol=1e-4, atol=1e-4)
    assert torch.allclose(gbasis, sum(p[6] for p in per), **tol)
    for h, ((act, od), p) in enumerate(zip(specs, per)):
        assert torch.allclose(gw0[h], p[0], **tol) and torch.allclose(gb0[h], p[1], **tol)
        assert torch.allclose(gw1[h], p[2], **tol) and torch.allclose(gb1[h], p[3], **tol)
        assert torch.allclose(gw2[h, :od], p[4], **tol) and torch.allclose(gb2[h, :od], p[5], **tol)
        assert torch.allclose(gxl[h], p[8], **tol)
        if od < 4:
            assert float(gw2[h, od:].abs().max()) == 0.0
