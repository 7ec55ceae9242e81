"""GPU: weighted-prototype kernels against torch einsum / normalize on the CPU."""
import pytest
import torch
import torch.nn.functional as F

from clip_gp_b200 import ops
from tests.helpers import rel_err, within

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("S,C,T,D", [(1, 3, 1, 4), (4, 100, 8, 1024), (10, 37, 32, 512), (33, 5, 64, 512), (3, 7, 5, 64), (20, 4, 9, 2048)])
def test_prototypes_forward_backward(S, C, T, D):
    g = torch.Generator().manual_seed(S * 100 + C)
    w = torch.rand(S, C, T, generator=g); w = w / w.sum(-1, keepdim=True)
    E = F.normalize(torch.randn(C, T, D, generator=g), dim=-1)
    dP = torch.randn(S, C, D, generator=g)
    wr = w.clone().requires_grad_(True)
    ref = torch.einsum("skm,kmd->skd", wr, E)
    ref.backward(dP)
    wd = w.cuda().requires_grad_(True)
    P = ops.prototypes(wd, E.cuda())
    P.backward(dP.cuda())
    assert within(P, ref, 1e-5) and within(wd.grad, wr.grad, 1e-5)
    P_hat, mh, mr = ops.prototypes_reduced(w.cuda(), E.cuda(), want_hat=True, want_mean_hat=True, want_mean_raw=True)
    ph = F.normalize(ref.detach(), dim=-1)
    assert within(P_hat, ph, 1e-5) and within(mh, ph.mean(0), 1e-5)
    assert within(mr, F.normalize(ref.detach().mean(0), dim=-1), 1e-5)
    # TaskRes residual branch (taskres.py:109-113)
    x = 0.1 * torch.randn(C, D, generator=g)
    t = ph + 0.5 * x.unsqueeze(0); t = t / t.norm(dim=-1, keepdim=True)
    T_hat, tm, _ = ops.prototypes_reduced(w.cuda(), E.cuda(), residual=x.cuda(), alpha=0.5, want_hat=True, want_mean_hat=True)
    assert within(T_hat, t, 1e-5) and within(tm, t.mean(0), 1e-5)


def test_bad_shapes_raise():
    with pytest.raises(RuntimeError):
        ops.prototypes(torch.rand(2, 3, 4).cuda(), torch.rand(3, 4, 6).cuda())     # D % 4 != 0
    with pytest.raises(RuntimeError):
        ops.prototypes(torch.rand(2, 3, 4), torch.rand(3, 4, 8))                   # CPU tensors: no fallback
