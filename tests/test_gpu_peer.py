"""GPU: the fused multi-GPU optimiser step over peer memory (csrc/peer.cu) on ONE device: world = 1 against the plain AdamW kernel, and
a two-rank run emulated with two streams (both kernels are resident at once, the flags live in ordinary device memory), so the
reduce-scatter / all-gather protocol and its epoch flags are exercised without a second GPU.  The real NVLink / CUDA-IPC path is
covered by tools/check_multi_gpu.py under torchrun."""
import ctypes as C

import pytest
import torch

from clip_gp_b200 import _lib

pytestmark = pytest.mark.gpu


def _adamw_ref(p, g, m, v, t, lr0, lr1, n0, b1=0.9, b2=0.999, eps=1e-8, wd=0.0):
    lr = torch.where(torch.arange(p.numel(), device=p.device) < n0, torch.tensor(lr0, device=p.device), torch.tensor(lr1, device=p.device)).double()
    g, p, m, v = g.double(), p.double(), m.double(), v.double()
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
    p = p * (1 - lr * wd) - (lr / bc1) * m / (v.sqrt() / bc2 ** 0.5 + eps)
    return p, m, v


def _args(world, rank, gs, ps, flags, m, v, n, n0, lr, step, local, loss, status):
    a = _lib.PeerArgs()
    a.world, a.rank = world, rank
    for q in range(world):
        a.g[q], a.p[q], a.flags[q] = gs[q].data_ptr(), ps[q].data_ptr(), flags[q].data_ptr()
    a.m, a.v, a.n, a.n_group0 = m.data_ptr(), v.data_ptr(), n, n0
    a.lr_dev, a.beta1, a.beta2, a.eps, a.weight_decay = lr.data_ptr(), 0.9, 0.999, 1e-8, 0.0
    a.step, a.local, a.loss_out, a.status = step.data_ptr(), local.data_ptr(), loss.data_ptr(), status.data_ptr()
    a.timeout_ns = int(2e9)
    return a


@pytest.mark.parametrize("n", [4096, 10007, 3])
def test_peer_adamw_world1_matches_adamw_formula(n):
    dev = torch.device("cuda")
    lib = _lib.load()
    gen = torch.Generator(device="cpu").manual_seed(n)
    g = torch.randn(n + 1, generator=gen).to(dev); p = torch.randn(n, generator=gen).to(dev)
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    flags = torch.zeros(2 * _lib.PEER_MAX, dtype=torch.int64, device=dev)
    lr = torch.tensor([1e-2, 1e-3], device=dev); step = torch.ones(1, dtype=torch.int64, device=dev)
    local = torch.zeros(2, dtype=torch.int64, device=dev); loss = torch.zeros(1, device=dev); status = torch.zeros(1, dtype=torch.int32, device=dev)
    n0 = n // 3
    pr, mr, vr = p.clone(), m.clone(), v.clone()
    a = _args(1, 0, [g], [p], [flags], m, v, n, n0, lr, step, local, loss, status)
    for t in (1, 2, 3):
        step.fill_(t)
        _lib.check(lib.clipgp_peer_adamw(C.byref(a), _lib.stream_ptr(dev)), "peer_adamw")
        pr, mr, vr = _adamw_ref(pr, g[:n], mr, vr, t, 1e-2, 1e-3, n0)
    torch.cuda.synchronize()
    assert int(status) == 0 and int(local[0]) == 3
    assert float((p.double() - pr).abs().max()) < 1e-6 and float(loss) == pytest.approx(float(g[n]))


def test_peer_adamw_two_ranks_emulated_on_two_streams():
    dev = torch.device("cuda")
    lib = _lib.load()
    n, n0, W = 50003, 20000, 2
    gen = torch.Generator(device="cpu").manual_seed(5)
    p0 = torch.randn(n, generator=gen)
    gs = [torch.zeros(n + 1, device=dev) for _ in range(W)]
    ps = [p0.clone().to(dev) for _ in range(W)]
    flags = [torch.zeros(2 * _lib.PEER_MAX, dtype=torch.int64, device=dev) for _ in range(W)]
    ms = [torch.zeros(n, device=dev) for _ in range(W)]; vs = [torch.zeros(n, device=dev) for _ in range(W)]
    lr = torch.tensor([1e-2, 1e-3], device=dev); step = torch.ones(1, dtype=torch.int64, device=dev)
    locs = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(W)]
    losses = [torch.zeros(1, device=dev) for _ in range(W)]; stats = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(W)]
    args = [_args(W, r, gs, ps, flags, ms[r], vs[r], n, n0, lr, step, locs[r], losses[r], stats[r]) for r in range(W)]
    streams = [torch.cuda.Stream(dev) for _ in range(W)]
    pr, mr, vr = p0.clone().to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    for t in (1, 2, 3):
        gt = [torch.randn(n + 1, generator=gen).to(dev) for _ in range(W)]
        for r in range(W):
            gs[r].copy_(gt[r])
        step.fill_(t)
        torch.cuda.synchronize()
        for r in range(W):
            with torch.cuda.stream(streams[r]):
                _lib.check(lib.clipgp_peer_adamw(C.byref(args[r]), _lib.stream_ptr(dev)), "peer_adamw")
        torch.cuda.synchronize()
        pr, mr, vr = _adamw_ref(pr, (gt[0] + gt[1])[:n], mr, vr, t, 1e-2, 1e-3, n0)
        assert [int(s) for s in stats] == [0, 0]
        assert torch.equal(ps[0], ps[1])                                   # every rank ends the step with identical parameters
        assert float((ps[0].double() - pr).abs().max()) < 1e-6
        assert float(losses[0]) == pytest.approx(float(gt[0][n] + gt[1][n])) and float(losses[1]) == float(losses[0])
