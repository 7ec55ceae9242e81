"""GPU: differentiable head building blocks (fp32 GEMM, row-normalise, softmax-CE, Tip cache logits) against torch
autograd through the oracle's restatement of the reference heads."""
import pytest
import torch
import torch.nn.functional as F

from clip_gp_b200 import ops, tc
from oracle import heads as oh
from tests.helpers import max_err, rel_err, within

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (48, 37, 128), (128, 1000, 512), (130, 10000, 77), (700, 300, 1024), (128, 512, 10000)])
def test_matmul_nt_forward_backward(M, N, K):
    g = torch.Generator().manual_seed(M * N + K)
    A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    dC = torch.randn(M, N, generator=g)
    Ar, Br = A.clone().requires_grad_(True), B.clone().requires_grad_(True)
    (0.7 * Ar @ Br.t()).backward(dC)
    Ad, Bd = A.cuda().requires_grad_(True), B.cuda().requires_grad_(True)
    C = ops.matmul_nt(Ad, Bd, 0.7)
    C.backward(dC.cuda())
    assert within(C, (0.7 * A @ B.t()), 1e-5)
    assert within(Ad.grad, Ar.grad, 1e-5) and within(Bd.grad, Br.grad, 1e-5)


@pytest.mark.parametrize("M,N,K", [(48, 36, 128), (128, 1000, 512), (130, 10000, 76), (700, 300, 1024), (33, 512, 10000), (50, 37, 128)])
def test_matmul_nt_tf32_forward_backward(M, N, K):
    """Tensor-core (kind::tf32) autograd product and both in-place-transposed adjoints; (50,37,128): N not a multiple of 4 -> FFMA."""
    g = torch.Generator().manual_seed(M * N + K)
    A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    dC = torch.randn(M, N, generator=g)
    Ar, Br = A.double().requires_grad_(True), B.double().requires_grad_(True)
    (0.7 * Ar @ Br.t()).backward(dC.double())
    Ad, Bd = A.cuda().requires_grad_(True), B.cuda().requires_grad_(True)
    C = ops.matmul_nt(Ad, Bd, 0.7, "tf32")
    C.backward(dC.cuda())
    tol = 1e-5 if N % 4 else 2e-3                              # TF32: 10-bit mantissa operands, fp32 accumulate; normwise
    assert max_err(C, 0.7 * A.double() @ B.double().t()) < tol
    assert max_err(Ad.grad, Ar.grad) < tol and max_err(Bd.grad, Br.grad) < tol
    assert ops.tf32_ok(Ad, Bd) == (N % 4 == 0)


def test_adapter_head_loss_and_grads_match_oracle():
    """adapter.py:401-428 composed from clipgp ops == oracle.heads.adapter_mc_ce (autograd)."""
    g = torch.Generator().manual_seed(1)
    B, D, C, S = 48, 128, 37, 4
    f = torch.randn(B, D, generator=g); y = torch.randint(0, C, (B,), generator=g)
    W = (torch.eye(D) + 0.05 * torch.randn(D, D, generator=g))
    P = torch.randn(S, C, D, generator=g)
    Wr, Pr = W.clone().requires_grad_(True), P.clone().requires_grad_(True)
    ref = oh.adapter_mc_ce(f, y, Wr, Pr, 30.0)
    ref.backward()
    Wd, Pd = W.cuda().requires_grad_(True), P.cuda().requires_grad_(True)
    f_hat = ops.row_normalize(ops.matmul_nt(f.cuda(), Wd))
    p_hat = ops.row_normalize(Pd).reshape(S * C, D)
    logits = ops.matmul_nt(f_hat, p_hat, 30.0).view(B * S, C)           # row (b, s)
    loss = ops.cross_entropy(logits, y.cuda(), rows_per_label=S)
    loss.backward()
    assert float(loss) == pytest.approx(float(ref), rel=1e-5)
    assert within(Wd.grad, Wr.grad, 1e-4) and within(Pd.grad, Pr.grad, 1e-4)
    # logit-mean form (adapter.py:247-249)
    lm = ops.matmul_nt(f_hat, p_hat, 30.0).view(B, S, C).mean(1)
    assert within(lm, oh.adapter_logits(f, W, P, 30.0), 1e-5)


@pytest.mark.parametrize("B,N_tr,C,D", [(16, 48, 12, 64), (128, 1600, 100, 1024), (33, 407, 37, 128)])
def test_tip_cache_logits_forward_backward(B, N_tr, C, D):
    g = torch.Generator().manual_seed(B + N_tr)
    f = F.normalize(torch.randn(B, D, generator=g), dim=-1)
    keys = F.normalize(torch.randn(N_tr, D, generator=g), dim=-1)
    lab = torch.randint(0, C, (N_tr,), generator=g)
    clip = 5.0 * torch.randn(B, C, generator=g)
    dout = torch.randn(B, C, generator=g)
    kr = keys.clone().requires_grad_(True)
    ref = oh.tip_logits(f, kr, oh.tip_cache_vals(lab, C), clip, 2.0, 20.0)
    ref.backward(dout)
    kd = keys.cuda().requires_grad_(True)
    out = ops.tip_logits(f.cuda(), kd, lab.cuda(), clip.cuda(), 2.0, 20.0, C)
    out.backward(dout.cuda())
    assert within(out, ref, 1e-5) and within(kd.grad, kr.grad, 1e-4)
    # fused tensor-core evaluation form (affinity never materialised); split operands -> fp32-grade
    order = torch.argsort(lab, stable=True)
    ks, ls = keys[order].cuda(), lab[order].cuda()
    o3 = clip.cuda().clone()
    from clip_gp_b200 import _lib
    fa, kb, li = tc.cast_bf16(f.cuda(), tc.SPLIT_A), tc.cast_bf16(ks, tc.SPLIT_B), ls.to(torch.int32).contiguous()   # keep alive
    _lib.check(_lib.load().clipgp_tc_tip_logits(fa.data_ptr(), B, kb.data_ptr(), N_tr, 3 * D, li.data_ptr(), 2.0, 20.0, o3.data_ptr(), C,
                                                _lib.stream_ptr(o3.device)), "clipgp_tc_tip_logits")
    assert rel_err(o3, ref) < 1e-3


@pytest.mark.parametrize("precision,tol", [("bf16x3", 1e-3), ("bf16", 3e-2)])
def test_tip_cache_logits_tensor_core_precisions(precision, tol):
    """Tip-Adapter-F step on the tcgen05 GEMMs (affinity and key gradient) against the oracle; bf16 carries a stated tolerance."""
    g = torch.Generator().manual_seed(11)
    B, N_tr, C, D = 128, 1600, 100, 256
    mu = torch.randn(C, D, generator=g)
    lab = torch.arange(C).repeat_interleave(16)
    keys = F.normalize(mu[lab] + 2.0 * torch.randn(N_tr, D, generator=g), dim=-1)
    yb = torch.randint(0, C, (B,), generator=g)
    f = F.normalize(mu[yb] + 2.0 * torch.randn(B, D, generator=g), dim=-1)
    clip = 5.0 * torch.randn(B, C, generator=g)
    dout = torch.randn(B, C, generator=g)
    kr = keys.clone().requires_grad_(True)
    ref = oh.tip_logits(f, kr, oh.tip_cache_vals(lab, C), clip, 2.0, 20.0)
    ref.backward(dout)
    kd = keys.cuda().requires_grad_(True)
    out = ops.tip_logits(f.cuda(), kd, lab.cuda(), clip.cuda(), 2.0, 20.0, C, precision)
    out.backward(dout.cuda())
    # logits: elementwise; key gradient: norm-wise (bf16: the stated tolerance) and, for the split operands, elementwise with the
    # floor at 1e-2 of the tensor's largest entry: every entry is a sum over the batch with cancellation, and small entries carry
    # the 2^-16 product rounding of the large terms (norm-wise the split path is at 1e-6)
    assert rel_err(out, ref) < tol and max_err(kd.grad, kr.grad) < tol
    if precision == "bf16x3":
        assert within(kd.grad, kr.grad, 1e-4)
    with pytest.raises(ValueError):
        ops.tip_logits(f.cuda(), kd, lab.cuda(), clip.cuda(), 2.0, 20.0, C, "fp8")


@pytest.mark.parametrize("N_tr,C,sort", [(600, 300, True), (777, 40, False), (4096, 4096, True)])
def test_tip_fused_epilogue_class_windows(N_tr, C, sort):
    """Fused tcgen05 Tip epilogue outside its fast case: more than 32 classes per 256-key tile (few-shot caches) and unsorted keys
    go through the direct-reduction fallback; results equal the one-hot formulation."""
    from clip_gp_b200 import _lib
    g = torch.Generator().manual_seed(N_tr)
    B, D = 150, 128
    lab = torch.randint(0, C, (N_tr,), generator=g) if C < N_tr else torch.arange(C)
    if sort:
        lab = lab.sort().values
    keys = F.normalize(torch.randn(N_tr, D, generator=g), dim=-1)
    f = F.normalize(torch.randn(B, D, generator=g), dim=-1)
    clip = 5.0 * torch.randn(B, C, generator=g)
    ref = oh.tip_logits(f, keys, oh.tip_cache_vals(lab, C), clip, 5.0, 10.0)
    o3 = clip.cuda().clone()
    fa, kb, li = tc.cast_bf16(f.cuda(), tc.SPLIT_A), tc.cast_bf16(keys.cuda(), tc.SPLIT_B), lab.to(torch.int32).cuda().contiguous()
    _lib.check(_lib.load().clipgp_tc_tip_logits(fa.data_ptr(), B, kb.data_ptr(), N_tr, 3 * D, li.data_ptr(), 5.0, 10.0, o3.data_ptr(), C,
                                                _lib.stream_ptr(o3.device)), "clipgp_tc_tip_logits")
    assert rel_err(o3, ref) < 1e-3


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("bf16x3", 1e-3)])
def test_tip_engine_matches_reference_loop(precision, tol):
    """TipAdapterEngine (fused, graph-captured Tip-Adapter-F step) against the reference formulation on the CPU: nn.Linear keys,
    AdamW(lr, eps=1e-4), per-step CosineAnnealingLR (tip_adapter.py:229-269), several steps incl. a smaller last batch."""
    from clip_gp_b200.tip_engine import TipAdapterEngine
    g = torch.Generator().manual_seed(21)
    N_tr, C, D, steps = 320, 20, 64, 5
    mu = torch.randn(C, D, generator=g)
    lab = torch.arange(C).repeat_interleave(16)
    keys = F.normalize(mu[lab] + 2.0 * torch.randn(N_tr, D, generator=g), dim=-1)
    batches = []
    for i in range(steps):
        B = 48 if i != 3 else 24
        yb = torch.randint(0, C, (B,), generator=g)
        f = F.normalize(mu[yb] + 2.0 * torch.randn(B, D, generator=g), dim=-1)
        batches.append((f, 100.0 * f @ F.normalize(mu, dim=-1).t(), yb))
    w = keys.clone().requires_grad_(True)
    opt = torch.optim.AdamW([w], lr=1e-2, eps=1e-4)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, steps)
    vals = oh.tip_cache_vals(lab, C)
    ref_losses = []
    for f, clip, yb in batches:
        loss = F.cross_entropy(oh.tip_logits(f, w, vals, clip, 2.0, 20.0), yb)
        opt.zero_grad(); loss.backward(); opt.step(); sched.step()
        ref_losses.append(float(loss))
    eng = TipAdapterEngine(keys.cuda(), lab.cuda(), C, 48, 2.0, 20.0, lr=1e-2, eps=1e-4, total_steps=steps, precision=precision)
    losses = [float(eng.train_step(f.cuda(), clip.cuda(), yb.cuda())) for f, clip, yb in batches]
    assert losses == pytest.approx(ref_losses, rel=max(tol, 1e-4), abs=2e-6)
    assert within(eng.keys, w, max(tol, 5e-5))


def test_tip_hyperparameter_search_matches_oracle():
    from clip_gp_b200 import heads
    g = torch.Generator().manual_seed(7)
    B, N_tr, C, D = 200, 160, 10, 64
    mu = torch.randn(C, D, generator=g)
    lab = torch.arange(C).repeat_interleave(16)
    keys = F.normalize(mu[lab] + 2.0 * torch.randn(N_tr, D, generator=g), dim=-1)
    y = torch.randint(0, C, (B,), generator=g)
    f = F.normalize(mu[y] + 3.0 * torch.randn(B, D, generator=g), dim=-1)
    clip = 100.0 * f @ F.normalize(mu, dim=-1).t()
    bb, ba, acc = oh.tip_search(f, y, keys, oh.tip_cache_vals(lab, C), clip, 2.0, 20.0)
    b2, a2, acc2 = heads.tip_search(f.cuda(), y.cuda(), keys.cuda(), lab.cuda(), clip.cuda(), C, 2.0, 20.0)
    assert (b2, a2) == (bb, ba) and acc2 == pytest.approx(acc)
    # image chunks + tensor-core affinity: same arg-max pair (the grid shares ONE affinity per chunk)
    b3, a3, acc3 = heads.tip_search(f.cuda(), y.cuda(), keys.cuda(), lab.cuda(), clip.cuda(), C, 2.0, 20.0, precision="bf16x3", chunk=64)
    assert (b3, a3) == (bb, ba) and acc3 == pytest.approx(acc)


@pytest.mark.parametrize("method", ["uniform", "val_weighted", "top3", "minmax"])
def test_template_weight_initialisation(method):
    """_get_template_weights (adapter.py:48-142): zero-shot accuracy of every template per class from the tcgen05 GEMM + arg-max
    epilogue, then the reference's top3 / minmax / log-softmax post-processing, against the line-by-line oracle."""
    from clip_gp_b200 import heads, synth
    from oracle import heads as oh
    wl = synth.make_workload("small"); shp = wl["shape"]
    cfg = type("Cfg", (), {"adapter": type("A", (), {"template_init_method": method})()})()
    E, f, y = wl["E"], wl["f_train"], wl["y_train"]
    ref = oh.template_weights(method, E, f, y, 100.0)
    w = heads.get_template_weights(cfg, E.cuda(), f.cuda(), y.cuda(), 100.0)
    assert w.shape == (shp.C, shp.T)
    assert torch.allclose(w.sum(-1).cpu(), torch.ones(shp.C), atol=1e-5)
    if method == "uniform":
        assert torch.equal(w.cpu(), ref)
        return
    w_ref, scores_ref = ref
    scores = heads.template_accuracy_scores(E.cuda(), f.cuda(), y.cuda()).cpu()
    if method == "val_weighted":
        assert torch.equal(scores, scores_ref)            # integer hit counts / class counts: bit-exact
    assert torch.allclose(w.cpu(), w_ref, atol=1e-6)
