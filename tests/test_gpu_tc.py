"""GPU: the tcgen05/TMEM/TMA GEMM against torch fp32 matmul of the same bf16-rounded operands, and its fused
calibration epilogue against the oracle metrics on the same logits."""
import pytest
import torch

from clip_gp_b200 import metrics as gm
from clip_gp_b200 import tc
from oracle import metrics as om
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


def _rand(R, K, g, scale=1.0):
    return (scale * torch.randn(R, K, generator=g)).cuda()


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 512), (300, 1000, 512), (128, 10000, 512), (1000, 512, 128),
                                   (77, 40, 72), (129, 257, 200), (4096, 1000, 1024)])
def test_gemm_store_matches_fp32_matmul_of_bf16_operands(M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A, B = _rand(M, K, g), _rand(N, K, g)
    Ab, Bb = tc.cast_bf16(A), tc.cast_bf16(B)
    assert torch.equal(Ab, A.to(torch.bfloat16)) and torch.equal(Bb, B.to(torch.bfloat16))
    C = tc.gemm_store(Ab, Bb, alpha=0.5)
    ref = 0.5 * (Ab.double() @ Bb.double().t())
    assert float((C.double() - ref).abs().max()) < 1e-4 * float(ref.abs().max())     # fp32 accumulation order only


def test_split_operands_reach_fp32_accuracy():
    g = torch.Generator().manual_seed(3)
    A, B = _rand(200, 512, g), _rand(1000, 512, g)
    A = torch.nn.functional.normalize(A, dim=-1); B = torch.nn.functional.normalize(B, dim=-1)
    ref = 100.0 * (A.double() @ B.double().t())
    C1 = tc.gemm_store(tc.cast_bf16(A), tc.cast_bf16(B), alpha=100.0)
    C3 = tc.gemm_store(tc.cast_bf16(A, tc.SPLIT_A), tc.cast_bf16(B, tc.SPLIT_B), alpha=100.0)
    e1 = float((C1.double() - ref).abs().max()); e3 = float((C3.double() - ref).abs().max())
    assert e3 < 1e-3 * float(ref.abs().max()) and e3 < 2e-3          # the fp32 gate of BASELINE.json (1e-3 relative)
    assert e1 < 0.1                                                   # stated bf16 tolerance: |dlogit| < 0.1 at scale 100
    assert e3 < e1 / 20


def test_k_wrap_accumulates_mc_samples_in_tmem():
    g = torch.Generator().manual_seed(4)
    S, D, C, M = 5, 128, 300, 260
    A = _rand(M, D, g); P = _rand(S * C, D, g).view(S, C, D)
    Ab = tc.cast_bf16(A)
    Bcat = tc.cast_bf16(P.permute(1, 0, 2).reshape(C, S * D).contiguous())      # row c = [p_1c | ... | p_Sc]
    C_wrap = tc.gemm_store(Ab, Bcat, alpha=1.0 / S)
    ref = (Ab.double().unsqueeze(0) @ P.to(torch.bfloat16).double().transpose(1, 2)).mean(0)
    assert float((C_wrap.double() - ref).abs().max()) < 1e-4 * float(ref.abs().max())


@pytest.mark.parametrize("M,C,D", [(128, 256, 64), (1000, 1000, 512), (333, 37, 128), (5000, 1000, 512)])
def test_fused_calibration_epilogue(M, C, D):
    g = torch.Generator().manual_seed(M + C)
    mu = torch.randn(C, D, generator=g)
    y = torch.randint(0, C, (M,), generator=g)
    f = torch.nn.functional.normalize(mu[y] + 1.5 * torch.randn(M, D, generator=g), dim=-1).cuda()
    P = torch.nn.functional.normalize(mu, dim=-1).cuda()
    y[::5] = torch.randint(0, C, (len(y[::5]),), generator=g)
    fb, Pb = tc.cast_bf16(f), tc.cast_bf16(P)
    conf, correct, hist, logits = tc.logits_calibration(fb, Pb, 30.0, y.cuda(), 10, want_logits=True)
    ref = tc.gemm_store(fb, Pb, 30.0)
    assert torch.equal(logits, ref)                                   # the optional logits copy is the same accumulator
    lg = logits.cpu()
    conf_ref, pred_ref, cor_ref = om.confidence(lg, y)
    assert int(hist[3, 0]) == int(cor_ref.sum()) == int(correct.sum())                 # bit-exact top-1 count
    assert torch.allclose(conf.cpu(), conf_ref, rtol=2e-5, atol=1e-7)                  # exp2-based online softmax
    e, b = om.compute_ece_with_bins(lg, y)
    gap = float((conf_ref[:, None] - torch.linspace(0, 1, 11)[None]).abs().min())
    cnt = gm.counters_from_hist(hist, M)
    ece, bins = gm.ece_from_counters(cnt)
    if gap > 1e-4:
        assert bins["bin_count"] == b["bin_count"]
    assert sum(bins["bin_count"]) == M
    assert ece == pytest.approx(e, rel=1e-3, abs=1e-3)


def _unsplit(op, K, Kp, mode):
    """Reassemble the fp32 value a (split) bf16 operand represents: hi (+ lo)."""
    op = op.float()
    if mode == tc.PLAIN:
        return op[:, :K]
    hi = op[:, :K]
    lo = op[:, 2 * Kp:2 * Kp + K] if mode == tc.SPLIT_A else op[:, Kp:Kp + K]
    return hi + lo


@pytest.mark.parametrize("R,K", [(128, 512), (48, 148), (33, 37), (130, 1000), (257, 64)])
@pytest.mark.parametrize("mode,modeT", [(tc.PLAIN, tc.PLAIN), (tc.SPLIT_A, tc.SPLIT_B), (tc.SPLIT_B, tc.SPLIT_A)])
def test_dual_layout_cast(R, K, mode, modeT):
    g = torch.Generator().manual_seed(R * 7 + K)
    x = _rand(R, K, g)
    out, outT = tc.cast_bf16_dual(x, mode, modeT)
    Kp, Rp = (K + 7) // 8 * 8, (R + 7) // 8 * 8
    ref_o = tc.cast_bf16(x, mode)                                       # the single-layout kernel is the reference layout
    seg = 3 if mode else 1
    for gseg in range(seg):
        assert torch.equal(out[:, gseg * Kp:gseg * Kp + K], ref_o[:, gseg * K:(gseg + 1) * K])
    ref_t = tc.cast_bf16(x.t().contiguous(), modeT)
    segT = 3 if modeT else 1
    for gseg in range(segT):
        assert torch.equal(outT[:, gseg * Rp:gseg * Rp + R], ref_t[:, gseg * R:(gseg + 1) * R])
    tol = 2.0 ** -15 if mode else 2.0 ** -8
    assert float((_unsplit(out, K, Kp, mode) - x).abs().max()) <= tol * float(x.abs().max())


@pytest.mark.parametrize("B,S,C", [(16, 3, 12), (48, 4, 37), (128, 10, 1000), (130, 1, 1000), (200, 2, 1500)])
@pytest.mark.parametrize("mode", [tc.PLAIN, tc.SPLIT_A])
@pytest.mark.parametrize("fused", [True, False])
def test_two_phase_softmax_ce_operands(B, S, C, mode, fused):
    """Loss and dlogits (both operand layouts) of the tensor-core step against torch cross_entropy + autograd."""
    g = torch.Generator().manual_seed(B + S + C)
    logits = (8.0 * torch.randn(B, S * C, generator=g)).cuda().requires_grad_(True)
    y = torch.randint(0, C, (B,), generator=g).cuda()
    ref = torch.nn.functional.cross_entropy(logits.view(B * S, C), y.repeat_interleave(S), reduction="mean")
    ref.backward()
    loss, out, outT = tc.softmax_ce_operands(logits.detach(), y, S, 1.0 / (B * S), mode, fused=fused)
    assert float(loss) == pytest.approx(float(ref), rel=1e-5)
    SC = S * C
    SCp, Bp = (SC + 7) // 8 * 8, (B + 7) // 8 * 8
    gref = logits.grad
    tol = (2.0 ** -15 if mode else 2.0 ** -8) * float(gref.abs().max()) + 1e-9
    assert float((_unsplit(out, SC, SCp, mode) - gref).abs().max()) <= tol
    assert float((_unsplit(outT, B, Bp, mode) - gref.t()).abs().max()) <= tol


@pytest.mark.parametrize("M,N,K", [(128, 512, 1536), (128, 10000, 512), (10000, 512, 384), (10000, 512, 2048), (128, 512, 30000),
                                   (512, 512, 384), (19000, 256, 1024), (77, 40, 72)])
def test_split_k_and_tail_wave_gemm(M, N, K):
    """The training-step GEMM entry (K split over work items for skinny outputs; tail-wave tiles split over the idle SMs; TMA-store
    epilogue for the rest) against the deterministic entry: same products, fp32 summation order only."""
    g = torch.Generator().manual_seed(M + 3 * N + K)
    Ab, Bb = tc.cast_bf16(_rand(M, K, g)), tc.cast_bf16(_rand(N, K, g))
    ref = 0.25 * (Ab.double() @ Bb.double().t())
    for _ in range(2):                                   # twice into the same buffer: the entry zeroes what it accumulates into
        C = tc.gemm_store(Ab, Bb, alpha=0.25, split_k=True, out=torch.full((M, N), 7.0, device="cuda"))
        assert float((C.double() - ref).abs().max()) < 2e-4 * float(ref.abs().max())
    C0 = tc.gemm_store(Ab, Bb, alpha=0.25)
    assert float((C0.double() - ref).abs().max()) < 1e-4 * float(ref.abs().max())


@pytest.mark.parametrize("R,D", [(7, 64), (300, 512), (1025, 128)])
@pytest.mark.parametrize("mode", [tc.PLAIN, tc.SPLIT_A, tc.SPLIT_B])
def test_fused_rownorm_cast(R, D, mode):
    """F.normalize fused with the operand cast == row-normalise kernel followed by the cast kernel, bit for bit."""
    from clip_gp_b200 import _lib, ops
    g = torch.Generator().manual_seed(R + D + mode)
    x = _rand(R, D, g, 3.0)
    ref = tc.cast_bf16(ops.row_normalize(x), mode)
    seg = 3 if mode else 1
    out = torch.empty(R, seg * D, dtype=torch.bfloat16, device="cuda")
    inv = torch.empty(R, dtype=torch.float32, device="cuda")
    _lib.check(_lib.load().clipgp_rownorm_cast(x.data_ptr(), R, D, None, inv.data_ptr(), out.data_ptr(), seg * D, D, mode,
                                               _lib.stream_ptr(x.device)), "rownorm_cast")
    # the fused kernel sums the squares four at a time, so the norm (and a few last bits of the unit rows) may differ by an ulp
    assert float((out.float() - ref.float()).abs().max()) <= 2.0 ** -7 * float(ref.float().abs().max())
    assert torch.allclose(inv, 1.0 / x.norm(dim=-1), rtol=1e-6)
    hi = out[:, :D].float()
    assert float((hi - torch.nn.functional.normalize(x, dim=-1)).abs().max()) <= 2.0 ** -8


@pytest.mark.parametrize("a_t", [False, True])
@pytest.mark.parametrize("b_t", [False, True])
@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (200, 1000, 512), (1000, 512, 1280), (128, 10000, 512), (77, 300, 100), (4096, 512, 128)])
def test_tf32_gemm_all_operand_layouts(M, N, K, a_t, b_t):
    """kind::tf32 on fp32 operands read in place, K-major and MN-major (transposed-in-place) operands, ragged M / N / K:
    against a float64 product of the TF32-rounded operands (exact up to fp32 accumulation) and against the fp32 product within
    the TF32 tolerance."""
    if (a_t and M % 4) or (b_t and N % 4) or (not a_t and K % 4) or (not b_t and K % 4):
        pytest.skip("row pitch must be a multiple of 16 bytes")
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g); B = torch.randn(N, K, generator=g)
    Ad = (A.t().contiguous() if a_t else A).cuda(); Bd = (B.t().contiguous() if b_t else B).cuda()
    C = tc.gemm_tf32(Ad, Bd, 0.5, a_t=a_t, b_t=b_t)
    ref = 0.5 * (A.double() @ B.double().t())
    scale = float(ref.abs().max())
    assert float((C.cpu().double() - ref).abs().max()) < 2e-3 * scale          # TF32: 10-bit mantissa products
    # the hardware truncates / rounds fp32 to TF32 (top 19 bits): both candidates bound the result far more tightly
    def tf32(x, rnd):
        xi = x.contiguous().view(torch.int32)
        xi = ((xi + (0x1000 if rnd else 0)) & ~0x1FFF)
        return xi.view(torch.float32)
    errs = []
    for rnd in (False, True):
        r2 = 0.5 * (tf32(A, rnd).double() @ tf32(B, rnd).double().t())
        errs.append(float((C.cpu().double() - r2).abs().max()) / scale)
    assert min(errs) < 2e-5, errs


def test_descriptor_cache_keys_on_geometry_and_follows_the_data():
    """The TMA descriptor cache of gemm_tc.cu is keyed on (pointer, geometry, layout): the same storage viewed with another shape
    gets its own descriptor, and a cached descriptor reads whatever the buffer holds now."""
    g = torch.Generator().manual_seed(5)
    buf_a = torch.empty(256 * 512, device="cuda"); buf_b = torch.empty(384 * 512, device="cuda")
    out = torch.empty(256 * 768, device="cuda")
    for rnd in range(3):
        for (M, N, K) in ((256, 384, 512), (128, 192, 1024), (64, 96, 2048), (256, 384, 512)):
            A = buf_a[: M * K].view(M, K); B = buf_b[: N * K].view(N, K)
            A.copy_(_rand(M, K, g).cuda() if not A.is_cuda else _rand(M, K, g)); B.copy_(_rand(N, K, g))
            C = tc.gemm_tf32(A, B, 0.5, out=out[: M * N].view(M, N))
            ref = 0.5 * (A.double() @ B.double().t())
            assert float((C.double() - ref).abs().max()) < 2e-3 * float(ref.abs().max()), (rnd, M, N, K)
