// Shared-memory carve-up of the per-class GP kernels.  Matrices that consecutive threads walk down a
// column of use an odd leading dimension (ldn = n|1, ldt = T|1) so the walk is bank-conflict free.
#pragma once
#include "gp_common.cuh"

namespace clipgp {
namespace gp {

constexpr int SCH = 32;  // MC samples processed per chunk

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }
__host__ __device__ inline size_t maxsz(size_t a, size_t b) { return a > b ? a : b; }

struct Dims {
    int T, n, d, ldn, ldt;
    size_t f_nn, f_nt, f_tt;  // element counts of [n][ldn], [n][ldt], [T][ldt]
    size_t tiles;             // floats of the two streamed chunk tiles
};

__host__ __device__ inline Dims make_dims(int T, int n, int d) {
    Dims D;
    D.T = T; D.n = n; D.d = d; D.ldn = n | 1; D.ldt = T | 1;
    D.f_nn = (size_t)n * D.ldn; D.f_nt = (size_t)n * D.ldt; D.f_tt = (size_t)T * D.ldt;
    D.tiles = (size_t)(((n + 3) & ~3) + ((T + 3) & ~3)) * KCP;
    return D;
}

// ---------------------------------------------------------------- forward
struct FwdLayout {
    size_t Ld, Ad, invd;                                   // double
    size_t Sig, R, AfBm, Lq, mu, mvec, invls, invdR, pool; // float
    size_t line;                                           // exchange lines of the whole-CTA factorisations (gp_block.cuh)
    size_t total;
};

constexpr size_t kLineBytes = (36 + 8 * 68 + 4 * 64) * 8;     // Blk4<65, 64>::kCholScratch doubles (>= the 2 * 4 * 68 of the back substitution)

__host__ __device__ inline FwdLayout make_fwd_layout(const Dims& D) {
    FwdLayout L;
    size_t o = 0;
    L.Ld = o;    o = align16(o + 8 * D.f_nn);
    L.Ad = o;    o = align16(o + 8 * D.f_nt);
    L.invd = o;  o = align16(o + 8 * (size_t)D.n);
    L.Sig = o;   o = align16(o + 4 * D.f_tt);
    L.R = o;     o = align16(o + 4 * D.f_tt);
    L.AfBm = o;  o = align16(o + 4 * maxsz(2 * D.f_nt, D.f_nn));   // Af | Bm, aliased by the scratch Gram K0 [n][ldn]
    L.Lq = o;    o = align16(o + 4 * D.f_nn);
    L.mu = o;    o = align16(o + 4 * (size_t)D.T);
    L.mvec = o;  o = align16(o + 4 * (size_t)D.n);
    L.invls = o; o = align16(o + 4 * (size_t)(D.d > 0 ? D.d : 1));
    L.invdR = o; o = align16(o + 4 * (size_t)D.T);
    // pool: chunk tiles during the Gram phase, then the sample chunk buffers f [SCH][ldt] and eps [T][SCH]
    L.pool = o;  o = align16(o + 4 * maxsz(D.tiles, (size_t)SCH * D.ldt + (size_t)D.T * SCH));
    L.line = o;  o = align16(o + kLineBytes);
    L.total = o;
    return L;
}

// ---------------------------------------------------------------- backward
// Persistent: Ld, invd, Af, dSig (becomes dK_XX), dAd, small vectors.  `pool` is re-carved per phase:
//   B1 (sparsemax + sampling + chol32 adjoint): R | scrF | df [SCH][ldt] | eps [T][SCH]
//   B3 (predictive adjoint)                  : Lq | Bm | dBm | dAf
//   B4 (fp64 solve + chol64 adjoint)         : scrD (double) | dLd (double) | dKzz (float, past both)
//   B5 (kernel adjoint)                      : dKzx | raw | tiles        (dKzz stays where B4 left it)
struct BwdLayout {
    size_t Ld, invd, dAd;                                   // double
    size_t Af, dSig, mvec, dmu, invls, invdR, dls, dzl, rs, cs, line, pool;
    size_t p_R, p_scrF, p_df, p_eps;                        // B1 (offsets from start of smem)
    size_t p_Lq, p_Bm, p_dBm, p_dAf;                        // B3
    size_t p_scrD, p_dLd, p_dKzz;                           // B4
    size_t p_dKzx, p_raw, p_tiles;                          // B5
    size_t total;
};

__host__ __device__ inline BwdLayout make_bwd_layout(const Dims& D) {
    BwdLayout L;
    size_t o = 0;
    L.Ld = o;    o = align16(o + 8 * D.f_nn);
    L.invd = o;  o = align16(o + 8 * (size_t)D.n);
    L.dAd = o;   o = align16(o + 8 * D.f_nt);
    L.Af = o;    o = align16(o + 4 * D.f_nt);
    L.dSig = o;  o = align16(o + 4 * D.f_tt);
    L.mvec = o;  o = align16(o + 4 * (size_t)D.n);
    L.dmu = o;   o = align16(o + 4 * (size_t)D.T);
    L.invls = o; o = align16(o + 4 * (size_t)(D.d > 0 ? D.d : 1));
    L.invdR = o; o = align16(o + 4 * (size_t)D.T);
    L.dls = o;   o = align16(o + 4 * (size_t)(D.d > 0 ? D.d : 1));
    L.dzl = o;   o = align16(o + 4 * (size_t)(D.d > 0 ? D.d : 1));
    L.rs = o;    o = align16(o + 4 * (size_t)(D.n + 3));
    L.cs = o;    o = align16(o + 4 * (size_t)(D.n + 3));
    L.line = o;  o = align16(o + kLineBytes);
    L.pool = o;
    size_t p = o;
    L.p_R = p;     p = align16(p + 4 * D.f_tt);
    L.p_scrF = p;  p = align16(p + 4 * D.f_tt);
    L.p_df = p;    p = align16(p + 4 * (size_t)SCH * D.ldt);
    L.p_eps = p;   p = align16(p + 4 * (size_t)D.T * SCH);
    size_t end = p;
    p = o;
    L.p_Lq = p;    p = align16(p + 4 * D.f_nn);
    L.p_Bm = p;    p = align16(p + 4 * D.f_nt);
    L.p_dBm = p;   p = align16(p + 4 * D.f_nt);
    L.p_dAf = p;   p = align16(p + 4 * D.f_nt);
    end = maxsz(end, p);
    p = o;
    L.p_scrD = p;  p = align16(p + 8 * D.f_nn);
    L.p_dLd = p;   p = align16(p + 8 * D.f_nn);
    L.p_dKzz = p;  p = align16(p + 4 * D.f_nn);
    end = maxsz(end, p);
    p = o;
    L.p_dKzx = p;  p = align16(p + 4 * D.f_nt);
    L.p_raw = p;   p = align16(p + 4 * maxsz(D.f_nn, D.f_nt));
    L.p_tiles = p; p = align16(p + 4 * D.tiles);
    // the B5 buffers must not reach dKzz (left in place by B4)
    if (p > L.p_dKzz) {
        // push dKzz (and the pool end) out of the way
        L.p_dKzz = p;
        end = maxsz(end, align16(p + 4 * D.f_nn));
    }
    end = maxsz(end, p);
    L.total = end;
    return L;
}

}  // namespace gp
}  // namespace clipgp
