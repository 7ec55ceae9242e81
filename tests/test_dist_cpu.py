"""CPU, world_size 2 over gloo: the sharding plumbing of clip_gp_b200/dist.py with the ORACLE as the per-rank compute
(the CUDA kernels need a GPU; what is tested here is partitioning + reduction logic)."""
import os

import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from clip_gp_b200 import dist as cd


def test_partitions_cover_everything():
    for n in (0, 1, 7, 50000, 50001):
        for w in (1, 2, 3, 8):
            rs = [cd.shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
    assert [cd.sample_split(10, r, 8)[1] for r in range(8)] == [2, 2, 1, 1, 1, 1, 1, 1]
    for S in (1, 10, 16, 100):
        for w in (1, 2, 4, 8):
            if S < w:
                continue
            sp = [cd.sample_split(S, r, w) for r in range(w)]
            assert sp[0][0] == 0 and sum(c for _, c in sp) == S and all(a[0] + a[1] == b[0] for a, b in zip(sp, sp[1:]))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from clip_gp_b200 import synth
        from oracle import gp as ogp, heads as oh, metrics as om, philox
        from oracle.train_step import OracleAdapter
        torch.manual_seed(0)
        # ---------------- eval: shard the images, all-reduce the integer counters, gather (conf, hit) for AECE
        g = torch.Generator().manual_seed(5)
        N, C = 1001, 23
        logits = 4.0 * torch.randn(N, C, generator=g)
        labels = torch.randint(0, C, (N,), generator=g)
        lo, hi = cd.shard_range(N, rank, world)
        conf, pred, correct = om.confidence(logits[lo:hi], labels[lo:hi])
        b = torch.linspace(0, 1, 11)
        hist = torch.zeros(4, 10, dtype=torch.int64)
        for i in range(10):
            inb = (conf > b[i]) & (conf <= b[i + 1])
            hist[0, i] = inb.sum(); hist[2, i] = correct[inb].sum()
            hist[1, i] = (conf[inb].double() * float(1 << 40)).to(torch.int64).sum()
        hist[3, 0] = correct.sum()
        h, conf_g, cor_g = cd.global_calibration(hist, conf, correct.to(torch.uint8), N, world)
        e_ref, bins_ref = om.compute_ece_with_bins(logits, labels)
        assert h[0].tolist() == bins_ref["bin_count"] and int(h[3, 0]) == om.top1_count(logits, labels)
        ece = sum(abs(int(h[1, i]) / float(1 << 40) / int(h[0, i]) - int(h[2, i]) / int(h[0, i])) * int(h[0, i]) / N
                  for i in range(10) if int(h[0, i]) > 0) * 100
        assert abs(ece - e_ref) < 1e-4
        a_ref, _, _, _ = om.aece_bins(*om.confidence(logits, labels)[::2])
        a_got, _, _, _ = om.aece_bins(conf_g, cor_g.bool())
        assert abs(a_got - a_ref) < 1e-6
        # ---------------- training: shard the MC samples, all-reduce the flat gradient; equals the single-process gradient
        wl = synth.make_workload("tiny"); shp = wl["shape"]
        S = 5
        f, y = wl["f_train"][: shp.B], wl["y_train"][: shp.B]

        def grads(s_off, s_cnt, kl_share):
            st = ogp.build_state(wl["E"], "rbf", shp.d)
            st.var_mean, st.chol_var = synth.trained_like_q(shp.C, shp.T + 1, 5)
            orc = OracleAdapter(st, shp.D, shots=shp.shots)
            eps = philox.eps_tensor(7, 0, shp.C, shp.T, s_cnt, s_offset=s_off, S_total=S)
            protos, _ = ogp.sample_prototypes(st, eps)
            ce = oh.adapter_mc_ce(f, y, orc.W, protos, 100.0) * (s_cnt / S)      # this rank's share of mean_s CE_s
            kl = ogp.kl_divergence(st.var_mean, st.chol_var).sum() * 0.01 * kl_share
            eye = torch.eye(shp.D)
            l2 = (orc.W - eye).pow(2).sum() * (0.5 / shp.shots) * kl_share
            (ce + kl + l2).backward()
            return torch.cat([orc.W.grad.reshape(-1), st.var_mean.grad.reshape(-1), st.chol_var.grad.reshape(-1)])

        off, cnt = cd.sample_split(S, rank, world)
        local = grads(off, cnt, 1.0 / world)
        cd.allreduce_sum_(local)
        full = grads(0, S, 1.0)
        assert float((local - full).abs().max()) < 1e-5 * float(full.abs().max()) + 1e-7
        # ---------------- hybrid sharding of the engine: classes sharded for the GP, MC samples for the logit path.
        # Each rank computes w for its classes (all samples), an all-reduce completes w; the CE share of its samples gives dw rows,
        # a second all-reduce completes dw; the GP adjoint then runs on the rank's classes only; gradients are summed.
        import copy
        import dataclasses
        c_lo, c_hi = cd.shard_range(shp.C, rank, world)

        def class_slice(st):
            kw = {}
            for fld in dataclasses.fields(st):
                v = getattr(st, fld.name)
                if fld.name == "kernel":
                    kp = copy.copy(v)
                    for nm in ("raw_lengthscale", "raw_outputscale", "raw_variance"):
                        t = getattr(kp, nm)
                        if t is not None:
                            setattr(kp, nm, t[c_lo:c_hi].clone().requires_grad_(True))
                    kw[fld.name] = kp
                elif torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == shp.C and fld.name not in ("pca_W",):
                    kw[fld.name] = v[c_lo:c_hi].clone()
                else:
                    kw[fld.name] = v
            return type(st)(**kw)

        st = ogp.build_state(wl["E"], "rbf", shp.d)
        st.var_mean, st.chol_var = synth.trained_like_q(shp.C, shp.T + 1, 5)
        sl = class_slice(st)
        sl.var_mean.requires_grad_(True); sl.chol_var.requires_grad_(True)
        eps_all = philox.eps_tensor(7, 0, shp.C, shp.T, S)                    # class / sample sharding never changes the draws
        w_own, _ = ogp.gp_weights(sl, eps_all[c_lo:c_hi])                     # [S, C_local, T], all samples
        w_full = torch.zeros(S, shp.C, shp.T)
        w_full[:, c_lo:c_hi] = w_own.detach()
        cd.allreduce_sum_(w_full)
        w_leaf = w_full[off:off + cnt].clone().requires_grad_(True)           # this rank's samples, all classes
        Wp = torch.eye(shp.D, requires_grad=True)
        protos = torch.einsum("skm,kmd->skd", w_leaf, st.templates)
        ce = oh.adapter_mc_ce(f, y, Wp, protos, 100.0) * (cnt / S) if cnt > 0 else Wp.sum() * 0.0
        l2 = (Wp - torch.eye(shp.D)).pow(2).sum() * (0.5 / shp.shots) / world
        (ce + l2).backward()
        dw_full = torch.zeros(S, shp.C, shp.T)
        if cnt > 0:
            dw_full[off:off + cnt] = w_leaf.grad
        cd.allreduce_sum_(dw_full)
        kl_own = ogp.kl_divergence(sl.var_mean, sl.chol_var).sum() * 0.01       # full weight: the class shard owns its KL terms
        ((w_own * dw_full[:, c_lo:c_hi]).sum() + kl_own).backward()
        g_m = torch.zeros(shp.C, shp.T + 1); g_m[c_lo:c_hi] = sl.var_mean.grad
        g_L = torch.zeros(shp.C, shp.T + 1, shp.T + 1); g_L[c_lo:c_hi] = sl.chol_var.grad
        hybrid = torch.cat([Wp.grad.reshape(-1), g_m.reshape(-1), g_L.reshape(-1)])
        cd.allreduce_sum_(hybrid)
        assert float((hybrid - full).abs().max()) < 1e-5 * float(full.abs().max()) + 1e-7
        # ---------------- batch sharding (the default multi-GPU mode): every rank its OWN batch, every loss term pre-divided by
        # the world size, gradients summed -> the gradient of the mean loss over the world * B rows; then the fused peer step's protocol
        # (csrc/peer.cu): reduce-scatter, AdamW on this rank's quarter-float4 slice only, all-gather == AdamW on the whole vector
        B = shp.B

        def grads_batch(f_, y_, share):
            st = ogp.build_state(wl["E"], "rbf", shp.d)
            st.var_mean, st.chol_var = synth.trained_like_q(shp.C, shp.T + 1, 5)
            orc = OracleAdapter(st, shp.D, shots=shp.shots)
            protos, _ = ogp.sample_prototypes(st, philox.eps_tensor(7, 0, shp.C, shp.T, S))       # same draws on every rank
            ce = oh.adapter_mc_ce(f_, y_, orc.W, protos, 100.0) * share
            kl = ogp.kl_divergence(st.var_mean, st.chol_var).sum() * 0.01 * share
            l2 = (orc.W - torch.eye(shp.D)).pow(2).sum() * (0.5 / shp.shots) * share
            (ce + kl + l2).backward()
            params = torch.cat([orc.W.detach().reshape(-1), st.var_mean.detach().reshape(-1), st.chol_var.detach().reshape(-1)])
            return torch.cat([orc.W.grad.reshape(-1), st.var_mean.grad.reshape(-1), st.chol_var.grad.reshape(-1)]), params

        fa, ya = wl["f_train"][: world * B], wl["y_train"][: world * B]
        g_loc, p0 = grads_batch(fa[rank * B:(rank + 1) * B], ya[rank * B:(rank + 1) * B], 1.0 / world)
        g_sum = g_loc.clone()
        cd.allreduce_sum_(g_sum)
        g_full, _ = grads_batch(fa, ya, 1.0)
        assert float((g_sum - g_full).abs().max()) < 1e-5 * float(g_full.abs().max()) + 1e-7

        def adamw(p, g_, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, t=1):
            m = (1 - b1) * g_; v = (1 - b2) * g_ * g_
            return p - (lr / (1 - b1 ** t)) * m / (v.sqrt() / (1 - b2 ** t) ** 0.5 + eps)

        n = g_sum.numel(); n4 = n >> 2; per = (n4 + world - 1) // world          # the slice formula of peer.cu
        lo, hi = 4 * per * rank, (4 * min(per * (rank + 1), n4) if rank < world - 1 else n)
        mine = torch.zeros(n)
        mine[lo:hi] = adamw(p0[lo:hi], g_sum[lo:hi])                            # reduce-scatter + AdamW on the own slice
        cd.allreduce_sum_(mine)                                                 # all-gather (disjoint slices)
        assert torch.equal(mine, adamw(p0, g_sum))
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        td.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_eval_and_train_sharding_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
