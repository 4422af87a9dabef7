"""world_size-2 gloo tests (CPU) of the N > 1 host logic: gather_tensor (row order, differentiable), the padded
all-gather / reduce-scatter helpers of the contrastive path, and the sharded-loss algebra (each rank owns its rows of
both logit blocks; W * local share; reduce-scatter of remote-row gradients) against the full-batch oracle."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent(
    """
    import os, sys
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    import b200mm
    from b200mm.distributed import gather_tensor, get_rank, get_world_size
    from b200mm.contrastive import _gather_rows, _scatter_grad
    from oracle import restated
    assert get_rank() == rank and get_world_size() == world

    # --- gather_tensor: cat order = rank order; backward = sum over ranks of the slice gradient
    torch.manual_seed(100 + rank)
    x = torch.randn(3, 5, requires_grad=True)
    g = gather_tensor(x, method="cat", back_gradient=True)
    assert g.shape == (3 * world, 5)
    assert torch.equal(g[3 * rank: 3 * rank + 3], x.detach())
    wgt = torch.arange(1, 3 * world + 1, dtype=torch.float32)[:, None] * (rank + 1)
    (g * wgt).sum().backward()
    expect = sum(torch.arange(1, 3 * world + 1, dtype=torch.float32)[3 * rank: 3 * rank + 3, None] * (r + 1) for r in range(world))
    assert torch.allclose(x.grad, expect.expand(3, 5)), (x.grad, expect)
    s = gather_tensor(x.detach(), method="stack")
    assert s.shape == (world, 3, 5) and not s.requires_grad
    sc = gather_tensor(torch.tensor(float(rank)))
    assert sc.tolist() == [float(r) for r in range(world)]

    # --- pad_tensors=True with UNEQUAL row counts (antmmf/utils/distributed_utils.py:145-176): rank r holds 2 + r rows
    from b200mm.distributed import gathered_sizes
    n_loc = 2 + rank
    y = (torch.arange(n_loc * 3, dtype=torch.float32).view(n_loc, 3) + 100 * rank).requires_grad_()
    sizes = gathered_sizes(y)
    assert sizes == [2 + r for r in range(world)]
    gy = gather_tensor(y, method="cat", back_gradient=True, pad_tensors=True)
    expect_rows = torch.cat([torch.arange((2 + r) * 3, dtype=torch.float32).view(2 + r, 3) + 100 * r for r in range(world)])
    assert gy.shape == expect_rows.shape and torch.equal(gy.detach(), expect_rows), gy
    wy = (torch.arange(1, gy.shape[0] + 1, dtype=torch.float32)[:, None] * (rank + 1)).expand_as(gy)
    (gy * wy).sum().backward()
    o0 = sum(sizes[:rank])
    exp_g = sum(torch.arange(1, gy.shape[0] + 1, dtype=torch.float32)[o0:o0 + n_loc, None] * (r + 1) for r in range(world))
    assert torch.allclose(y.grad, exp_g.expand(n_loc, 3)), (y.grad, exp_g)
    ng = gather_tensor(y.detach(), method="cat", pad_tensors=True)
    assert not ng.requires_grad and torch.equal(ng, expect_rows)
    try:
        gather_tensor(y.detach(), method="stack", pad_tensors=True)
        raise AssertionError("stack of unequal rows must fail like the reference's torch.stack")
    except RuntimeError:
        pass
    # equal sizes with pad_tensors=True take the plain path
    z = torch.full((2, 2), float(rank))
    assert torch.equal(gather_tensor(z, method="cat", pad_tensors=True), torch.cat([torch.full((2, 2), float(r)) for r in range(world)]))

    # --- padded gather / reduce-scatter helpers
    a = torch.full((3, 4), float(rank + 1))
    all_a, Bg = _gather_rows(a, None)
    assert Bg == 3 * world and all_a.shape[0] %% 8 == 0 and torch.equal(all_a[:Bg], torch.cat([torch.full((3, 4), float(r + 1)) for r in range(world)]))
    assert float(all_a[Bg:].abs().sum()) == 0.0
    back = _scatter_grad(all_a * (rank + 1), 3, None)
    assert torch.allclose(back, torch.full((3, 4), float(rank + 1) * sum(r + 1 for r in range(world))))

    # --- sharded contrastive algebra vs the full-batch oracle (fp64 emulation of what the kernels compute per rank)
    torch.manual_seed(7)
    B, E, alpha = 4, 6, 9.0
    I_all = torch.nn.functional.normalize(torch.randn(B * world, E, dtype=torch.float64), dim=-1)
    T_all = torch.nn.functional.normalize(torch.randn(B * world, E, dtype=torch.float64), dim=-1)
    I_loc = I_all[rank * B:(rank + 1) * B].clone().requires_grad_()
    T_loc = T_all[rank * B:(rank + 1) * B].clone().requires_grad_()
    Ig = gather_tensor(I_loc, method="cat", back_gradient=True)
    Tg = gather_tensor(T_loc, method="cat", back_gradient=True)
    A = alpha * I_loc @ Tg.t()
    Bt = alpha * T_loc @ Ig.t()
    idx = torch.arange(B) + rank * B
    local = ((torch.logsumexp(A, 1) - A[torch.arange(B), idx]).sum() + (torch.logsumexp(Bt, 1) - Bt[torch.arange(B), idx]).sum()) / (2 * B * world)
    (world * local).backward()
    # oracle: full loss, gradient w.r.t. this rank's rows; DDP would average the per-rank parameter gradients
    If, Tf = I_all.clone().requires_grad_(), T_all.clone().requires_grad_()
    full = restated.symmetric_info_nce(alpha * If @ Tf.t())
    full.backward()
    tot = torch.tensor([float(world * local)], dtype=torch.float64)
    dist.all_reduce(tot)
    assert abs(float(tot) / world - float(full)) < 1e-10
    assert torch.allclose(I_loc.grad / world, If.grad[rank * B:(rank + 1) * B], atol=1e-12)
    assert torch.allclose(T_loc.grad / world, Tf.grad[rank * B:(rank + 1) * B], atol=1e-12)
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(os.path.dirname(os.path.abspath(__file__)), f"ok_{rank}"), "w").write("ok")
    """
)


def test_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    port = 29500 + (os.getpid() % 400)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert (tmp_path / "ok_0").exists() and (tmp_path / "ok_1").exists()


WORKER_STAGE2 = textwrap.dedent(
    """
    import os, sys
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    import b200mm
    from oracle import restated
    from tests import emulated_ops
    from tests.test_cross_host_cpu import build_scorer, rel_l2
    BF = torch.bfloat16
    fx = torch.load(os.path.join(%r, "tests", "golden", "stage2.pt"), weights_only=False)
    heads = fx["config"]["heads"]
    sc = build_scorer(fx).to(BF).train()
    # identical global data on every rank; each rank owns 3 texts and 3 videos
    g = torch.Generator().manual_seed(5)
    b, St, Sv, H = 3, 6, 3, 64
    seq_all = torch.randn(b * world, St, H, generator=g).to(BF).float()
    vis_all = torch.randn(b * world, Sv, H, generator=g).to(BF).float()
    am_all = torch.ones(b * world, St, dtype=torch.long); am_all[1, 4:] = 0; am_all[4, 3:] = 0
    vm_all = torch.ones(b * world, Sv, dtype=torch.long); vm_all[5, 2:] = 0
    l1 = torch.randn(b * world, b * world, generator=g) + 2.0 * torch.eye(b * world)
    R = torch.randn(world, b, b, generator=g)
    sl = slice(rank * b, (rank + 1) * b)
    seq = seq_all[sl].clone().requires_grad_()
    vis = vis_all[sl].clone().requires_grad_()
    with emulated_ops.patched():
        l2 = sc.cross_similarity_hard_mining((vis, vm_all[sl], None, 1, None), (seq, am_all[sl], None, b, None), l1.clone(), "top_k")
        (l2 * R[rank]).sum().backward()
    # oracle: BOTH ranks' computations in one process on the concatenated data
    sd = {k: v.to(BF).float().requires_grad_(True) for k, v in fx["state_dict"].items()}
    o_seq, o_vis = seq_all.clone().requires_grad_(), vis_all.clone().requires_grad_()
    total, mine = 0.0, None
    for r in range(world):
        chosen = restated.hard_mining_indices(l1, r * b, b, "top_k")
        assert chosen.min() >= 0 and chosen.max() < b * world and torch.equal(torch.diagonal(chosen), torch.arange(b) + r * b)
        o_l2 = restated.cross_similarity_hard_mining(sd, o_seq[r * b:(r + 1) * b], am_all[r * b:(r + 1) * b], o_vis, vm_all, chosen, heads)
        total = total + (o_l2 * R[r]).sum()
        if r == rank:
            mine = o_l2.detach()
    total.backward()
    assert rel_l2(l2, mine) < 2e-2, rel_l2(l2, mine)
    # local text gradient = this rank's term only; local video gradient = the sum over ALL ranks' terms (reduce-scatter of gather_tensor)
    assert rel_l2(seq.grad, o_seq.grad[sl]) < 8e-2, rel_l2(seq.grad, o_seq.grad[sl])
    assert rel_l2(vis.grad, o_vis.grad[sl]) < 8e-2, rel_l2(vis.grad, o_vis.grad[sl])
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(os.path.dirname(os.path.abspath(__file__)), f"ok2_{rank}"), "w").write("ok")
    """
)


def test_world2_gloo_stage2_hard_mining(tmp_path):
    """Stage-2 hard-negative mining across ranks: local texts scored against videos GATHERED from every rank (global indices from the
    level-1 matrix, beg_idx = rank * bsz), video-token gradients returned home by the reduce-scatter — vs the single-process oracle."""
    script = tmp_path / "worker_stage2.py"
    script.write_text(WORKER_STAGE2 % (ROOT, ROOT))
    port = 29900 + (os.getpid() % 90)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert (tmp_path / "ok2_0").exists() and (tmp_path / "ok2_1").exists()


WORKER_LOSSES = textwrap.dedent(
    """
    import os, sys
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    import b200mm
    from b200mm.contrastive import clip_contrastive_loss, mil_nce_loss
    from oracle import restated
    from tests import emulated_ops
    BF = torch.bfloat16

    def rel(a, b):
        a, b = a.detach().float(), b.detach().float()
        return float((a - b).norm() / b.norm().clamp_min(1e-12))

    def mean_over_ranks(x):
        t = torch.tensor([float(x)], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t) / world

    g = torch.Generator().manual_seed(11)
    B, E = 5, 16   # 5 rows per rank: the gathered 10 rows are padded to 16 inside _gather_rows
    I_all = torch.nn.functional.normalize(torch.randn(B * world, E, generator=g), dim=-1).to(BF)
    T_all = torch.nn.functional.normalize(I_all.float() + 0.7 * torch.randn(B * world, E, generator=g), dim=-1).to(BF)
    sl = slice(rank * B, (rank + 1) * B)
    with emulated_ops.patched():
        # ---- symmetric InfoNCE (CNCLIP.contrastive_loss): the REAL _ContrastiveFn host code, sharded over 2 ranks
        i_loc, t_loc = I_all[sl].clone().requires_grad_(), T_all[sl].clone().requires_grad_()
        ls = torch.tensor(2.0, requires_grad=True)
        loss = clip_contrastive_loss(i_loc, t_loc, ls)
        loss.backward()
        If, Tf, lsf = I_all.float().requires_grad_(), T_all.float().requires_grad_(), torch.tensor(2.0, requires_grad=True)
        full = restated.symmetric_info_nce(lsf.exp() * If @ Tf.t())
        full.backward()
        assert abs(mean_over_ranks(loss) - float(full)) < 1e-4 * float(full), (float(loss), float(full))
        # DDP averages parameter gradients over ranks: (1/W) * sum_r d(W * share_r) = d(global loss)
        assert rel(i_loc.grad.float() / world, If.grad[sl]) < 2e-2 and rel(t_loc.grad.float() / world, Tf.grad[sl]) < 2e-2
        assert abs(mean_over_ranks(ls.grad) - float(lsf.grad)) < 2e-2 * max(1.0, abs(float(lsf.grad)))
        # ---- MIL-NCE, n_clips = 1 and 2 (forward_stage1 of base_vtp)
        for n in (1, 2):
            V_all = torch.nn.functional.normalize(T_all.float().repeat_interleave(n, 0) + 0.6 * torch.randn(B * world * n, E, generator=g), dim=-1).to(BF)
            v_loc = V_all[rank * B * n:(rank + 1) * B * n].clone().requires_grad_()
            t_loc = T_all[sl].clone().requires_grad_()
            loss = mil_nce_loss(v_loc, t_loc, None, n_clips=n)
            loss.backward()
            Vf, Tf = V_all.float().requires_grad_(), T_all.float().requires_grad_()
            full = restated.mil_nce_clips(restated.l1_simi_matrix(Tf, Vf, n))
            full.backward()
            assert abs(mean_over_ranks(loss) - float(full)) < 2e-4 * float(full), (n, float(loss), float(full))
            assert rel(v_loc.grad.float() / world, Vf.grad[rank * B * n:(rank + 1) * B * n]) < 3e-2, n
            assert rel(t_loc.grad.float() / world, Tf.grad[sl]) < 3e-2, n
    # opt-in guard: ranks with different numbers of pairs raise (on every rank) instead of corrupting / hanging the all-gather
    os.environ["B200MM_CHECK_SHAPES"] = "1"
    try:
        with emulated_ops.patched():
            n_bad = B + rank
            clip_contrastive_loss(torch.randn(n_bad, E).to(BF), torch.randn(n_bad, E).to(BF), torch.tensor(2.0))
        raise AssertionError("unequal per-rank batches must raise under B200MM_CHECK_SHAPES=1")
    except ValueError as e:
        assert "different numbers of pairs" in str(e), e
    finally:
        os.environ["B200MM_CHECK_SHAPES"] = "0"
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(os.path.dirname(os.path.abspath(__file__)), f"ok3_{rank}"), "w").write("ok")
    """
)


import pytest  # noqa: E402


@pytest.mark.parametrize("backend", ["gathered_grad", "two_sided"])
def test_world2_gloo_contrastive_losses_host_code(tmp_path, backend):
    """The production host code of the sharded losses (b200mm.contrastive._ContrastiveFn / _MilNceClipsFn: padded all-gather, per-rank
    logit blocks, W x local share, reduce-scatter of remote-row gradients) on 2 gloo ranks over the torch stand-ins of the kernels — vs
    the full-batch oracle: mean over ranks of the loss = global loss, DDP-averaged gradients = global gradient."""
    script = tmp_path / "worker_losses.py"
    script.write_text(WORKER_LOSSES % ROOT)
    port = 29800 + (os.getpid() % 90)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, B200MM_CONTRASTIVE=backend))  # "two_sided": LSE all-gather instead of the gradient reduce-scatter
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert (tmp_path / "ok3_0").exists() and (tmp_path / "ok3_1").exists()
