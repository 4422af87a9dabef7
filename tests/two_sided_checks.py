"""Worker of tests/test_zz_contrastive_two_sided_gpu.py: the GPU checks of the opt-in "two_sided" contrastive backend and its grouped
kernels (b200mm_contrast_lse_partials_pair / b200mm_contrast_softgrad_pair), run in a SEPARATE process so that a defect in these not yet
hardware-verified kernels (written after round 2's GPU budget was spent) can neither poison the CUDA context of the verified suite nor hang
it: the parent kills this process on a timeout. Prints one `TWO_SIDED <check> OK|FAIL ...` line per check and `TWO_SIDED ALL OK`."""
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from oracle import restated

BF = torch.bfloat16


def _rel(a, b):
    return float((a.float().cpu() - b.float().cpu()).norm() / b.float().cpu().norm().clamp_min(1e-12))


def close(got, ref, tol, what=""):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = (got - ref).abs().max() / ref.abs().max().clamp_min(1e-6)
    assert torch.isfinite(got).all(), what
    assert float(err) <= tol, f"{what}: max err / scale = {float(err):.3e} > {tol}"


def check_clip(C, B, E):
    C.set_backend("two_sided")
    g = torch.Generator().manual_seed(B)
    a = F.normalize(torch.randn(B, E, generator=g), dim=-1).to(BF)
    b = F.normalize(torch.randn(B, E, generator=g), dim=-1).to(BF)
    ls = torch.tensor(2.3, device="cuda", requires_grad=True)
    a1, b1 = a.cuda().requires_grad_(), b.cuda().requires_grad_()
    loss = C.clip_contrastive_loss(a1, b1, ls)
    (loss * 1.7).backward()
    C.set_backend("gathered_grad")
    ls0 = torch.tensor(2.3, device="cuda", requires_grad=True)
    a0, b0 = a.cuda().requires_grad_(), b.cuda().requires_grad_()
    loss0 = C.clip_contrastive_loss(a0, b0, ls0)
    (loss0 * 1.7).backward()
    af, bf_, lsf = a.float().requires_grad_(), b.float().requires_grad_(), torch.tensor(2.3, requires_grad=True)
    ref = restated.symmetric_info_nce(lsf.exp() * af @ bf_.t())
    (ref * 1.7).backward()
    assert abs(float(loss) - float(ref)) < 1e-4 * max(1.0, abs(float(ref))), (float(loss), float(ref))
    assert abs(float(loss) - float(loss0)) < 1e-5 * max(1.0, abs(float(ref))), (float(loss), float(loss0))
    ea, eb = _rel(a1.grad, af.grad), _rel(b1.grad, bf_.grad)
    assert ea < 1e-2 and eb < 1e-2, (ea, eb)
    assert abs(float(ls.grad) - float(lsf.grad)) < 5e-3 * max(1.0, abs(float(lsf.grad))), (float(ls.grad), float(lsf.grad))
    return f"loss {float(loss):.6f} (oracle {float(ref):.6f}) dimg {ea:.2e} dtxt {eb:.2e} (default backend {_rel(a0.grad, af.grad):.2e})"


def check_mil(C, B, E):
    C.set_backend("two_sided")
    g = torch.Generator().manual_seed(100 + B)
    v = F.normalize(torch.randn(B, E, generator=g), dim=-1).to(BF)
    t = F.normalize(torch.randn(B, E, generator=g), dim=-1).to(BF)
    v1, t1 = v.cuda().requires_grad_(), t.cuda().requires_grad_()
    loss = C.mil_nce_loss(v1, t1)
    loss.backward()
    vf, tf = v.float().requires_grad_(), t.float().requires_grad_()
    ref = restated.mil_nce_n1(restated.l1_simi_matrix(tf, vf, 1).view(B, B))
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-4 * max(1.0, abs(float(ref))), (float(loss), float(ref))
    ev, et = _rel(v1.grad, vf.grad), _rel(t1.grad, tf.grad)
    assert ev < 1e-2 and et < 1e-2, (ev, et)
    return f"loss {float(loss):.6f} (oracle {float(ref):.6f}) dvideo {ev:.2e} dtext {et:.2e}"


def check_pair_kernels(ops, M, N, E, off, flags, dsub):
    """Grouped (two problems per launch) forms with device-resident scalars: lse_partials_pair equals the two single launches bit for bit;
    softgrad_pair equals the two-sided formula  alpha*coef*(wr e^{z - rl_m} + wc e^{z - cl_n})  (the -dsub diagonal term is the caller's,
    in fp32) and its dscale counts the full gradient including that term."""
    alpha, coef, gout = 14.3, 0.37, 1.7
    gen = torch.Generator().manual_seed(M + N)
    Npad = (N + 7) // 8 * 8
    a0 = F.normalize(torch.randn(M, E, generator=gen), dim=-1).to(BF).cuda()
    a1 = F.normalize(torch.randn(M, E, generator=gen), dim=-1).to(BF).cuda()
    b0 = torch.zeros(Npad, E, dtype=BF)
    b1 = torch.zeros(Npad, E, dtype=BF)
    b0[:N] = F.normalize(torch.randn(N, E, generator=gen), dim=-1).to(BF)
    b1[:N] = F.normalize(torch.randn(N, E, generator=gen), dim=-1).to(BF)
    b0, b1 = b0.cuda(), b1.cuda()
    alpha_dev = torch.full((1,), alpha, device="cuda")
    pA, pB = ops.contrast_lse_partials_pair(a0, b0[:N], a1, b1[:N], 1.0, off, alpha_dev=alpha_dev)
    sA = ops.contrast_lse_partials(a0, b0[:N], alpha, off)
    sB = ops.contrast_lse_partials(a1, b1[:N], alpha, off)
    for got, want in zip(pA + pB, sA + sB):
        assert torch.equal(got, want), "grouped launch differs from the single launches"
    z0 = alpha * a0.float() @ b0[:N].float().t()
    z1 = alpha * a1.float() @ b1[:N].float().t()
    rl0, rl1 = torch.logsumexp(z0, 1), torch.logsumexp(z1, 1)
    cl0 = torch.randn(N, device="cuda") + rl0.mean()   # "row LSEs of the other ranks": arbitrary per-column normalisers
    cl1 = torch.randn(N, device="cuda") + rl1.mean()
    dscale = torch.zeros(1, device="cuda")
    G0, G1 = ops.contrast_softgrad_pair(a0, b0, a1, b1, N, 1.0, off, rl0.contiguous(), cl0, rl1.contiguous(), cl1, coef, dsub, flags, dscale,
                                        alpha_dev=alpha_dev, coef_dev=torch.full((1,), gout, device="cuda"))
    rows, cols = torch.arange(M, device="cuda"), torch.arange(M, device="cuda") + off
    for k, (G, z, rl, cl) in enumerate(((G0, z0, rl0, cl0), (G1, z1, rl1, cl1))):
        er, ec = torch.exp(z - rl[:, None]), torch.exp(z - cl[None, :])
        if flags[2 * k]:
            er[rows, cols] = 0
        if flags[2 * k + 1]:
            ec[rows, cols] = 0
        g = coef * gout * (er + ec)
        ref = torch.zeros(M, Npad, device="cuda")
        ref[:, :N] = alpha * g
        close(G, ref, 8e-3, f"two-sided G{k}")
        assert float(G[:, N:].float().abs().max()) == 0.0 if Npad > N else True
        if k == 0:
            full = g.clone()
            full[rows, cols] -= coef * gout * dsub
            close(dscale, (full * z).sum().view(1), 2e-3, "dscale (problem 0 only, incl. the diagonal term)")
    return "grouped launch == single launches (bit-exact); two-sided tiles and dscale within tolerance"


def main():
    import b200mm
    import b200mm.contrastive as C

    ops = b200mm.ops
    checks = [(f"pair_kernels[{c}]", lambda c=c: check_pair_kernels(ops, *c)) for c in
              [(6, 6, 32, 0, (0, 0, 0, 0), 2.0), (200, 1000, 64, 300, (0, 1, 1, 0), 1.0), (128, 520, 768, 128, (0, 0, 0, 0), 2.0),
               (300, 2048, 96, 1024, (0, 0, 0, 0), 2.0)]]
    checks += [(f"clip[B={B},E={E}]", lambda B=B, E=E: check_clip(C, B, E)) for B, E in [(6, 32), (37, 64), (300, 768), (1024, 768)]]
    checks += [(f"mil[B={B},E={E}]", lambda B=B, E=E: check_mil(C, B, E)) for B, E in [(4, 32), (37, 64), (256, 512)]]
    bad = 0
    for name, fn in checks:
        try:
            msg = fn()
            torch.cuda.synchronize()
            print(f"TWO_SIDED {name} OK {msg}", flush=True)
        except Exception as e:  # noqa: BLE001
            bad += 1
            print(f"TWO_SIDED {name} FAIL {type(e).__name__}: {str(e)[:300]}", flush=True)
            traceback.print_exc()
            if "CUDA" in str(e) or "launch" in str(e):
                break  # a trapped kernel poisons the context
    if bad == 0:
        print("TWO_SIDED ALL OK", flush=True)
    sys.exit(0 if bad == 0 else 1)


if __name__ == "__main__":
    main()
