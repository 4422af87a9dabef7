"""The contrastive-loss kernels at BASELINE.json configs[2] per-rank size (rank 0 of 8: B = 1024 local rows of each modality against the
B_g = 8192 gathered rows, E = 768), timed alone with CUDA events — SURVEY.md §8d "contrastive kernel roofline": both fractions are
reported (tensor: 16·B·B_g·E FLOP; HBM: algorithmic bytes = gathered I,T read + their gradients written + LSE vectors) next to the
reference's unfused arithmetic (materialised fp32 logits, log_softmax, autograd) run by torch on the same GPU.

  python tools/contrast_bench.py [E]        (ITERS=1 for ncu captures)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from b200mm import ops

BF = torch.bfloat16
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    pk = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))
    PEAK_TF, PEAK_TF_BURST, PEAK_HBM = pk["bf16_tflops_sustained"], pk["bf16_tflops"], pk["hbm_gbs"]
except Exception:
    PEAK_TF, PEAK_TF_BURST, PEAK_HBM = 1400.0, 1650.0, 6650.0
ITERS = int(os.environ.get("ITERS", "20"))


def timeit(fn):
    for _ in range(min(3, ITERS)):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(ITERS):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / ITERS


def timeit_graph(fn):
    """Device time of the launch sequence: captured once into a CUDA graph and replayed (no per-launch host cost)."""
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
        with torch.cuda.graph(g, stream=side):
            keep = fn()  # noqa: F841 (keeps the captured outputs alive)
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(ITERS):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / ITERS


def run(B, Bg, E, alpha=14.285):
    torch.manual_seed(0)
    a_all = torch.nn.functional.normalize(torch.randn(Bg, E, device="cuda"), dim=-1).to(BF)
    b_all = torch.nn.functional.normalize(torch.randn(Bg, E, device="cuda"), dim=-1).to(BF)
    a, b = a_all[:B].contiguous(), b_all[:B].contiguous()
    off, coef = 0, 1.0 / (2 * Bg)
    state = {}

    def fwd():
        pA = ops.contrast_lse_partials(a, b_all, alpha, off)
        pB = ops.contrast_lse_partials(b, a_all, alpha, off)
        loss = torch.zeros(1, device="cuda")
        state["lseA"] = ops.contrast_lse_merge(pA[:2], None, pA[2], False, loss)
        state["lseB"] = ops.contrast_lse_merge(pB[:2], None, pB[2], False, loss)
        return loss

    def bwd():
        ds = torch.zeros(1, device="cuda")
        GA = ops.contrast_softgrad(a, b_all, Bg, alpha, off, state["lseA"], coef, 1.0, False, ds)
        GB = ops.contrast_softgrad(b, a_all, Bg, alpha, off, state["lseB"], coef, 1.0, False, ds)
        da = ops.gemm(GA, b_all, b_mn=True, out_f32=True)
        db = ops.gemm(GB, a_all, b_mn=True, out_f32=True)
        db_all = ops.gemm(GA, a, a_mn=True, b_mn=True, out_f32=True)
        da_all = ops.gemm(GB, b, a_mn=True, b_mn=True, out_f32=True)
        return da, db, da_all, db_all

    # round 2: both directions per grouped launch, two-sided gradient tiles (no gathered-row gradient GEMMs, no reduce-scatter), device scalars
    alpha_dev = torch.full((1,), alpha, device="cuda")
    gout = torch.ones(1, device="cuda")

    def fwd2():
        pA, pB = ops.contrast_lse_partials_pair(a, b_all, b, a_all, 1.0, off, alpha_dev=alpha_dev)
        loss = torch.zeros(1, device="cuda")
        state["lseA"] = ops.contrast_lse_merge(pA[:2], None, pA[2], False, loss)
        state["lseB"] = ops.contrast_lse_merge(pB[:2], None, pB[2], False, loss)
        return loss

    def bwd2():
        ds = torch.zeros(1, device="cuda")
        GA, GB = ops.contrast_softgrad_pair(a, b_all, b, a_all, Bg, 1.0, off, state["lseA"], state["lseB_all"], state["lseB"], state["lseA_all"],
                                            1.0 / (2 * Bg), 2.0, (0, 0, 0, 0), ds, alpha_dev=alpha_dev, coef_dev=gout)
        dcoef = gout * (-2.0 / (2 * Bg)) * alpha_dev
        da = torch.addcmul(ops.gemm(GA, b_all, b_mn=True, out_f32=True), b.float(), dcoef)
        db = torch.addcmul(ops.gemm(GB, a_all, b_mn=True, out_f32=True), a.float(), dcoef)
        return da.to(BF), db.to(BF), ds

    fwd()
    # the row LSEs of the other ranks (what the forward all-gathers): here the LSE of every gathered row against this rank's columns is
    # not the real thing, but the kernel's work is identical — use full-width LSEs computed once with torch
    state["lseA_all"] = torch.logsumexp(alpha * a_all.float() @ b_all.float().t(), 1)
    state["lseB_all"] = torch.logsumexp(alpha * b_all.float() @ a_all.float().t(), 1)
    v1 = (timeit_graph(fwd), timeit_graph(bwd))
    fwd, bwd = fwd2, bwd2
    fwd()
    t_f_eager, t_b_eager = timeit(fwd), timeit(bwd)
    try:  # the headline figures: device time without host launch cost (in training these launches queue behind the encoder kernels)
        t_f, t_b = timeit_graph(fwd), timeit_graph(bwd)
        timing = "cuda-graph replay (device time)"
    except Exception as e:  # noqa: BLE001
        t_f, t_b, timing = t_f_eager, t_b_eager, f"eager launches (graph capture failed: {type(e).__name__})"
    fl_f, fl_b = 4.0 * B * Bg * E, 12.0 * B * Bg * E
    bytes_alg = 2 * Bg * E * 2 + 2 * Bg * E * 2 + 2 * Bg * 4  # read gathered I,T; write their gradients (bf16); LSE vectors
    # what this implementation additionally moves: the bf16 softmax-gradient blocks G [B, B_g] x 2, written once and read twice
    bytes_g = 2 * B * Bg * 2 * 2  # round 2: written once, read once

    # the reference's arithmetic, unfused, by torch on the same GPU: fp32 logits of this rank's rows, log_softmax, autograd
    af, bf_ = a.float().requires_grad_(), b.float().requires_grad_()
    aa, ba = a_all.float().requires_grad_(), b_all.float().requires_grad_()
    idx = torch.arange(B, device="cuda")

    def ref():
        for t in (af, bf_, aa, ba):
            t.grad = None
        A = alpha * af @ ba.t()
        Bt = alpha * bf_ @ aa.t()
        loss = (torch.nn.functional.cross_entropy(A, idx, reduction="sum") + torch.nn.functional.cross_entropy(Bt, idx, reduction="sum")) * coef
        loss.backward()
        return loss

    t_ref = timeit(ref)
    # same with bf16 tensor-core matmuls (what .cuda().bfloat16() modules would do)
    ab, bb, aab, bab = (t.detach().to(BF).requires_grad_() for t in (af, bf_, aa, ba))

    def ref16():
        for t in (ab, bb, aab, bab):
            t.grad = None
        A = (alpha * ab @ bab.t()).float()
        Bt = (alpha * bb @ aab.t()).float()
        loss = (torch.nn.functional.cross_entropy(A, idx, reduction="sum") + torch.nn.functional.cross_entropy(Bt, idx, reduction="sum")) * coef
        loss.backward()
        return loss

    t_ref16 = timeit(ref16)
    t = t_f + t_b
    out = {
        "shape": {"B_local": B, "B_global": Bg, "E": E},
        "round1_launch_sequence_ms": {"fwd": round(v1[0], 4), "bwd": round(v1[1], 4), "total": round(v1[0] + v1[1], 4)},
        "timing": timing, "fwd_ms": round(t_f, 4), "bwd_ms": round(t_b, 4), "total_ms": round(t, 4),
        "eager_launch_fwd_ms": round(t_f_eager, 4), "eager_launch_bwd_ms": round(t_b_eager, 4),
        "tflops": round((fl_f + fl_b) / t / 1e9, 1), "frac_of_sustained_bf16_peak": round((fl_f + fl_b) / t / 1e9 / PEAK_TF, 3),
        "fwd_tflops": round(fl_f / t_f / 1e9, 1), "bwd_tflops": round(fl_b / t_b / 1e9, 1),
        "algorithmic_MB": round(bytes_alg / 1e6, 1), "algorithmic_GBps": round(bytes_alg / t / 1e6, 1),
        "frac_of_hbm_peak_algorithmic": round(bytes_alg / t / 1e6 / PEAK_HBM, 4),
        "moved_MB_incl_G_blocks": round((bytes_alg + bytes_g) / 1e6, 1), "frac_of_hbm_peak_moved": round((bytes_alg + bytes_g) / t / 1e6 / PEAK_HBM, 4),
        "t_min_tensor_us": round((fl_f + fl_b) / PEAK_TF / 1e6, 1), "t_min_hbm_us": round(bytes_alg / PEAK_HBM / 1e3, 1),
        "roofline_frac_max_bound": round(max((fl_f + fl_b) / PEAK_TF / 1e9, bytes_alg / PEAK_HBM / 1e6) / t, 3),
        "torch_unfused_fp32_ms": round(t_ref, 3), "torch_unfused_bf16_ms": round(t_ref16, 3),
        "speedup_vs_torch_fp32": round(t_ref / t, 2), "speedup_vs_torch_bf16": round(t_ref16 / t, 2),
    }
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(1024, 8192, int(sys.argv[1]))
    else:
        run(1024, 8192, 768)
        run(1024, 8192, 64)
        run(1024, 1024, 768)   # the N = 1 bench shape
        run(512, 4096, 1024)   # configs[4]
