"""bench.py — image-text pairs/s, forward+backward, of the ViT+BERT contrastive hot path (BASELINE.json `metric`).

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (the reference algorithm on the host CPU cores)

A "step" = one fused forward + contrastive loss + backward over one synthetic batch (optimizer step and data loading
excluded). N = 1 runs BASELINE.json configs[1]: M2_Encoder ViT-L/14 + BERT-base bf16, per-GPU batch
1024, local contrastive; N > 1 keeps 1024 pairs per GPU (weak scaling) and contrasts over the all-gathered global batch
(configs[2] at N = 8), with the NCCL all-reduce of the parameter gradients inside the timed region (flat buckets after backward by
default, `--grad-sync ddp` = DistributedDataParallel's overlapped buckets; N > 1 lines carry per-rank step times and GEMM rates).

JSON keys beyond the base contract:  roofline (dominant kernel = the tcgen05 GEMM, timed live per launch with CUDA
events), cpu_baseline (oracle port on the host cores, bounded sample), e2e (public-API call with pinned HOST inputs,
H2D + loss D2H inside the timed region), clocks, gpu_launches.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "image-text pairs/sec fwd+bwd"
FLOP_PER_PAIR = {  # algorithmic fwd+bwd FLOPs per pair, SURVEY.md §8d (GEMMs 24*W^2 + attention 4*L*W per token-layer, x3)
    "ViT-L-14": 526.0e9,
    "ViT-B-16": 145.0e9,
    # ViT-H/14 at 224 (width 1280, 32 layers, 257 tokens, head_dim 80) + BERT-large (width 1024, 24 layers) — configs[4]'s towers at the
    # resolution CONFIGS["ViT-H-14"] ships (cn_model.py:95-113); attention runs on the mma.sync fallback kernels (head_dim 80)
    "ViT-H-14": 3 * ((24 * 1280**2 + 4 * 257 * 1280) * 32 * 257 + (24 * 1024**2 + 4 * 77 * 1024) * 24 * 77),
    # M2-Encoder (BEiT-3 multiway, 197 image tokens + 52 text tokens through the same 21+3 / 9+3 layers), same counting rule
    "M2-Encoder-1B": 3 * (24 * 1024**2 + 4 * 197 * 1024) * 24 * 197 + 3 * (24 * 1024**2 + 4 * 52 * 1024) * 24 * 52,
    "M2-Encoder-0.4B": 3 * (24 * 768**2 + 4 * 197 * 768) * 12 * 197 + 3 * (24 * 768**2 + 4 * 52 * 768) * 12 * 52,
}


def synth_batch(B, res, L, vocab, seed, device="cpu"):
    """SURVEY.md §8d synthetic inputs: N(0,1) images; ids with [CLS]=101, [SEP]=102 at the end of a U{8..L} span, [PAD]=0 after."""
    g = torch.Generator().manual_seed(seed)
    image = torch.randn(B, 3, res, res, generator=g)
    ids = torch.randint(1, vocab, (B, L), generator=g)
    ids[:, 0] = 101
    lens = torch.randint(8, L + 1, (B,), generator=g)
    pos = torch.arange(L)[None, :]
    ids[pos == (lens[:, None] - 1)] = 102
    ids[pos >= lens[:, None]] = 0
    return image.to(device), ids.to(device)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        d = json.load(open(path))
        return d.get("bf16_tflops_sustained", 1406.8), d.get("hbm_gbs", 6489.3), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


VTP_FRAMES = 8


class _VtpStep(torch.nn.Module):
    """BASELINE.json configs[3] geometry: base_vtp video-text retrieval (arch 'clip'), ViT-B/16 over 8 frames + BERT-base, level-1 MIL-NCE +
    level-2 cross-modal scoring with hard-negative mining ('top_k') and 'median' row weights — one training forward of
    b200mm.vtp.B200VideoTextRetrieval behind the (frames, text) -> loss interface of the bench."""

    def __init__(self, model):
        super().__init__()
        self.m = model

    def contrastive_loss(self, image, text, group=None):
        B = text.shape[0]
        F_ = image.shape[0] // B
        img_input = dict(image_data=image.view(B, F_, *image.shape[1:]),
                         image_pad_mask=torch.zeros((B, F_) + tuple(image.shape[2:]), dtype=torch.bool, device=image.device),  # "no padding"
                         image_n_clips=[1] * B, image_num_frames=[F_] * B)
        mask = (text != 0).long()
        caption = dict(caption_raw_input_ids=text, caption_input_ids=text, caption_input_mask=mask)
        out = self.m(img_input, caption)
        return out["losses"]["level1_similarity_loss"] + out["losses"]["level2_similarity_loss"]


class _M2Step(torch.nn.Module):
    """M2-Encoder ITC step behind the (image, text) -> loss interface of the bench: text_masks = ids != [PAD]."""

    def __init__(self, m2):
        super().__init__()
        self.m2 = m2

    def parameters(self, recurse=True):
        return self.m2.parameters(recurse)

    def contrastive_loss(self, image, text, group=None):
        return self.m2.itc_loss(image, text, (text != 0).long(), group)


def build_model(name, device, ckpt_every, keep_act=0, keep_ln=0, image_res=0, dropout=0.0):
    from b200mm.modules import CNCLIP, CONFIGS, M2_CONFIGS, M2Encoder

    if name == "base_vtp-ViT-B-16":
        from b200mm import vtp

        c = dict(CONFIGS["ViT-B-16"])
        vcfg = dict(training_stage="stage1+stage2", arch_type="clip", hidden_size=c["text_hidden_size"], with_moco=False,
                    hard_example_mining=True, re_sample_method="top_k", re_weight_method="median",
                    image_encoder=dict(type="B200VitImageEncoder", params=dict(
                        model_name="-", input_resolution=c["image_resolution"], patch_size=c["vision_patch_size"], width=c["vision_width"],
                        layers=c["vision_layers"], out_dim=c["embed_dim"], pretrained=False)),
                    text_encoder=dict(type="B200RobertBertEncoder", params=dict(
                        pretrained=False, hidden_size=c["text_hidden_size"], intermediate_size=c["text_intermediate_size"],
                        num_hidden_layers=c["text_num_hidden_layers"], num_attention_heads=c["text_num_attention_heads"],
                        vocab_size=c["vocab_size"], max_position_embeddings=c["text_max_position_embeddings"],
                        out_dim=c["text_hidden_size"], hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)))
        torch.manual_seed(0)
        m = vtp.B200VideoTextRetrieval(vcfg).to(device).to(torch.bfloat16).train()
        return _VtpStep(m), dict(image_resolution=c["image_resolution"], vocab_size=c["vocab_size"], vision_layers=c["vision_layers"],
                                 frames=VTP_FRAMES)
    if name in M2_CONFIGS:
        cfg = dict(M2_CONFIGS[name])
        torch.manual_seed(0)
        m2 = M2Encoder(**cfg).to(device).to(torch.bfloat16).train()
        if ckpt_every > 0:
            m2.set_grad_checkpointing(True)
        if keep_act > 0:
            m2.set_keep_activation(keep_act)
        return _M2Step(m2), dict(image_resolution=cfg["image_size"], vocab_size=cfg["vocab_size"], vision_layers=cfg["encoder_layers"])

    cfg = dict(CONFIGS[name])
    if image_res:
        cfg["image_resolution"] = image_res  # e.g. 336 for BASELINE.json configs[4]; positional embeddings are random-init at that size
    # the headline runs with p = 0 (same as the reference arm it is compared with); `--dropout 0.1` = the reference's training config
    cfg["text_hidden_dropout_prob"] = dropout
    cfg["text_attention_probs_dropout_prob"] = dropout
    torch.manual_seed(0)  # identical weights on every rank
    model = CNCLIP(**cfg)
    model = model.to(device).to(torch.bfloat16).train()
    if ckpt_every > 0:
        model.visual.set_grad_checkpointing(True, every=ckpt_every)
    if keep_act > 0:
        model.visual.set_keep_activation(keep_act)
    if keep_ln > 0:
        model.visual.set_keep_layernorm(keep_ln)
    return model, cfg


class TrainStep(torch.nn.Module):
    """The unit DDP wraps: forward = both encoders + fused global contrastive loss."""

    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, image, text):
        return self.model.contrastive_loss(image, text)


def run_ours(args):
    import torch.distributed as dist

    import b200mm
    from b200mm import ops
    from b200mm.gradcache import cnclip_gradcache_step

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert world == args.gpus or world == 1, f"WORLD_SIZE={world} but --gpus {args.gpus}"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    b200mm._lib.check(b200mm._lib.load().b200mm_check_device(), "b200mm_check_device")

    B, L = args.batch, args.seq_len
    model, cfg = build_model(args.model, device, args.ckpt_every, args.keep_act, args.keep_ln, args.image_res, args.dropout)
    res = cfg["image_resolution"]
    step_mod = TrainStep(model)
    if world > 1 and args.micro_batch == 0 and args.grad_sync == "ddp":
        step_mod = torch.nn.parallel.DistributedDataParallel(step_mod, device_ids=[local_rank], gradient_as_bucket_view=True,
                                                             static_graph=True)
    image_h, text_h = synth_batch(B * cfg.get("frames", 1), res, L, cfg["vocab_size"], 1234 + rank)
    text_h = text_h[:B].contiguous()
    image_h = image_h.to(torch.bfloat16).pin_memory()
    text_h = text_h.pin_memory()
    image_d, text_d = image_h.to(device), text_h.to(device)

    def step(img, txt):
        for p in model.parameters():
            p.grad = None
        if args.micro_batch > 0:
            # GradCache two-pass driver (full-batch negatives at micro-batch memory; the extra forward is NOT counted as useful work)
            return cnclip_gradcache_step(model, img, txt, args.micro_batch)
        loss = step_mod(img, txt)
        loss.backward()
        if world > 1 and args.grad_sync == "flat":
            # A/B against DDP's bucketed overlap: the parameter gradients averaged in flat buckets AFTER backward (nothing shares the SMs
            # with the GEMMs; the all-reduce is exposed instead)
            from b200mm.gradcache import allreduce_grads

            allreduce_grads(list(model.parameters()))
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident timing (value) ----------------
    try:
        for _ in range(args.warmup):
            step(image_d, text_d)
    except torch.OutOfMemoryError:
        # the keep-activation policy is a memory-for-time knob: fall back to recomputing every block's activated hidden
        if (args.keep_act == 0 and args.keep_ln == 0) or not hasattr(model, "visual"):
            raise
        for p in model.parameters():
            p.grad = None
        torch.cuda.empty_cache()
        args.keep_act = args.keep_ln = 0
        model.visual.set_keep_activation(0)
        model.visual.set_keep_layernorm(0)
        for _ in range(args.warmup):
            step(image_d, text_d)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ops.LAUNCHES = 0
    ops.GEMM_PROFILE = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = step(image_d, text_d)
    e1.record()
    barrier()
    ms_rank = e0.elapsed_time(e1)
    ms_total = max_over_ranks(ms_rank)
    clocks = sampler.stop() if rank == 0 else None
    launches = ops.LAUNCHES
    prof, ops.GEMM_PROFILE = ops.GEMM_PROFILE, None
    loss_val = float(loss.detach())
    ms_step = ms_total / args.steps
    pairs_per_s = B * world / (ms_step * 1e-3)

    # roofline of the dominant kernel (tcgen05 GEMM): algorithmic FLOPs / CUDA-event duration, over every launch of the timed region
    gemm_flops = sum(r[0] for r in prof)
    gemm_ms = sum(r[1].elapsed_time(r[2]) for r in prof)
    peak_tf, peak_hbm, peak_src = peaks()
    gemm_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0

    dominant = dominant_launch(prof, peak_tf) if rank == 0 else None
    per_rank = None
    if world > 1:
        # where the weak-scaling loss comes from: every rank's own step time and GEMM rate (the reported step is the slowest rank's)
        t = torch.tensor([ms_rank / args.steps, gemm_tf, gemm_ms / args.steps], device=device, dtype=torch.float64)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank = {"ms_per_step": [round(float(x[0]), 2) for x in allt], "gemm_tflops": [round(float(x[1]), 1) for x in allt],
                    "gemm_ms_per_step": [round(float(x[2]), 2) for x in allt], "grad_sync": args.grad_sync}

    if args.profile and rank == 0:
        write_profile(args, prof, step, image_d, text_d, B)

    # ---------------- end-to-end through the public API with HOST inputs (e2e) ----------------
    def e2e_step():
        img = image_h.to(device, non_blocking=True)
        txt = text_h.to(device, non_blocking=True)
        ls = step(img, txt)
        return float(ls.detach())  # device -> host read of the step's result

    e2e_steps = 0 if args.skip_e2e else args.steps
    e2e_step()
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(max(1, e2e_steps)):
        e2e_step()
    t1.record()
    barrier()
    e2e_ms = max_over_ranks(t0.elapsed_time(t1)) / max(1, e2e_steps)
    e2e_val = B * world / (e2e_ms * 1e-3)
    peak_mem = torch.cuda.max_memory_allocated(device) / 2**30

    out = None
    if rank == 0:
        fpp = FLOP_PER_PAIR.get(args.model)
        if args.model in ("ViT-H-14", "ViT-L-14") and (res != 224 or L != 77):
            # the table entries are for 224^2 / 77 tokens: same counting rule at the resolution / text length actually run (configs[4]: 336^2)
            vw, vl, tw, tl, ps = (1280, 32, 1024, 24, 14) if args.model == "ViT-H-14" else (1024, 24, 768, 12, 14)
            li = (res // ps) ** 2 + 1
            fpp = 3 * ((24 * vw**2 + 4 * li * vw) * vl * li + (24 * tw**2 + 4 * L * tw) * tl * L)
        if args.model.startswith("base_vtp"):
            # per pair, fwd+bwd = 3 x (8 frames of ViT-B/16 + BERT-base at L + B cross-encoder passes over L + 2 tokens): hard mining scores every
            # text against B videos, so the cross-encoder work per pair grows with the per-GPU batch
            w = 768
            fpp = 3.0 * (VTP_FRAMES * 35.1e9 + (24 * w * w + 4 * L * w) * 12 * L + B * (24 * w * w + 4 * (L + 2) * w) * 12 * (L + 2))
        out = {
            "metric": METRIC, "value": round(pairs_per_s, 2), "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic (seeded N(0,1) images, random ids with [CLS]/[SEP]/[PAD]; random-init weights, reference init)",
            "config": {"workload": (f"BASELINE.json configs[{4 if args.model == 'ViT-H-14' else (1 if world == 1 else 2)}]{' towers' if args.model == 'ViT-H-14' and (res != 336 or B * world != 4096) else ''}: CNCLIP {args.model} + BERT-{'large' if args.model == 'ViT-H-14' else 'base'}, fwd + fused contrastive loss + bwd"
                                    if args.model in FLOP_PER_PAIR and not args.model.startswith(("M2", "base_vtp")) else
                                    (f"BASELINE.json configs[3] geometry: base_vtp video-text (arch clip) ViT-B/16 x {VTP_FRAMES} frames + BERT-base, level-1 MIL-NCE + "
                                     f"level-2 cross-modal scoring of B x B mined pairs (86-token sequences) + weighted MIL-NCE, fwd + bwd; {B} pairs per GPU")
                                    if args.model.startswith("base_vtp") else
                                    f"prj/M2_Encoder {args.model} (BEiT-3 multiway): infer_image + infer_text + symmetric ITC on both head pairs, fwd + bwd"),
                       "model": args.model, "per_gpu_batch": B, "global_batch": B * world, "image_res": res, "seq_len": L,
                       "parallelism": f"dp{world}" + ((" + embedding all-gather / grad reduce-scatter, " + ("DDP bucketed grad all-reduce overlapped with backward" if args.grad_sync == "ddp" else "flat-bucket NCCL grad all-reduce after backward") + ", in the timed region") if world > 1 else ""),
                       "dropout": args.dropout, "recompute": f"GradCache two-pass, micro-batch {args.micro_batch} (second forward not counted)" if args.micro_batch else f"checkpoint every {args.ckpt_every} ViT block(s)" if args.ckpt_every else f"none (selective save: LN outputs recomputed in {max(0, cfg['vision_layers'] - args.keep_ln)} and the activated MLP hidden in {max(0, cfg['vision_layers'] - args.keep_act)} of {cfg['vision_layers']} ViT blocks)",
                       "l2_policy": "inputs and activations (>= 0.5 GB per tensor) exceed the 126 MB L2; no flush needed",
                       "loss": round(loss_val, 5), "peak_mem_gib": round(peak_mem, 1)},
            "e2e": {"value": round(e2e_val, 2), "unit": "pairs/s", "ms_per_step": round(e2e_ms, 3),
                    "h2d_bytes_per_step": image_h.numel() * 2 + text_h.numel() * 8, "d2h_bytes_per_step": 4},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (all tcgen05 GEMM launches of the timed region)",
                         "achieved": round(gemm_tf, 1), "peak": peak_tf, "unit": "TFLOP/s", "frac": round(gemm_tf / peak_tf, 4),
                         "peak_source": peak_src, "launches": len(prof), "share_of_step": round(gemm_ms / ms_total, 4),
                         "traffic": (dominant.get("traffic") or (dominant.get("nearest_captured_launch") or {}).get("traffic")) if dominant else None,
                         "traffic_kernel": ((dominant["kernel"] if dominant.get("traffic") else (dominant.get("nearest_captured_launch") or {}).get("kernel"))
                                            if dominant else None),
                         "dominant_launch": dominant,
                         "whole_step_frac": round(pairs_per_s / world * fpp / 1e12 / peak_tf, 4) if fpp else None},
        }
        if per_rank is not None:
            out["per_rank"] = per_rank
        headline = args.model in FLOP_PER_PAIR and not args.model.startswith(("M2", "base_vtp"))
        if world == 1 and not args.no_gpu_baseline and headline and not args.image_res:
            del step_mod, model
            torch.cuda.empty_cache()
            try:
                out["gpu_eager_baseline"] = gpu_eager_baseline(args, device)
            except Exception as e:  # noqa: BLE001 — a comparison point must never take the bench line down
                out["gpu_eager_baseline"] = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
        if world == 1 and not args.no_cpu_baseline and headline:
            try:
                out["cpu_baseline"] = cpu_baseline(args, steps=2, warmup=0)  # ~10 s of CPU work at batch 16 on 16 cores
            except Exception as e:  # noqa: BLE001
                out["cpu_baseline"] = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out), flush=True)


def dominant_launch(gemm_prof, peak_tf):
    """The GEMM shape/epilogue that takes the most time in the timed region: its own TFLOP/s, algorithmic bytes per launch, and the DRAM
    traffic ncu measured for one launch of the SAME kernel instantiation at the SAME shape (profiles/ncu_traffic_latest.json, produced from
    the committed `ncu --set full` capture of tools/prof_kernels.py) — null when that capture does not cover this shape."""
    from collections import defaultdict

    groups = defaultdict(lambda: [0, 0.0, 0.0])
    for flops, a, b, splits, tag in gemm_prof:
        g = groups[tag + (splits,)]
        g[0] += 1
        g[1] += a.elapsed_time(b)
        g[2] += flops
    if not groups:
        return None
    path = os.path.join(ROOT, "profiles", "ncu_traffic_latest.json")
    cap = json.load(open(path)) if os.path.isfile(path) else {}

    def describe(tag, n, ms, fl):
        M, N, K, a_mn, b_mn, bias, act, aux, dact, res, f32, splits = tag
        tf = fl / (ms * 1e-3) / 1e12
        alg = 2 * (M * K + N * K) + (4 if f32 else 2) * M * N * (1 + aux) + 2 * M * N * (dact + res) + 2 * N * bias
        flavor = (bias) | (aux << 1) | (res << 2) | (dact << 3) | (f32 << 4) | (act << 6)
        name = f"gemm_tcgen05_kernel<{a_mn}, {b_mn}, 0, 2, {flavor}>"
        out = {"kernel": name, "shape_MNK": [M, N, K], "launches": n, "avg_ms": round(ms / n, 4), "achieved": round(tf, 1), "frac": round(tf / peak_tf, 4),
               "algorithmic_bytes": alg, "traffic": None}
        k = cap.get("kernels", {}).get(name)
        if k and splits == 1 and cap.get("shapes", {}).get(name) == [M, N, K]:
            out["traffic"] = k["dram_bytes"]
            out["traffic_over_algorithmic"] = round(k["dram_bytes"] / alg, 3)
            out["traffic_source"] = cap.get("source")
        return out

    ranked = sorted(groups.items(), key=lambda kv: -kv[1][1])
    out = describe(ranked[0][0], *ranked[0][1])
    if out["traffic"] is None:
        # the capture on file does not cover the launch with the largest time share: report, labelled as such, the measured traffic of the
        # next-largest launch that it does cover (same GEMM family; the time shares of the top flavours are within a few per cent)
        total = sum(v[1] for v in groups.values())
        for tag, (n, ms, fl) in ranked[1:6]:
            alt = describe(tag, n, ms, fl)
            if alt["traffic"] is not None:
                alt["share_of_gemm_time"] = round(ms / total, 4)
                out["share_of_gemm_time"] = round(ranked[0][1][1] / total, 4)
                out["nearest_captured_launch"] = alt
                break
    return out


def write_profile(args, gemm_prof, step, image_d, text_d, B):
    """Per-op CUDA-event breakdown of one extra step + per-shape GEMM table -> gpurun_out/op_profile.json (diagnostics only)."""
    import b200mm
    from collections import defaultdict

    recs = []
    b200mm._lib.enable_profile(recs)
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    w0.record()
    step(image_d, text_d)
    w1.record()
    torch.cuda.synchronize()
    b200mm._lib.disable_profile()
    by_op = defaultdict(lambda: [0, 0.0])
    for name, a, b in recs:
        by_op[name][0] += 1
        by_op[name][1] += a.elapsed_time(b)
    shapes = defaultdict(lambda: [0, 0.0, 0.0])
    for flops, a, b, splits, tag in gemm_prof:
        k = str(tag + (splits,))
        shapes[k][0] += 1
        shapes[k][1] += a.elapsed_time(b)
        shapes[k][2] += flops
    table = sorted(((k, n, ms, fl / (ms * 1e-3) / 1e12 if ms > 0 else 0) for k, (n, ms, fl) in shapes.items()), key=lambda r: -r[2])
    out = {"step_ms_profiled": w0.elapsed_time(w1), "ops": {k: {"calls": v[0], "ms": round(v[1], 3)} for k, v in sorted(by_op.items(), key=lambda kv: -kv[1][1])},
           "gemm_shapes(M,N,K,a_mn,b_mn,bias,act,aux,dact,res,f32,splits)": [{"shape": k, "calls": n, "ms_total": round(ms, 3), "tflops": round(tf, 1)} for k, n, ms, tf in table]}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"op_profile_b{B}.json"), "w"), indent=1)


def _reference_cnclip(args):
    """The UNMODIFIED reference CNCLIP (antmmf/modules/vision/backbone/clip/cn_model.py) imported from baseline/_ref/ — the byte copies
    baseline/install_ref.py makes of the reference's plain-PyTorch files (sha256 in baseline/_ref/MANIFEST.json). None when the copies are
    absent (then the oracle port is timed instead)."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    try:
        import ref_import
    finally:
        sys.path.pop(0)
    if not ref_import.available():
        return None
    return ref_import.build_cnclip(args.model, seed=0, dropout=0.0)


def _reference_low_precision(model, dtype):
    """The reference's own low-precision recipe (convert_weights, cn_model.py:276-305) with the dtype swapped for bf16: Conv / Linear /
    nn.MultiheadAttention parameters, the whole BertModel, `text_projection` and `proj` go to `dtype`; the ViT LayerNorms keep fp32
    parameters (their forward casts the input to fp32, clip/model.py:213-219)."""
    from torch import nn

    def _convert(l):
        if isinstance(l, (nn.Conv1d, nn.Conv2d, nn.Linear)):
            l.weight.data = l.weight.data.to(dtype)
            if l.bias is not None:
                l.bias.data = l.bias.data.to(dtype)
        if isinstance(l, nn.MultiheadAttention):
            for attr in ["in_proj_weight", "q_proj_weight", "k_proj_weight", "v_proj_weight", "in_proj_bias", "bias_k", "bias_v"]:
                t = getattr(l, attr)
                if t is not None:
                    t.data = t.data.to(dtype)
        if type(l).__name__ == "BertModel":
            l.to(dtype)
        for name in ["text_projection", "proj"]:
            t = getattr(l, name, None)
            if isinstance(t, torch.Tensor):
                t.data = t.data.to(dtype)

    model.apply(_convert)
    return model


def gpu_eager_baseline(args, device):
    """SURVEY.md §8d "GPU comparison point": the reference's own modules after `.cuda().bfloat16()` under eager PyTorch on the SAME GPU
    (nn.MultiheadAttention -> torch's fused SDPA, BertSelfAttention's matmul-softmax, F.cross_entropy on materialised logits), fwd + bwd at a
    batch that fits eager's activation memory. Same model / sequence length / synthetic data as the B200 arm."""
    model = _reference_cnclip(args)
    if model is None:
        return {"unavailable": "baseline/_ref not installed"}
    model = _reference_low_precision(model.to(device), torch.bfloat16).train()
    Bc = args.eager_batch
    cfg_res = model.visual.input_resolution if hasattr(model.visual, "input_resolution") else 224
    image, text = synth_batch(Bc, cfg_res, args.seq_len, model.bert.embeddings.word_embeddings.num_embeddings, 1234)
    image, text = image.to(device).to(torch.bfloat16), text.to(device)
    target = torch.arange(Bc, device=device)

    def one_step():
        for p_ in model.parameters():
            p_.grad = None
        _, _, lpi, lpt = model(image, text)
        loss = 0.5 * (torch.nn.functional.cross_entropy(lpi.float(), target) + torch.nn.functional.cross_entropy(lpt.float(), target))
        loss.backward()
        return loss

    try:
        for _ in range(2):
            one_step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            one_step()
        e1.record()
        torch.cuda.synchronize()
    except torch.OutOfMemoryError:
        return {"unavailable": f"eager reference modules run out of memory at batch {Bc}"}
    ms = e0.elapsed_time(e1) / 3
    return {"value": round(Bc / ms * 1e3, 1), "unit": "pairs/s", "ms_per_step": round(ms, 2), "batch": Bc,
            "kind": "unmodified reference CNCLIP (baseline/_ref) on the GPU in bf16 by the reference's own convert_weights rule (ViT LayerNorm fp32), eager PyTorch incl. torch SDPA inside nn.MultiheadAttention; 3 steps after 2 warm-up"}


def cpu_baseline(args, steps, warmup):
    """The reference's CPU path on a BOUNDED sample of the workload (same model / config / sequence length, batch `cpu_batch` instead of
    1024), all host threads: the reference's own modules from baseline/_ref/ when they are installed (kind "reference": CNCLIP.forward +
    the symmetric cross-entropy of its logits + backward, fp32 eager PyTorch), else the oracle port oracle/restated.py (kind "port")."""
    from b200mm.modules import CONFIGS

    cfg = dict(CONFIGS[args.model])
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    Bc = args.cpu_batch
    image, text = synth_batch(Bc, cfg["image_resolution"], args.seq_len, cfg["vocab_size"], 1234)
    ref_model = _reference_cnclip(args)
    if ref_model is not None:
        ref_model = ref_model.float().train()
        target = torch.arange(Bc)

        def one_step():
            for p_ in ref_model.parameters():
                p_.grad = None
            _, _, lpi, lpt = ref_model(image, text)
            loss = 0.5 * (torch.nn.functional.cross_entropy(lpi, target) + torch.nn.functional.cross_entropy(lpt, target))
            loss.backward()

        kind, what = "reference", "unmodified reference CNCLIP from baseline/_ref (cn_model.py / model.py / modeling_bert.py), fp32 eager"
    else:
        from b200mm.modules import CNCLIP
        from oracle import restated

        torch.manual_seed(0)
        model = CNCLIP(**cfg)  # parameter container only (reference init); the arithmetic below is the oracle's
        sd = {k: v.detach().clone().requires_grad_(torch.is_floating_point(v)) for k, v in model.state_dict().items()}
        del model
        vh = cfg["vision_width"] // cfg.get("vision_head_width", 64)

        def one_step():
            for v in sd.values():
                v.grad = None
            _, _, logits, _ = restated.cnclip_forward(sd, image, text, vh, cfg["text_num_attention_heads"])
            restated.symmetric_info_nce(logits).backward()

        kind, what = "port", "oracle/restated.py fp32 eager"
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        one_step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    best = min(times)
    return {"value": round(Bc / best, 3), "unit": "pairs/s", "cores": threads, "kind": kind,
            "sample": f"{what}, same model/seq_len, batch {Bc} instead of {args.batch}; best of {steps} step(s) after {warmup} warm-up",
            "s_per_step": round(best, 2)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores — the unmodified modules installed under
    baseline/_ref/ by baseline/install_ref.py (they travel to the GPU box with the snapshot); the oracle port only if they are absent."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline(args, steps=max(1, args.steps), warmup=min(1, args.warmup))
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": round(cb["s_per_step"] * 1e3, 1), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"CNCLIP {args.model} + BERT-base fwd + symmetric InfoNCE + bwd, CPU eager fp32 ({cb['kind']}), bounded sample batch {args.cpu_batch}",
                      "model": args.model, "seq_len": args.seq_len},
           "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200mm", choices=["b200mm", "reference"])
    ap.add_argument("--model", default="ViT-L-14")
    ap.add_argument("--batch", type=int, default=1024, help="pairs per GPU")
    ap.add_argument("--seq-len", type=int, default=77)
    ap.add_argument("--image-res", type=int, default=0, help="override the model's image resolution (336 = BASELINE.json configs[4])")
    ap.add_argument("--dropout", type=float, default=0.0, help="BERT hidden / attention-probability dropout (the reference's CN-CLIP configs train with 0.1)")
    ap.add_argument("--grad-sync", default="flat", choices=["ddp", "flat"],
                    help="N > 1: parameter gradients averaged in flat bf16 buckets after backward (default: measured faster on 8 x B200, the NCCL kernels of "
                         "DDP's overlapped buckets take SMs from the persistent tcgen05 GEMMs: profiles/r02e_bench_8gpu_{ddp,flat}.log), or 'ddp' = "
                         "torch DistributedDataParallel's bucketed all-reduce overlapped with backward")
    ap.add_argument("--ckpt-every", type=int, default=0, help="re-run every k-th ViT block in backward (0 = never)")
    ap.add_argument("--micro-batch", type=int, default=0, help="> 0: run the step through the GradCache two-pass driver with this micro-batch")
    ap.add_argument("--keep-act", type=int, default=0, help="ViT blocks that keep the activated MLP hidden instead of recomputing it (memory for time)")
    ap.add_argument("--keep-ln", type=int, default=16, help="ViT blocks that keep both LayerNorm outputs instead of recomputing them (memory for time)")
    ap.add_argument("--cpu-batch", type=int, default=16, help="batch of the bounded CPU-baseline sample (fp32 eager needs ~0.6 GB of host RAM per pair)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the eager-PyTorch bf16 run of the reference modules on the same GPU")
    ap.add_argument("--eager-batch", type=int, default=128, help="batch of the eager GPU baseline (eager keeps every activation)")
    ap.add_argument("--profile", action="store_true", help="write gpurun_out/op_profile_b<B>.json (per-op CUDA-event breakdown)")
    ap.add_argument("--skip-e2e", action="store_true", help="diagnostics only: skip the host-input leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
