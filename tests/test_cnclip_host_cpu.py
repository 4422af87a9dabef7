"""Host logic of the headline path (CNCLIP: ViT + BERT towers, block-level autograd glue with its save / recompute policies) without a
GPU: b200mm.modules.CNCLIP over the torch stand-ins of the kernels (tests/emulated_ops.py) must reproduce the golden vectors of the
unmodified reference for every memory policy — plain, block checkpointing, keep-activation, keep-LayerNorm."""
import os

import pytest
import torch

from oracle import restated
from tests import emulated_ops

BF = torch.bfloat16


def rel_l2(got, ref):
    got, ref = got.detach().float(), ref.detach().float()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-12))


@pytest.mark.parametrize("name", ["cnclip_tiny.pt", "cnclip_tiny_h80.pt"])
@pytest.mark.parametrize("policy", ["plain", "checkpoint", "keep_act", "keep_ln", "keep_both"])
def test_cnclip_glue_reproduces_reference_golden(golden_dir, name, policy):
    from b200mm.modules import CNCLIP

    fx = torch.load(os.path.join(golden_dir, name), weights_only=False)
    m = CNCLIP(**fx["config"])
    m.load_state_dict(fx["state_dict"])
    m = m.to(BF).train()
    m.set_grad_checkpointing(policy == "checkpoint")
    if policy in ("keep_act", "keep_both"):
        m.visual.set_keep_activation(1)
    if policy in ("keep_ln", "keep_both"):
        m.visual.set_keep_layernorm(2)
    with emulated_ops.patched():
        img, txt = m.encode_normalized(fx["image"].to(BF), fx["text"])
        assert rel_l2(img, fx["image_features"]) < 2e-2 and rel_l2(txt, fx["text_features"]) < 2e-2
        loss = m.contrastive_loss(fx["image"].to(BF), fx["text"])  # the production fused-loss host code (b200mm.contrastive)
        assert abs(float(loss) - float(fx["loss"])) < 3e-2 * float(fx["loss"])
        loss.backward()
    # calibrator: the reference arithmetic itself in eager bf16 (tiny random-init fixtures are ill-conditioned — the image features of
    # cnclip_tiny_h80 have pairwise cosine 0.96, so the contrastive gradient is a difference of nearly equal vectors; eager bf16 is 7 % off)
    sdb = {k: (v.detach().to(BF).requires_grad_(True) if torch.is_floating_point(v) else v) for k, v in fx["state_dict"].items()}
    cfg = fx["config"]
    _, _, e_logits, _ = restated.cnclip_forward(sdb, fx["image"].to(BF), fx["text"], cfg["vision_width"] // cfg["vision_head_width"],
                                                cfg["text_num_attention_heads"])
    restated.symmetric_info_nce(e_logits.float()).backward()
    checked = 0
    for n, p in m.named_parameters():
        ref = fx["grads"].get(n)
        if ref is None or float(ref.abs().max()) < 1e-5:
            continue
        if n == "logit_scale":  # scalar with heavy cancellation: absolute bound
            assert abs(float(p.grad) - float(ref)) < 5e-2 * max(1.0, abs(float(ref))), (float(p.grad), float(ref))
            continue
        assert p.grad is not None, n
        assert rel_l2(p.grad, ref) < max(8e-2, 3.0 * rel_l2(sdb[n].grad, ref)), (n, policy, rel_l2(p.grad, ref), rel_l2(sdb[n].grad, ref))
        checked += 1
    assert checked > 40
    assert float(m.bert.embeddings.word_embeddings.weight.grad[0].abs().max()) == 0.0  # padding_idx row


@pytest.mark.parametrize("in_dtype", [torch.float32, torch.float16])
def test_modules_accept_fp32_fp16_inputs_under_autocast(golden_dir, in_dtype):
    """Trainer coupling (SURVEY.md §8b iv): the unmodified trainer may call the model under fp16 autocast with fp32 / fp16 inputs and fp32
    master parameters (base_trainer.py:334-349); the modules cast at their boundary and give the same features as the bf16 call."""
    from b200mm.modules import CNCLIP

    fx = torch.load(os.path.join(golden_dir, "cnclip_tiny.pt"), weights_only=False)
    m = CNCLIP(**fx["config"])
    m.load_state_dict(fx["state_dict"])
    m.train()  # fp32 master parameters: cast to bf16 per call
    with emulated_ops.patched():
        ref_img, ref_txt = m.to(BF).encode_normalized(fx["image"].to(BF), fx["text"])
        m = m.float()
        with torch.autocast("cpu", dtype=torch.bfloat16):
            img, txt = m.encode_normalized(fx["image"].to(in_dtype), fx["text"])
    assert img.dtype == BF and txt.dtype == BF
    assert rel_l2(img, ref_img) < 1e-2 and rel_l2(txt, ref_txt) < 1e-2
    assert rel_l2(img, fx["image_features"]) < 2e-2
