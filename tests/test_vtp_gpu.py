"""GPU parity of b200mm.vtp — the base_vtp video-text retrieval model (arch 'clip': ViT frames + BERT, level-1 MIL-NCE / MoCo, level-2
cross-modal scoring with hard-negative mining; BASELINE.json configs[3] geometry at toy size) — against the oracle composition of
tests/vtp_common.py, all kernels through the C-ABI."""
import pytest
import torch

from tests import vtp_common

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rel_l2(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-12))


def _model(cfg):
    from b200mm import vtp

    torch.manual_seed(0)
    model = vtp.B200VideoTextRetrieval(cfg)
    vtp_common.randomize(model)
    return model.cuda().to(BF).train()


@pytest.mark.parametrize("hard,method", [(True, "top_k"), (True, "nearliest"), (False, "top_k")])
def test_vtp_forward_matches_oracle_composition(hard, method):
    import b200mm

    cfg = vtp_common.make_config(hard=hard, re_sample=method)
    model = _model(cfg)
    img_input, caption = vtp_common.make_batch(device="cuda")
    out = model(img_input, caption)
    chosen = b200mm.cross.hard_mining_indices(out["l1_simi"], 0, 6, method).cpu() if hard else None
    ref, _ = vtp_common.oracle_forward(model, img_input, caption, cfg, chosen=chosen)
    assert set(out) == {"losses", "l1_simi", "l2_simi"}
    assert rel_l2(out["l1_simi"], ref["l1_simi"]) < 2e-2, rel_l2(out["l1_simi"], ref["l1_simi"])
    l1 = float(out["losses"]["level1_similarity_loss"])
    assert abs(l1 - float(ref["l1_loss"])) < 2e-2 * float(ref["l1_loss"]), (l1, float(ref["l1_loss"]))
    assert out["l2_simi"].shape == (6, 6) and out["l2_simi"].dtype == torch.float32
    assert rel_l2(out["l2_simi"], ref["l2_simi"]) < 3e-2, rel_l2(out["l2_simi"], ref["l2_simi"])
    l2 = float(out["losses"]["level2_similarity_loss"])
    # the weighted loss uses weights from THIS run's level-1 diagonal; the oracle's come from its own fp32 diagonal (same rule)
    assert abs(l2 - float(ref["l2_loss"])) < 3e-2 * float(ref["l2_loss"]), (l2, float(ref["l2_loss"]))
    (out["losses"]["level1_similarity_loss"] + out["losses"]["level2_similarity_loss"]).backward()
    for n, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad.float()).all(), n


def test_vtp_gradient_routing_matches_oracle():
    cfg = vtp_common.make_config(hard=False)
    model = _model(cfg)
    img_input, caption = vtp_common.make_batch(device="cuda")
    g = torch.Generator().manual_seed(9)
    R1, R2, R3 = torch.randn(6, vtp_common.HID, generator=g), torch.randn(6, vtp_common.HID, generator=g), torch.randn(6, 6, generator=g)
    cap_input, vis_input, _, _ = model.module.get_l2_input(img_input, caption)
    out = model.forward_stage(cap_input + (caption,), vis_input + (img_input,), True)
    ((cap_input[2].float() * R1.cuda()).sum() + (vis_input[2].float() * R2.cuda()).sum() + (out["l2_simi"] * R3.cuda()).sum()).backward()
    ref, leaves = vtp_common.oracle_forward(model, img_input, caption, cfg)
    ((ref["text_n"] * R1).sum() + (ref["clip_n"] * R2).sum() + (ref["l2_simi"] * R3).sum()).backward()
    named = model.state_dict(keep_vars=True)
    checked = 0
    for n, leaf in leaves.items():
        if n not in named or leaf.grad is None or float(leaf.grad.abs().max()) < 1e-4:
            continue
        assert named[n].grad is not None, n
        assert rel_l2(named[n].grad, leaf.grad) < 0.12, (n, rel_l2(named[n].grad, leaf.grad))
        checked += 1
    assert checked > 40


def test_vtp_moco_stage1_runs_and_enqueues():
    """with_moco (the reference default, univl_video_ret.py:29,263-312): momentum key encoders, both queue losses on the fused kernels,
    enqueue of the gathered keys; the loss equals the oracle's moco_nce on the same embeddings / queues."""
    from oracle import restated

    cfg = vtp_common.make_config(stage="stage1", hard=False, with_moco=True)
    model = _model(cfg)
    img_input, caption = vtp_common.make_batch(device="cuda")
    out = model(img_input, caption)
    mu = model.moco_utils
    assert mu is not None and int(mu.txt_queue_ptr) == 6 and int(mu.img_queue_ptr) == 6
    loss = out["losses"]["level1_similarity_loss"]
    assert torch.isfinite(loss) and float(loss) > 0
    loss.backward()
    assert model.module.img_encoder.visual.proj.grad is not None and model.module.text_encoder.text_projection.grad is not None
    # oracle on the same query / key embeddings and the queues BEFORE this step's enqueue (columns 0..5 were overwritten afterwards)
    with torch.no_grad():
        cap_input, vis_input, _, _ = model.module.get_l2_input(img_input, caption)
        q_t, q_v = cap_input[2].float().cpu(), vis_input[2].float().cpu()
        key_v = model.module.forward_img_encoder(**img_input, img_encoder=mu.img_encoder_k)["clip_feature"].float().cpu()
        key_t = model.module.forward_text_encoder(caption["caption_raw_input_ids"], caption["caption_input_mask"], txt_encoder=mu.txt_encoder_k)["pooled_output"].float().cpu()
    tq, iq = mu.txt_queue.float().cpu().clone(), mu.img_queue.float().cpu().clone()
    # negatives outside the overwritten columns are unchanged; compare the loss restricted to a queue where those columns are restored is
    # not possible after the fact, so bound instead: the enqueued keys are near-duplicates of the positives (momentum copy at step 1)
    ref_v = restated.moco_nce((q_v * key_t).sum(-1, keepdim=True), q_v @ tq[:, 6:].to(BF).float(), 0.05)
    ref_t = restated.moco_nce((q_t * key_v).sum(-1, keepdim=True), q_t @ iq[:, 6:].to(BF).float(), 0.05)
    ref = float((ref_v + ref_t) / 2)
    assert abs(float(loss) - ref) < 0.05 * ref, (float(loss), ref)
