"""Host logic of b200mm.vtp (the base_vtp video-text retrieval model, arch 'clip') without a GPU: registry construction from a
reference-style config, the reference's batch dictionaries, frame pooling → level-1 matrix → hard-negative selection → pair scoring →
weighted level-2 loss, over torch stand-ins of the kernels; compared with the oracle composition (tests/vtp_common.py)."""
import pytest
import torch

from oracle import restated
from tests import emulated_ops, vtp_common

BF = torch.bfloat16


def rel_l2(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-12))


@pytest.mark.parametrize("hard", [True, False])
def test_vtp_forward_matches_oracle_composition(hard, monkeypatch):
    import b200mm
    from b200mm import vtp

    cfg = vtp_common.make_config(hard=hard)
    torch.manual_seed(0)
    model = vtp.B200VideoTextRetrieval(cfg)
    vtp_common.randomize(model)
    model = model.to(BF).train()
    # state-dict prefixes of the reference model (univl_video_ret.py:22-28, univl_video_base.py:24-48)
    keys = set(model.state_dict())
    assert "similarity_dense.0.weight" in keys and "similarity_dense.2.bias" in keys
    assert "module.text_encoder.text_projection" in keys and "module.img_encoder.visual.conv1.weight" in keys
    assert "module.text_encoder.encoder.layer.0.attention.self.query.weight" in keys
    img_input, cap_input = vtp_common.make_batch()
    with emulated_ops.patched():
        out = model(img_input, cap_input)
        ref, leaves = vtp_common.oracle_forward(model, img_input, cap_input, cfg)
        assert set(out) == {"losses", "l1_simi", "l2_simi"} and set(out["losses"]) == {"level1_similarity_loss", "level2_similarity_loss"}
        assert rel_l2(out["l1_simi"], ref["l1_simi"]) < 2e-2
        assert abs(float(out["losses"]["level1_similarity_loss"]) - float(ref["l1_loss"])) < 2e-2 * float(ref["l1_loss"])
        if hard:  # the selection is made from this run's own level-1 matrix: re-score the oracle on the same selection
            chosen = b200mm.cross.hard_mining_indices(out["l1_simi"], 0, 6, "top_k")
            ref, leaves = vtp_common.oracle_forward(model, img_input, cap_input, cfg, chosen=chosen)
        assert out["l2_simi"].shape == (6, 6)
        assert rel_l2(out["l2_simi"], ref["l2_simi"]) < 3e-2, rel_l2(out["l2_simi"], ref["l2_simi"])
        assert abs(float(out["losses"]["level2_similarity_loss"]) - float(ref["l2_loss"])) < 2e-2 * float(ref["l2_loss"])
        (out["losses"]["level1_similarity_loss"] + out["losses"]["level2_similarity_loss"]).backward()
    for n, p in model.named_parameters():  # every parameter on the path receives a finite gradient
        if "k_encoder" in n:
            continue
        assert p.grad is not None and torch.isfinite(p.grad.float()).all(), n


def test_vtp_gradient_routing_matches_oracle(monkeypatch):
    """Gradient ROUTING through the composed model (towers <- embeddings, towers + head <- pair scores). The NCE losses of a random-init
    model are cancellation-dominated (all scores nearly equal), so the scalar differentiated here is a fixed random projection of the
    level-1 embeddings and of the full level-2 score matrix; the losses' own gradients are kernel-level tests."""
    from b200mm import vtp

    cfg = vtp_common.make_config(hard=False)
    torch.manual_seed(0)
    model = vtp.B200VideoTextRetrieval(cfg)
    vtp_common.randomize(model)
    model = model.to(BF).train()
    img_input, caption = vtp_common.make_batch()
    g = torch.Generator().manual_seed(9)
    R1, R2, R3 = torch.randn(6, vtp_common.HID, generator=g), torch.randn(6, vtp_common.HID, generator=g), torch.randn(6, 6, generator=g)
    with emulated_ops.patched():
        cap_input, vis_input, _, _ = model.module.get_l2_input(img_input, caption)
        out = model.forward_stage(cap_input + (caption,), vis_input + (img_input,), True)
        ((cap_input[2].float() * R1).sum() + (vis_input[2].float() * R2).sum() + (out["l2_simi"] * R3).sum()).backward()
    ref, leaves = vtp_common.oracle_forward(model, img_input, caption, cfg)
    ((ref["text_n"] * R1).sum() + (ref["clip_n"] * R2).sum() + (ref["l2_simi"] * R3).sum()).backward()
    named = model.state_dict(keep_vars=True)  # includes the aliased names (text_encoder.module.* = text_encoder.{embeddings,encoder}.*)
    checked = 0
    for n, leaf in leaves.items():
        if n not in named or leaf.grad is None or float(leaf.grad.abs().max()) < 1e-4:
            continue
        assert named[n].grad is not None, n
        assert rel_l2(named[n].grad, leaf.grad) < 0.12, (n, rel_l2(named[n].grad, leaf.grad))
        checked += 1
    assert checked > 30


def test_vtp_stage1_only_and_unsupported_arch():
    from b200mm import vtp

    cfg = vtp_common.make_config(stage="stage1", hard=False)
    m = vtp.B200VideoTextRetrieval(cfg)
    assert not hasattr(m, "similarity_dense") and m.module.with_cross_encoder is False
    with pytest.raises(NotImplementedError):
        vtp.B200VideoTextRetrieval(dict(cfg, arch_type="univl"))


def test_univl_model_plugin_interface(monkeypatch):
    """The registered model (antmmf registry.register_model API): build(), forward(sample_list) with the reference's key-prefix grouping
    (univl_model.py:35-50), get_optimizer_parameters with the reference's four groups (univl_video_ret.py:478-537)."""
    import types

    from b200mm import vtp
    from b200mm.registry import registry

    assert registry.get_model_class("b200_univl") is vtp.B200Univl
    vtp.install_as_univl()
    assert registry.get_model_class("univl") is vtp.B200Univl
    cfg = dict(vtp_common.make_config(hard=True), training_head_type="video_text_retrieval", encoder_lr_decay=0.01)
    torch.manual_seed(0)
    m = registry.get_model_class("univl")(cfg)
    m.build()
    vtp_common.randomize(m.model)
    m = m.to(BF).train()
    img_input, caption = vtp_common.make_batch()
    sample_list = {**img_input, **caption}
    with emulated_ops.patched():
        out = m(sample_list)
        direct = m.model(img_input, caption)
    assert set(out["losses"]) == {"level1_similarity_loss", "level2_similarity_loss"}
    assert torch.equal(out["l2_simi"], direct["l2_simi"])
    opt_cfg = types.SimpleNamespace(optimizer_attributes=types.SimpleNamespace(params=types.SimpleNamespace(lr=1e-4, weight_decay=0.05)))
    groups = m.get_optimizer_parameters(opt_cfg)
    assert len(groups) == 4 and groups[0]["lr"] == groups[2]["lr"] == 1e-4 * 0.01 and "lr" not in groups[1] and groups[2]["weight_decay"] == 0.0
    ids = [id(p) for g in groups for p in g["params"]]
    assert len(ids) == len(set(ids)) == len(list(m.model.parameters()))
    named = dict(m.model.named_parameters())
    assert any(p is named["similarity_dense.0.weight"] for p in groups[1]["params"])
    assert any(p is named["similarity_dense.0.bias"] for p in groups[3]["params"])
    assert any(p is named["module.img_encoder.visual.conv1.weight"] for p in groups[0]["params"])
    with pytest.raises(NotImplementedError):
        bad = vtp.B200Univl(dict(cfg, training_head_type="pretraining"))
        bad.build()


def test_vtp_moco_stage1_host_logic():
    """with_moco (the reference default): momentum key encoders (fp32 copies, EMA), both queue losses, enqueue after the loss and BEFORE
    backward (the order forward_stage1 uses: the saved queue image must not be modified in place) — vs the oracle's moco_nce."""
    from b200mm import vtp

    cfg = vtp_common.make_config(stage="stage1", hard=False, with_moco=True)
    torch.manual_seed(0)
    model = vtp.B200VideoTextRetrieval(cfg)
    vtp_common.randomize(model)
    model = model.to(BF).train()
    img_input, caption = vtp_common.make_batch()
    with emulated_ops.patched():
        out = model(img_input, caption)
        mu = model.moco_utils
        assert int(mu.txt_queue_ptr) == 6 and int(mu.img_queue_ptr) == 6 and set(dict(mu.named_buffers())) >= {"txt_queue", "img_queue"}
        loss = out["losses"]["level1_similarity_loss"]
        loss.backward()  # must not trip autograd's version check on the queue image
        with torch.no_grad():
            cap_input, vis_input, _, _ = model.module.get_l2_input(img_input, caption)
            q_t, q_v = cap_input[2].float(), vis_input[2].float()
            key_v = model.module.forward_img_encoder(**img_input, img_encoder=mu.img_encoder_k)["clip_feature"].float()
            key_t = model.module.forward_text_encoder(caption["caption_raw_input_ids"], caption["caption_input_mask"], txt_encoder=mu.txt_encoder_k)["pooled_output"].float()
    # the first 6 queue columns were overwritten by the enqueue AFTER the loss was taken: compare on the untouched remainder (2 % of the
    # 256 / 16384 negatives)
    ref_v = restated.moco_nce((q_v * key_t).sum(-1, keepdim=True), q_v @ mu.txt_queue.float()[:, 6:].to(BF).float(), 0.05)
    ref_t = restated.moco_nce((q_t * key_v).sum(-1, keepdim=True), q_t @ mu.img_queue.float()[:, 6:].to(BF).float(), 0.05)
    ref = float((ref_v + ref_t) / 2)
    assert abs(float(loss) - ref) < 0.05 * ref, (float(loss), ref)
    assert model.module.img_encoder.visual.proj.grad is not None and model.module.text_encoder.text_projection.grad is not None
    assert all(p.grad is None for p in mu.txt_encoder_k.parameters())
    # EMA: the key encoders moved towards the query encoders by (1 - M)
    w_q = model.module.text_encoder.text_projection.detach().float()
    w_k = mu.txt_encoder_k.text_projection.detach().float()
    assert w_k.dtype == torch.float32 and float((w_k - w_q).abs().max()) < 1e-2
