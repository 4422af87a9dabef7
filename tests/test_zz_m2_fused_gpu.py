"""GPU check of the M²-Encoder multiway split INSIDE a sequence (fused vision + language input; SURVEY.md §8 rows M1-M3, not on the ITC
path): BEiT3.forward(textual_tokens, visual_tokens, text_padding_position) and Encoder.forward(token_embeddings, multiway_split_position=s)
against the oracle's fused forward (pinned to the unmodified reference in tests/test_oracle.py). Written after the round's GPU budget was
spent: the host glue is verified on CPU over emulated kernels (tests/test_m2_host_cpu.py); this file sorts last on purpose."""
import pytest
import torch

from oracle import restated

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rel_l2(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-12))


def test_m2_fused_vision_language_input():
    from b200mm.modules import M2Encoder

    torch.manual_seed(0)
    m = M2Encoder(image_size=32, patch_size=8, vocab_size=64, encoder_embed_dim=64, encoder_attention_heads=2, encoder_layers=1,
                  beit3_vl_layers=1, out_embed_dim=32, max_text_len=8, max_source_positions=16).cuda().to(BF)
    ids = torch.randint(1, 64, (3, 8), device="cuda")
    pad = torch.zeros(3, 8, dtype=torch.long, device="cuda")
    pad[:, 6:] = 1
    # fused vision + language input: multiway split inside the sequence (two per-expert token matrices, joint attention)
    img = torch.randn(3, 3, 32, 32, device="cuda")
    fused = m.backbone(textual_tokens=ids, visual_tokens=img, text_padding_position=pad)
    assert fused["encoder_out"].shape == (3, 17 + 8, 64) and fused["multiway_split_position"] == 17
    assert torch.isfinite(fused["encoder_out"].float()).all()
    sd = {k: v.detach().float().cpu() for k, v in m.state_dict().items()}
    ref = restated.m2_fused_forward(sd, img.to(BF).float().cpu(), ids.cpu(), (1 - pad).cpu(), 2)
    valid = torch.cat([torch.ones(3, 17, dtype=torch.bool), pad.cpu() == 0], 1)
    assert rel_l2(fused["encoder_out"].float().cpu()[valid], ref[valid]) < 5e-2
    mixed = m.backbone_vl(src_tokens=None, token_embeddings=fused["encoder_out"], multiway_split_position=17)
    assert mixed["encoder_out"].shape == (3, 25, 64) and torch.isfinite(mixed["encoder_out"].float()).all()
    fused["encoder_out"].float().square().mean().backward()
    assert m.backbone.encoder.layers[0].ffn.A.fc1.weight.grad is not None and m.backbone.encoder.layers[0].ffn.B.fc1.weight.grad is not None
