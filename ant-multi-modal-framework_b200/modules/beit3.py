"""B200-native M²-Encoder (BEiT-3 multiway transformer) — drop-in for the ITC path of prj/M2_Encoder.

Mirrors, with the same attribute / state-dict names, so that reference checkpoints load by key:
  BEiT3                 vlmo/torchscale/model/BEiT3.py:15-96            text_embed, vision_embed.{proj,mask_token,cls_token}, encoder
  Encoder / EncoderLayer vlmo/torchscale/architecture/encoder.py:29-482  layers.{i}.{self_attn,self_attn_layer_norm,ffn,final_layer_norm}, layer_norm,
                                                                         embed_positions.{A,B}
  MultiheadAttention    vlmo/torchscale/component/multihead_attention.py {q,k,v,out}_proj.{A,B}, inner_attn_ln.{A,B}
  FeedForwardNetwork    vlmo/torchscale/component/feedforward_network.py ffn.{A,B}.{fc1,fc2,ffn_layernorm}
  MultiwayNetwork       vlmo/torchscale/component/multiway_network.py    .A (vision expert) / .B (language expert), split_position −1 / 0
  VLMo (ITC part)       vlmo/modules/vlmo_module.py:131-405              backbone, backbone_vl, itc_*_proj.fc, logit_scale, logit_vl_scale,
                                                                         infer_image(batch) / infer_text(batch) with the reference dict keys

The nn.Linear / nn.LayerNorm / nn.Embedding objects below are PARAMETER CONTAINERS with the reference names; their own
forward is never called — every layer runs through b200mm.functional (hand-written CUDA behind the C-ABI).
Supported configuration = what the shipped configs use (configs/Encoder_0.4B.json, Encoder_1B.json): pre-LN + sub-LN,
no deepnorm / MoE / relative position bias, dropout and drop-path 0, head_dim 64; XPOS (args.xpos_rel_pos, off in the shipped configs) is
supported as an in-place rotary step on the fused QKV output. A multiway split inside one sequence (fused vision+language input,
split_position > 0; not used by the ITC path) runs the per-token parts per expert on two token matrices and attention on the joint
sequence (`forward_tokens_mixed`).
"""
import math

import numpy as np
import torch
from torch import nn

from .. import functional as Fn
from ..contrastive import clip_contrastive_loss

BF16 = torch.bfloat16
MASK_BIAS = -30000.0  # additive key bias standing for masked_fill(-inf) (multihead_attention.py:129-135): exp() of it is exactly 0 in fp32


def _bf16(t):
    return t if t.dtype == BF16 else t.to(BF16)


def xpos_tables(L, head_dim, scale_base=512, device="cpu"):
    """(q_cos, q_sin, k_cos, k_sin), f32 [L, head_dim/2]: the tables XPOS.forward builds for offset 0 — q with `scale`, k with 1/scale
    (downscale=True) — written with the reference's own expressions (xpos_relative_position.py:9-13, :41-61) so the values are identical."""
    base = (torch.arange(0, head_dim, 2) + 0.4 * head_dim) / (1.4 * head_dim)
    min_pos = -(L + 0) // 2
    max_pos = L + 0 + min_pos
    scale = base ** torch.arange(min_pos, max_pos, 1).to(base).div(scale_base)[:, None]
    seq_len, dim = scale.shape
    inv_freq = 1.0 / (10000 ** (torch.arange(0, dim) / dim))
    sinusoid = torch.einsum("i , j -> i j", torch.arange(0, seq_len, dtype=torch.float), inv_freq).to(scale)
    sin, cos = torch.sin(sinusoid), torch.cos(sinusoid)
    out = (cos * scale, sin * scale, cos * (1 / scale), sin * (1 / scale))
    return tuple(t.float().contiguous().to(device) for t in out)


class XPOS(nn.Module):
    """State-dict twin of the reference XPOS module (xpos_relative_position.py:35-40): the registered buffer `scale` [head_dim/2]."""

    def __init__(self, head_dim, scale_base=512):
        super().__init__()
        self.head_dim, self.scale_base = head_dim, scale_base
        self.register_buffer("scale", (torch.arange(0, head_dim, 2) + 0.4 * head_dim) / (1.4 * head_dim))


class MultiwayNetwork(nn.Module):
    """Two copies of a parameter container: .A (vision expert) and .B (language expert) — multiway_network.py:24-45."""

    def __init__(self, make):
        super().__init__()
        self.A = make()
        self.B = make()
        self.split_position = -1

    def way(self, split_position):
        """Expert for a whole-sequence call: -1 -> A (vision), 0 -> B (language). A split INSIDE the sequence (> 0) is handled by the
        callers' *_mixed paths, which ask for both experts."""
        if split_position == -1:
            return self.A
        if split_position == 0:
            return self.B
        raise ValueError(f"MultiwayNetwork.way: split_position {split_position} selects no single expert")


class FeedForwardNetwork(nn.Module):
    def __init__(self, embed_dim, ffn_dim, eps):
        super().__init__()
        self.fc1 = nn.Linear(embed_dim, ffn_dim)
        self.fc2 = nn.Linear(ffn_dim, embed_dim)
        self.ffn_layernorm = nn.LayerNorm(ffn_dim, eps=eps)


class MultiheadAttention(nn.Module):
    def __init__(self, embed_dim, num_heads, eps, xpos_rel_pos=False, xpos_scale_base=512):
        super().__init__()
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.xpos_rel_pos, self.xpos_scale_base = xpos_rel_pos, xpos_scale_base
        self._xpos_cache = {}
        if xpos_rel_pos:
            self.xpos = XPOS(embed_dim // num_heads, xpos_scale_base)
        self.head_dim = embed_dim // num_heads
        self.scaling = self.head_dim ** -0.5
        lin = lambda: nn.Linear(embed_dim, embed_dim, bias=True)  # noqa: E731
        self.k_proj = MultiwayNetwork(lin)
        self.v_proj = MultiwayNetwork(lin)
        self.q_proj = MultiwayNetwork(lin)
        self.out_proj = MultiwayNetwork(lin)
        self.inner_attn_ln = MultiwayNetwork(lambda: nn.LayerNorm(embed_dim, eps=eps))

    def xpos_tables_for(self, L, device):
        """XPOS tables for sequences of length L (None when args.xpos_rel_pos is off, the shipped default)."""
        if not self.xpos_rel_pos:
            return None
        key = (L, str(device))
        if key not in self._xpos_cache:
            self._xpos_cache[key] = xpos_tables(L, self.head_dim, self.xpos_scale_base, device)
        return self._xpos_cache[key]


class EncoderLayer(nn.Module):
    """architecture/encoder.py:29-168 with encoder_normalize_before=True, subln=True, alpha=1."""

    def __init__(self, embed_dim, num_heads, ffn_dim, eps, xpos_rel_pos=False, xpos_scale_base=512):
        super().__init__()
        self.embed_dim, self.eps = embed_dim, eps
        self.self_attn = MultiheadAttention(embed_dim, num_heads, eps, xpos_rel_pos, xpos_scale_base)
        self.self_attn_layer_norm = MultiwayNetwork(lambda: nn.LayerNorm(embed_dim, eps=eps))
        self.ffn = MultiwayNetwork(lambda: FeedForwardNetwork(embed_dim, ffn_dim, eps))
        self.final_layer_norm = MultiwayNetwork(lambda: nn.LayerNorm(embed_dim, eps=eps))
        self.checkpoint = False
        self.keep_act = False

    def layer_params(self, sp):
        a = self.self_attn
        ln1, ln2, ffn, iln = self.self_attn_layer_norm.way(sp), self.final_layer_norm.way(sp), self.ffn.way(sp), a.inner_attn_ln.way(sp)
        q, k, v, o = a.q_proj.way(sp), a.k_proj.way(sp), a.v_proj.way(sp), a.out_proj.way(sp)
        return tuple(_bf16(t) for t in (ln1.weight, ln1.bias, q.weight, q.bias, k.weight, k.bias, v.weight, v.bias, iln.weight, iln.bias,
                                        o.weight, o.bias, ln2.weight, ln2.bias, ffn.fc1.weight, ffn.fc1.bias, ffn.ffn_layernorm.weight,
                                        ffn.ffn_layernorm.bias, ffn.fc2.weight, ffn.fc2.bias))

    def forward_tokens(self, x2d, key_bias, B, L, split_position):
        return Fn.M2EncoderLayerFn.apply(x2d, *self.layer_params(split_position), key_bias, B, L, self.self_attn.num_heads, self.eps,
                                         self.checkpoint, self.keep_act, self.self_attn.xpos_tables_for(L, x2d.device))


    def _pre_post_params(self, sp):
        a = self.self_attn
        ln1, ln2, ffn, iln = self.self_attn_layer_norm.way(sp), self.final_layer_norm.way(sp), self.ffn.way(sp), a.inner_attn_ln.way(sp)
        q, k, v, o = a.q_proj.way(sp), a.k_proj.way(sp), a.v_proj.way(sp), a.out_proj.way(sp)
        pre = tuple(_bf16(t) for t in (ln1.weight, ln1.bias, q.weight, q.bias, k.weight, k.bias, v.weight, v.bias))
        post = tuple(_bf16(t) for t in (iln.weight, iln.bias, o.weight, o.bias, ln2.weight, ln2.bias, ffn.fc1.weight, ffn.fc1.bias,
                                        ffn.ffn_layernorm.weight, ffn.ffn_layernorm.bias, ffn.fc2.weight, ffn.fc2.bias))
        return pre, post

    def forward_tokens_mixed(self, xA, xB, idx, key_bias, B, L):
        """Multiway split inside the sequence (multiway_network.py:38-45): rows of expert A (vision) and B (language) are kept in two
        token matrices; only attention sees the joint [B*L] order (`idx`: merge / takeA / takeB row-id lists)."""
        H = self.self_attn.num_heads
        preA, postA = self._pre_post_params(-1)
        preB, postB = self._pre_post_params(0)
        qa = Fn.M2PreAttnFn.apply(xA, *preA, self.eps)
        qb = Fn.M2PreAttnFn.apply(xB, *preB, self.eps)
        qkv = Fn.GatherRowsFn.apply(torch.cat([qa, qb], dim=0), idx["merge"])
        xp = self.self_attn.xpos_tables_for(L, qkv.device)
        if xp is not None:
            qkv = Fn.XposFn.apply(qkv, xp, B, L, H)
        a = Fn.AttentionFn.apply(qkv, key_bias, B, L, H)
        aA = Fn.GatherRowsFn.apply(a, idx["takeA"])
        aB = Fn.GatherRowsFn.apply(a, idx["takeB"])
        return Fn.M2PostAttnFn.apply(aA, xA, *postA, self.eps), Fn.M2PostAttnFn.apply(aB, xB, *postB, self.eps)


def split_index(B, L, s, device):
    """Row-id lists between the joint token order [B*L] and the two per-expert matrices A [B*s] (tokens < s) and B [B*(L-s)]."""
    b = torch.arange(B, device=device)[:, None]
    l = torch.arange(L, device=device)[None, :]
    merge = torch.where(l < s, b * s + l, B * s + b * (L - s) + (l - s)).reshape(-1).contiguous()   # into cat([A, B]) rows
    takeA = (b * L + torch.arange(s, device=device)[None, :]).reshape(-1).contiguous()
    takeB = (b * L + s + torch.arange(L - s, device=device)[None, :]).reshape(-1).contiguous()
    return {"merge": merge, "takeA": takeA, "takeB": takeB}


class PositionalEmbedding(nn.Embedding):
    """component/embedding.py:93-110: positions start at 2 (fairseq convention)."""


class _MultiwayEmbedding(nn.Module):
    def __init__(self, a, b):
        super().__init__()
        self.A, self.B = a, b


class Encoder(nn.Module):
    """architecture/encoder.py:171-482 for the configuration named in the module docstring. Works on token matrices
    [B*L, W] (batch-first rows); `forward` keeps the reference's keyword interface and returns the same dict keys."""

    def __init__(self, embed_dim=768, attention_heads=12, ffn_dim=3072, layers=12, eps=1e-5, embed_positions=None, xpos_rel_pos=False,
                 xpos_scale_base=512):
        super().__init__()
        self.embed_dim, self.eps = embed_dim, eps
        self.embed_positions = embed_positions
        self.layers = nn.ModuleList([EncoderLayer(embed_dim, attention_heads, ffn_dim, eps, xpos_rel_pos, xpos_scale_base) for _ in range(layers)])
        self.num_layers = layers
        self.layer_norm = MultiwayNetwork(lambda: nn.LayerNorm(embed_dim, eps=eps))
        # subln init (encoder.py:262-269): fc1 / fc2 / out_proj / v_proj scaled by sqrt(log(2·layers))
        init_scale = math.sqrt(math.log(layers * 2))
        with torch.no_grad():
            for name, p in self.named_parameters():
                if p.dim() == 2 and ("fc1" in name or "fc2" in name or "out_proj" in name or "v_proj" in name):
                    p.normal_(0.0, 0.02).mul_(init_scale)
                elif p.dim() == 2 and "proj" in name:
                    p.normal_(0.0, 0.02)
                elif name.endswith("bias"):
                    p.zero_()

    def set_grad_checkpointing(self, enable=True):
        for layer in self.layers:
            layer.checkpoint = bool(enable)

    def set_keep_activation(self, n_layers):
        """Keep the sub-LN-normalised FFN hidden (4·W per token) of the first `n_layers` layers for backward instead of
        recomputing gelu + sub-LN from the pre-activation (memory for time)."""
        for i, layer in enumerate(self.layers):
            layer.keep_act = i < n_layers

    def forward_tokens(self, x2d, B, L, split_position, drop=None, key_bias=None, mask_input=True):
        """x2d: [B*L, W] token embeddings with positions added; drop: uint8 [B*L] (1 = padded row) or None."""
        if drop is not None and mask_input:
            x2d = Fn.MaskRowsFn.apply(x2d, drop)  # encoder.py:440
        for layer in self.layers:
            x2d = layer.forward_tokens(x2d, key_bias, B, L, split_position)
        ln = self.layer_norm.way(split_position)
        return Fn.LayerNormFn.apply(x2d, _bf16(ln.weight), _bf16(ln.bias), self.eps)

    def forward_tokens_mixed(self, xA, xB, B, L, s, drop=None, key_bias=None, mask_input=True):
        """Fused vision + language input: xA [B*s, W] (expert A rows), xB [B*(L-s), W]; drop / key_bias in the JOINT order ([B*L] / [B, L]).
        Returns the joint [B*L, W] hidden after the per-expert final layer_norm."""
        idx = split_index(B, L, s, xA.device)
        if drop is not None and mask_input:
            xA = Fn.MaskRowsFn.apply(xA, drop[idx["takeA"]].contiguous())
            xB = Fn.MaskRowsFn.apply(xB, drop[idx["takeB"]].contiguous())
        for layer in self.layers:
            xA, xB = layer.forward_tokens_mixed(xA, xB, idx, key_bias, B, L)
        lnA, lnB = self.layer_norm.way(-1), self.layer_norm.way(0)
        xA = Fn.LayerNormFn.apply(xA, _bf16(lnA.weight), _bf16(lnA.bias), self.eps)
        xB = Fn.LayerNormFn.apply(xB, _bf16(lnB.weight), _bf16(lnB.bias), self.eps)
        return Fn.GatherRowsFn.apply(torch.cat([xA, xB], dim=0), idx["merge"])

    def forward(self, src_tokens=None, encoder_padding_mask=None, attn_mask=None, return_all_hiddens=False, token_embeddings=None,
                multiway_split_position=None, features_only=False, incremental_state=None, positions=None, **kwargs):
        if src_tokens is not None or attn_mask is not None or incremental_state is not None or positions is not None or return_all_hiddens:
            raise NotImplementedError("b200mm M2 Encoder.forward: only token_embeddings (+ encoder_padding_mask, multiway_split_position)")
        if self.embed_positions is not None:
            raise NotImplementedError("b200mm M2 Encoder.forward: an encoder with positional embeddings is driven through BEiT3.forward")
        B, L, W = token_embeddings.shape
        sp = -1 if multiway_split_position is None else multiway_split_position
        drop, key_bias = _padding(encoder_padding_mask)
        x2d = _bf16(token_embeddings).reshape(B * L, W).contiguous()
        if 0 < sp < L:
            idx = split_index(B, L, sp, x2d.device)
            out = self.forward_tokens_mixed(Fn.GatherRowsFn.apply(x2d, idx["takeA"]), Fn.GatherRowsFn.apply(x2d, idx["takeB"]), B, L, sp, drop, key_bias)
        else:
            out = self.forward_tokens(x2d, B, L, sp if sp <= 0 else 0, drop, key_bias)
        return {"encoder_out": out.view(B, L, W), "encoder_embedding": token_embeddings, "encoder_padding_mask": encoder_padding_mask,
                "encoder_states": [], "l_aux": [None] * self.num_layers, "multiway_split_position": multiway_split_position}


def _padding(mask):
    """[B, L] padding mask (non-zero = padded) -> (uint8 [B*L] row flags, f32 [B, L] additive key bias), or (None, None)."""
    if mask is None:
        return None, None
    pad = mask.to(torch.bool)
    drop = pad.reshape(-1).to(torch.uint8).contiguous()
    key_bias = torch.zeros(pad.shape, device=pad.device, dtype=torch.float32).masked_fill_(pad, MASK_BIAS)
    return drop, key_bias


class VisionEmbedding(nn.Module):
    """component/embedding.py:27-83 (parameter container; the arithmetic is M2VisionEmbedFn)."""

    def __init__(self, img_size, patch_size, in_chans, embed_dim):
        super().__init__()
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))

    def num_position_embeddings(self):
        return self.num_patches + 1


class BEiT3(nn.Module):
    """model/BEiT3.py:15-96. forward(textual_tokens=…, text_padding_position=…) or forward(visual_tokens=…)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, vocab_size=64010, encoder_embed_dim=768, encoder_attention_heads=12,
                 encoder_ffn_embed_dim=3072, encoder_layers=12, max_source_positions=1024, layernorm_eps=1e-5, xpos_rel_pos=False,
                 xpos_scale_base=512):
        super().__init__()
        W = encoder_embed_dim
        self.text_embed = nn.Embedding(vocab_size, W)
        nn.init.normal_(self.text_embed.weight, mean=0, std=W ** -0.5)
        self.vision_embed = VisionEmbedding(img_size, patch_size, in_chans, W)
        embed_positions = _MultiwayEmbedding(PositionalEmbedding(self.vision_embed.num_position_embeddings() + 2, W),
                                             PositionalEmbedding(max_source_positions, W))
        self.encoder = Encoder(W, encoder_attention_heads, encoder_ffn_embed_dim, encoder_layers, layernorm_eps, embed_positions, xpos_rel_pos,
                               xpos_scale_base)

    def forward_tokens(self, textual_tokens=None, visual_tokens=None, text_padding_position=None):
        """-> ([B*L, W] hidden after the final layer_norm, B, L, drop, key_bias)"""
        enc = self.encoder
        if textual_tokens is None:
            ve = self.vision_embed
            B = visual_tokens.shape[0]
            L = ve.num_position_embeddings()
            x = Fn.M2VisionEmbedFn.apply(_bf16(visual_tokens).contiguous(), _bf16(ve.proj.weight), _bf16(ve.proj.bias), _bf16(ve.cls_token),
                                         _bf16(enc.embed_positions.A.weight[2: L + 2]))
            return enc.forward_tokens(x, B, L, -1), B, L, None, None
        B, Lt = textual_tokens.shape
        drop_t, bias_t = _padding(text_padding_position)
        xt = Fn.M2TextEmbedFn.apply(_bf16(self.text_embed.weight), textual_tokens, _bf16(enc.embed_positions.B.weight[2: Lt + 2]), drop_t)
        if visual_tokens is None:
            return enc.forward_tokens(xt, B, Lt, 0, drop_t, bias_t, mask_input=False), B, Lt, drop_t, bias_t
        # fused vision + language input (model/BEiT3.py:68-86): [vision tokens ; text tokens], split position = number of vision tokens
        if visual_tokens.shape[0] != B:
            raise NotImplementedError("b200mm BEiT3: fused input with repeated text (image batch a multiple of the text batch, BEiT3.py:71-74)")
        ve = self.vision_embed
        Lv = ve.num_position_embeddings()
        xv = Fn.M2VisionEmbedFn.apply(_bf16(visual_tokens).contiguous(), _bf16(ve.proj.weight), _bf16(ve.proj.bias), _bf16(ve.cls_token),
                                      _bf16(enc.embed_positions.A.weight[2: Lv + 2]))
        L = Lv + Lt
        drop = key_bias = None
        if text_padding_position is not None:
            pad = torch.cat([torch.zeros((B, Lv), dtype=torch.bool, device=textual_tokens.device), text_padding_position.to(torch.bool)], dim=1)
            drop, key_bias = _padding(pad)
        # the embeddings already zeroed the padded text rows (architecture/encoder.py:440); vision rows are never padded
        return enc.forward_tokens_mixed(xv, xt, B, L, Lv, drop, key_bias, mask_input=False), B, L, drop, key_bias

    def forward(self, textual_tokens=None, visual_tokens=None, text_padding_position=None, attn_mask=None, vision_masked_position=None,
                incremental_state=None, positions=None):
        if attn_mask is not None or vision_masked_position is not None or incremental_state is not None or positions is not None:
            raise NotImplementedError("b200mm BEiT3.forward: attn_mask / masked positions / incremental decoding are not on the ITC path")
        h, B, L, _, _ = self.forward_tokens(textual_tokens, visual_tokens, text_padding_position)
        split = -1 if textual_tokens is None else (0 if visual_tokens is None else self.vision_embed.num_position_embeddings())
        return {"encoder_out": h.view(B, L, -1), "encoder_padding_mask": text_padding_position, "multiway_split_position": split}


class ITCHead(nn.Module):
    """vlmo/modules/heads.py:17-24."""

    def __init__(self, hidden_size, out_size):
        super().__init__()
        self.fc = nn.Linear(hidden_size, out_size, bias=False)
        nn.init.normal_(self.fc.weight, mean=0.0, std=0.02)


class Pooler(nn.Module):
    """vlmo/modules/heads.py:4-14 — parameters only (state-dict compatibility; the ITC path does not call it)."""

    def __init__(self, hidden_size):
        super().__init__()
        self.dense = nn.Linear(hidden_size, hidden_size)


M2_CONFIGS = {
    # configs/Encoder_0.4B.json + vlmo/config.py defaults (beit_version "base"), configs/Encoder_1B.json ("large")
    "M2-Encoder-0.4B": dict(image_size=224, patch_size=16, vocab_size=115244, encoder_embed_dim=768, encoder_attention_heads=12,
                            encoder_layers=9, beit3_vl_layers=3, out_embed_dim=768, max_text_len=52),
    "M2-Encoder-1B": dict(image_size=224, patch_size=16, vocab_size=115244, encoder_embed_dim=1024, encoder_attention_heads=16,
                          encoder_layers=21, beit3_vl_layers=3, out_embed_dim=1024, max_text_len=52),
}


class M2Encoder(nn.Module):
    """ITC part of VLMo (vlmo/modules/vlmo_module.py:131-405): `infer_image` / `infer_text` take and return the reference's
    batch dicts; `itc_loss` is the symmetric InfoNCE over both head pairs on the fused similarity / log-softmax kernels."""

    def __init__(self, image_size=224, patch_size=16, vocab_size=115244, encoder_embed_dim=768, encoder_attention_heads=12, encoder_layers=9,
                 beit3_vl_layers=3, out_embed_dim=768, max_text_len=52, mlp_ratio=4, max_source_positions=1024, xpos_rel_pos=False,
                 xpos_scale_base=512):
        super().__init__()
        W = encoder_embed_dim
        self.img_size, self.num_features, self.out_features, self.max_text_len = image_size, W, out_embed_dim, max_text_len
        self.backbone = BEiT3(image_size, patch_size, 3, vocab_size, W, encoder_attention_heads, int(W * mlp_ratio), encoder_layers,
                              max_source_positions, xpos_rel_pos=xpos_rel_pos, xpos_scale_base=xpos_scale_base)
        self.use_vl = beit3_vl_layers > 0
        if self.use_vl:
            self.backbone_vl = Encoder(W, encoder_attention_heads, int(W * mlp_ratio), beit3_vl_layers, xpos_rel_pos=xpos_rel_pos,
                                       xpos_scale_base=xpos_scale_base)
        self.norm = nn.LayerNorm(W, eps=1e-6)  # present in the reference state dict, unused on the ITC path (:176)
        self.pooler = Pooler(W)
        self.itc_text_proj = ITCHead(W, out_embed_dim)
        self.itc_image_proj = ITCHead(W, out_embed_dim)
        self.itc_vl_text_proj = ITCHead(W, out_embed_dim)
        self.itc_vl_image_proj = ITCHead(W, out_embed_dim)
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        self.logit_vl_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))

    def set_grad_checkpointing(self, enable=True):
        self.backbone.encoder.set_grad_checkpointing(enable)
        if self.use_vl:
            self.backbone_vl.set_grad_checkpointing(enable)

    def set_keep_activation(self, n_layers):
        n0 = len(self.backbone.encoder.layers)
        self.backbone.encoder.set_keep_activation(n_layers)
        if self.use_vl:
            self.backbone_vl.set_keep_activation(max(0, n_layers - n0))

    def _heads(self, h, hv, B, L, proj, proj_vl):
        f = Fn.RowNormFn.apply(Fn.ClsLinearFn.apply(h, _bf16(proj.fc.weight), B, L))
        fv = Fn.RowNormFn.apply(Fn.ClsLinearFn.apply(hv, _bf16(proj_vl.fc.weight), B, L))
        return f, fv

    @staticmethod
    def img_norm(img):
        """inception_normalize (vlmo/transforms/utils.py:48, applied inside VLMo.infer_image, vlmo_module.py:17,385): (x - 0.5) / 0.5."""
        return ((img.float() - 0.5) / 0.5).to(img.dtype)

    def infer_image(self, batch, mask_image=False, image_token_type_idx=1, image_embeds=None, image_masks=None):
        """vlmo_module.py:364-405, same batch dictionary as the reference: batch["image"][0] is the loader output in [0, 1]; the inception
        normalisation of :385 is applied here (one elementwise pass over the pixels), so the method is a drop-in."""
        if mask_image:
            raise NotImplementedError("b200mm M2Encoder.infer_image: masked image modelling is not on the ITC path")
        imgkey = f"image_{image_token_type_idx - 1}" if f"image_{image_token_type_idx - 1}" in batch else "image"
        img = self.img_norm(batch[imgkey][0])
        h, B, L, _, _ = self.backbone.forward_tokens(visual_tokens=img)
        hv = self.backbone_vl.forward_tokens(h, B, L, -1)
        f, fv = self._heads(h, hv, B, L, self.itc_image_proj, self.itc_vl_image_proj)
        return {"image_feats": h.view(B, L, -1), "cls_feats": f, "cls_vlffn_feats": fv}

    def infer_text(self, batch, mask_text=False, with_text_embed=True):
        """vlmo_module.py:323-362; returns the reference's keys (cls_feats, cls_vlffn_feats, text_embed = backbone.text_embed(text_ids),
        :332) plus the language hiddens as "text_hidden". with_text_embed=False skips the embedding lookup nobody on the ITC path reads."""
        do_mlm = "_mlm" if mask_text else ""
        text_ids = batch[f"text_ids{do_mlm}"]
        text_padding_position = 1 - batch["text_masks"]
        h, B, L, drop, key_bias = self.backbone.forward_tokens(textual_tokens=text_ids, text_padding_position=text_padding_position)
        hv = self.backbone_vl.forward_tokens(h, B, L, -1, drop, key_bias)  # expert A on the language hiddens (:343)
        f, fv = self._heads(h, hv, B, L, self.itc_text_proj, self.itc_vl_text_proj)
        ret = {"cls_feats": f, "cls_vlffn_feats": fv, "text_hidden": h.view(B, L, -1)}
        if with_text_embed:
            ret["text_embed"] = self.backbone.text_embed(text_ids)
        return ret

    def itc_loss(self, image, text_ids, text_masks, group=None):
        """Symmetric InfoNCE on (cls_feats, logit_scale) and (cls_vlffn_feats, logit_vl_scale); similarity as m2_encoder.py:92-95."""
        i = self.infer_image({"image": [image]})
        t = self.infer_text({"text_ids": text_ids, "text_masks": text_masks}, with_text_embed=False)
        return (clip_contrastive_loss(i["cls_feats"], t["cls_feats"], self.logit_scale, group)
                + clip_contrastive_loss(i["cls_vlffn_feats"], t["cls_vlffn_feats"], self.logit_vl_scale, group))
