"""B200-native VisionTransformer — drop-in for antmmf/modules/vision/backbone/clip/model.py:275-335.

Same constructor arguments, forward signature, initialisation and state-dict keys
(conv1.weight, class_embedding, positional_embedding, ln_pre.*, transformer.resblocks.{i}.{ln_1,attn,mlp.c_fc,
mlp.c_proj,ln_2}.*, ln_post.*, proj) as the reference, so reference checkpoints load by key
(clip_visual_encoder.py:46-71). The arithmetic runs in the b200mm CUDA kernels through b200mm.functional.
"""
from collections import OrderedDict

import torch
from torch import nn

from .. import functional as Fn

BF16 = torch.bfloat16


def _bf16(t):
    """Kernels take bf16; fp32 master parameters are cast per call (autograd casts the gradient back)."""
    return t if t.dtype == BF16 else t.to(BF16)


class _MHAParams(nn.Module):
    """Parameter container with nn.MultiheadAttention's key names (in_proj_weight, in_proj_bias, out_proj.*)."""

    def __init__(self, d_model):
        super().__init__()
        ref = nn.MultiheadAttention(d_model, 1)  # reference init (xavier in_proj, zero biases), clip/model.py:231
        self.in_proj_weight = nn.Parameter(ref.in_proj_weight.detach().clone())
        self.in_proj_bias = nn.Parameter(ref.in_proj_bias.detach().clone())
        self.out_proj = nn.Linear(d_model, d_model)
        with torch.no_grad():
            self.out_proj.weight.copy_(ref.out_proj.weight)
            self.out_proj.bias.copy_(ref.out_proj.bias)


class ResidualAttentionBlock(nn.Module):
    """clip/model.py:227-256 (pre-LN, QuickGELU MLP). attn_mask is not supported (the ViT never passes one)."""

    def __init__(self, d_model: int, n_head: int, attn_mask=None):
        super().__init__()
        if attn_mask is not None:
            raise NotImplementedError("b200mm ResidualAttentionBlock: attn_mask is only used by the CLIP text tower, not on this path")
        self.n_head = n_head
        self.attn = _MHAParams(d_model)
        self.ln_1 = nn.LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)), ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = nn.LayerNorm(d_model)
        self.checkpoint = False
        self.keep_act = False
        self.keep_ln = False

    def block_params(self):
        return tuple(_bf16(p) for p in (self.ln_1.weight, self.ln_1.bias, self.attn.in_proj_weight, self.attn.in_proj_bias,
                                        self.attn.out_proj.weight, self.attn.out_proj.bias, self.ln_2.weight, self.ln_2.bias,
                                        self.mlp.c_fc.weight, self.mlp.c_fc.bias, self.mlp.c_proj.weight, self.mlp.c_proj.bias))

    def forward_tokens(self, x2d, B, L):
        return Fn.VitBlockFn.apply(x2d, *self.block_params(), B, L, self.n_head, self.ln_1.eps, self.checkpoint, self.keep_act, self.keep_ln)


class Transformer(nn.Module):
    def __init__(self, width: int, layers: int, heads: int, attn_mask=None):
        super().__init__()
        self.width = width
        self.layers = layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask) for _ in range(layers)])


class VisionTransformer(nn.Module):
    def __init__(self, input_resolution: int, patch_size: int, width: int, layers: int, heads: int, output_dim: int):
        super().__init__()
        self.input_resolution = input_resolution
        self.output_dim = output_dim
        self.conv1 = nn.Conv2d(in_channels=3, out_channels=width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))

    def set_grad_checkpointing(self, enable=True, every=1):
        """Keep only each block's input and re-run its forward in backward (for `every`-th blocks)."""
        for i, blk in enumerate(self.transformer.resblocks):
            blk.checkpoint = bool(enable) and (i % every == 0)

    def set_keep_activation(self, n_blocks):
        """Keep the activated MLP hidden (4*width per token) of the first `n_blocks` blocks for backward instead of recomputing
        it from the pre-activation (2 bytes * 4 * width per token and block of extra memory; saves one HBM pass per block)."""
        for i, blk in enumerate(self.transformer.resblocks):
            blk.keep_act = i < n_blocks

    def set_keep_layernorm(self, n_blocks):
        """Keep both LayerNorm outputs (2 x width per token, 4*width bytes) of the first `n_blocks` blocks for backward instead of
        recomputing them: half the memory of `set_keep_activation` per block and more time saved (two HBM passes per block)."""
        for i, blk in enumerate(self.transformer.resblocks):
            blk.keep_ln = i < n_blocks

    def forward_features(self, x: torch.Tensor):
        """[B, 3, R, R] -> token matrix [B*L, width] after the last block (no ln_post)."""
        B = x.shape[0]
        x = _bf16(x).contiguous()
        t = Fn.VitStemFn.apply(x, _bf16(self.conv1.weight), _bf16(self.class_embedding), _bf16(self.positional_embedding),
                               _bf16(self.ln_pre.weight), _bf16(self.ln_pre.bias), self.ln_pre.eps)
        L = self.positional_embedding.shape[0]
        for blk in self.transformer.resblocks:
            t = blk.forward_tokens(t, B, L)
        return t, B, L

    def forward(self, x: torch.Tensor):
        t, B, L = self.forward_features(x)
        if self.proj is None:
            raise NotImplementedError("b200mm VisionTransformer: proj=None is not used on this path")
        return Fn.ClsHeadFn.apply(t, _bf16(self.ln_post.weight), _bf16(self.ln_post.bias), _bf16(self.proj), B, L, self.ln_post.eps)
