"""B200-native BERT encoder — drop-in for antmmf/modules/vision/backbone/clip/modeling_bert.py (Chinese-CLIP BERT:
post-LN, erf-GELU, no pooler) with the same config object, forward signatures, initialisation and state-dict keys:

  embeddings.{word,position,token_type}_embeddings.weight, embeddings.LayerNorm.*,
  encoder.layer.{i}.attention.self.{query,key,value}.*, .attention.output.{dense,LayerNorm}.*,
  .intermediate.dense.*, .output.{dense,LayerNorm}.*

Dropout (modeling_bert.py:84,124,158,180,232): in training mode the three sites per layer (attention probabilities, self-output,
output) and the embedding dropout run inside the kernels — fused into the attention softmax, the two projection-GEMM epilogues and one
in-place pass — with counter-based masks regenerated from per-site 64-bit seeds in backward (`ops.next_dropout_seed`, re-based with
`b200mm.ops.manual_seed`). Eval mode and p = 0 take the dropout-free kernels, bit-identical to before.
"""
import torch
from torch import nn

from .. import functional as Fn
from .vit import _bf16

BF16 = torch.bfloat16


class BertConfig(object):
    """Same fields as clip/configuration_bert.py:56-99 (the subset the encoder reads)."""

    def __init__(self, vocab_size_or_config_json_file=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                 intermediate_size=3072, hidden_act="gelu", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1,
                 max_position_embeddings=512, type_vocab_size=2, initializer_range=0.02, layer_norm_eps=1e-12,
                 output_attentions=False, output_hidden_states=False, **kwargs):
        if not isinstance(vocab_size_or_config_json_file, int):
            raise ValueError("b200mm BertConfig: pass the vocabulary size as an int")
        self.vocab_size = vocab_size_or_config_json_file
        self.hidden_size = hidden_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.intermediate_size = intermediate_size
        self.hidden_act = hidden_act
        self.hidden_dropout_prob = hidden_dropout_prob
        self.attention_probs_dropout_prob = attention_probs_dropout_prob
        self.max_position_embeddings = max_position_embeddings
        self.type_vocab_size = type_vocab_size
        self.initializer_range = initializer_range
        self.layer_norm_eps = layer_norm_eps
        self.output_attentions = output_attentions
        self.output_hidden_states = output_hidden_states
        for k, v in kwargs.items():
            setattr(self, k, v)


class BertEmbeddings(nn.Module):
    """modeling_bert.py:66-103; `inputs_embeds` variant of prj/base_vtp/.../clip_text_encoder.py:36-60."""

    def __init__(self, config):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=0)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, input_ids=None, token_type_ids=None, position_ids=None, inputs_embeds=None):
        if position_ids is not None:
            raise NotImplementedError("b200mm BertEmbeddings: explicit position_ids are not used on this path (arange is fused)")
        if inputs_embeds is None:
            B, L = input_ids.shape
            table = _bf16(self.word_embeddings.weight)
            ids = input_ids.reshape(-1).contiguous()
            from_embeds = False
        else:
            B, L = inputs_embeds.shape[:2]
            table = _bf16(inputs_embeds).reshape(B * L, -1).contiguous()
            ids = torch.arange(B * L, device=table.device)
            from_embeds = True
        if token_type_ids is None:
            token_type_ids = torch.zeros((B, L), dtype=torch.long, device=ids.device)
        drop = (self.dropout.p, Fn.ops.next_dropout_seed()) if self.training and self.dropout.p > 0 else None
        y = Fn.BertEmbeddingsFn.apply(table, ids, _bf16(self.position_embeddings.weight), _bf16(self.token_type_embeddings.weight),
                                      token_type_ids.reshape(-1).contiguous(), _bf16(self.LayerNorm.weight), _bf16(self.LayerNorm.bias),
                                      L, self.LayerNorm.eps, self.word_embeddings.padding_idx, from_embeds, drop)
        return y.view(B, L, -1)


class _SelfAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.query = nn.Linear(config.hidden_size, config.hidden_size)
        self.key = nn.Linear(config.hidden_size, config.hidden_size)
        self.value = nn.Linear(config.hidden_size, config.hidden_size)


class _SelfOutput(nn.Module):
    def __init__(self, config, in_features):
        super().__init__()
        self.dense = nn.Linear(in_features, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class _Attention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.self = _SelfAttention(config)
        self.output = _SelfOutput(config, config.hidden_size)


class _Intermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)


class BertLayer(nn.Module):
    """modeling_bert.py:253-270."""

    def __init__(self, config):
        super().__init__()
        if config.hidden_act != "gelu":
            raise NotImplementedError(f"b200mm BertLayer: hidden_act={config.hidden_act!r}; the path uses erf-GELU ('gelu')")
        self.num_heads = config.num_attention_heads
        self.hidden_dropout_prob = float(config.hidden_dropout_prob)
        self.attention_probs_dropout_prob = float(config.attention_probs_dropout_prob)
        self.attention = _Attention(config)
        self.intermediate = _Intermediate(config)
        self.output = _SelfOutput(config, config.intermediate_size)
        self.checkpoint = False

    def layer_params(self):
        a, o = self.attention, self.output
        return tuple(_bf16(p) for p in (a.self.query.weight, a.self.query.bias, a.self.key.weight, a.self.key.bias, a.self.value.weight,
                                        a.self.value.bias, a.output.dense.weight, a.output.dense.bias, a.output.LayerNorm.weight,
                                        a.output.LayerNorm.bias, self.intermediate.dense.weight, self.intermediate.dense.bias,
                                        o.dense.weight, o.dense.bias, o.LayerNorm.weight, o.LayerNorm.bias))

    def forward_tokens(self, x2d, key_bias, B, L):
        drop = None
        if self.training and (self.hidden_dropout_prob > 0 or self.attention_probs_dropout_prob > 0):
            seeds = [Fn.ops.next_dropout_seed() for _ in range(3)]  # attention probabilities, self-output, output
            drop = (self.hidden_dropout_prob, self.attention_probs_dropout_prob, *seeds)
        return Fn.BertLayerFn.apply(x2d, *self.layer_params(), key_bias, B, L, self.num_heads, self.output.LayerNorm.eps, self.checkpoint, drop)

    def forward(self, hidden_states, attention_mask=None, head_mask=None):
        if head_mask is not None:
            raise NotImplementedError("b200mm BertLayer: head_mask is not supported")
        B, L, Hd = hidden_states.shape
        y = self.forward_tokens(_bf16(hidden_states).reshape(B * L, Hd).contiguous(), _key_bias(attention_mask, B, L), B, L)
        return (y.view(B, L, Hd),)


def _key_bias(extended_mask, B, L):
    """Additive mask [B,1,1,L] (or [B,L]) -> fp32 [B, L] key bias, or None."""
    if extended_mask is None:
        return None
    kb = extended_mask.reshape(B, L) if extended_mask.numel() == B * L else None
    if kb is None:
        raise NotImplementedError("b200mm: only key-padding masks broadcast over heads and queries ([B,1,1,L]) are supported")
    return kb.float().contiguous()


class BertEncoder(nn.Module):
    """modeling_bert.py:273-314. Returns a tuple whose first element is the last hidden state."""

    def __init__(self, config):
        super().__init__()
        self.output_attentions = config.output_attentions
        self.output_hidden_states = config.output_hidden_states
        self.grad_checkpointing = False
        self.layer = nn.ModuleList([BertLayer(config) for _ in range(config.num_hidden_layers)])

    def forward(self, hidden_states, attention_mask=None, head_mask=None):
        if self.output_attentions:
            raise NotImplementedError("b200mm BertEncoder: attention probabilities are never materialised (output_attentions)")
        if head_mask is not None and any(h is not None for h in head_mask):
            raise NotImplementedError("b200mm BertEncoder: head_mask is not supported")
        B, L, Hd = hidden_states.shape
        kb = _key_bias(attention_mask, B, L)
        x = _bf16(hidden_states).reshape(B * L, Hd).contiguous()
        all_hidden = ()
        for layer in self.layer:
            if self.output_hidden_states:
                all_hidden = all_hidden + (x.view(B, L, Hd),)
            layer.checkpoint = layer.checkpoint or self.grad_checkpointing
            x = layer.forward_tokens(x, kb, B, L)
        out = x.view(B, L, Hd)
        outputs = (out,)
        if self.output_hidden_states:
            outputs = outputs + (all_hidden + (out,),)
        return outputs


class BertModel(nn.Module):
    """modeling_bert.py:421-534 (no pooler; returns (sequence_output, None, ...))."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.embeddings = BertEmbeddings(config)
        self.encoder = BertEncoder(config)
        self.apply(self._init_weights)

    def _init_weights(self, module):
        """modeling_bert.py:408-418."""
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
        if isinstance(module, nn.Linear) and module.bias is not None:
            module.bias.data.zero_()

    def set_grad_checkpointing(self, enable=True):
        self.encoder.grad_checkpointing = enable
        for layer in self.encoder.layer:
            layer.checkpoint = enable

    def forward(self, input_ids, attention_mask=None, token_type_ids=None, position_ids=None, head_mask=None):
        if head_mask is not None:
            raise NotImplementedError("b200mm BertModel: head_mask is not supported")
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        # modeling_bert.py:487-497: additive key mask 0 / -10000
        ext = (1.0 - attention_mask.unsqueeze(1).unsqueeze(2).float()) * -10000.0
        emb = self.embeddings(input_ids, token_type_ids=token_type_ids, position_ids=position_ids)
        enc = self.encoder(emb, ext)
        return (enc[0], None) + enc[1:]
