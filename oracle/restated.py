"""TEST INFRASTRUCTURE — not product code.  CPU restatement (oracle) of the AntMMF ViT+BERT contrastive hot path.

Plain PyTorch tensor arithmetic (fp32 by default, fp64 on request), written from the reference's algorithm and
operating directly on a state-dict with the reference's parameter names. It exists so that parity can be checked on
the GPU box, where ``/root/reference`` is absent; ``tests/test_oracle.py`` pins it (a) against the committed golden
vectors under ``tests/golden/`` that were produced by the real reference (``oracle/make_golden.py``) and (b), when
the reference tree is present, against the reference modules themselves.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may import this file.
The reference's own tests do not pin this path (it has no tests): parity is pinned by (a)/(b) only.

Every function cites the reference code it restates (paths relative to the AntMMF tree).
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------------------------
def layer_norm(x, w, b, eps):
    """antmmf/modules/vision/backbone/clip/model.py:213-219 (fp32 compute) and modeling_bert.py:63 (nn.LayerNorm)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def quick_gelu(x):
    """clip/model.py:222-224."""
    return x * torch.sigmoid(1.702 * x)


def gelu_erf(x):
    """clip/modeling_bert.py:31-37."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def mha(q, k, v, heads, key_bias=None, prob_keep=None, p_drop=0.0):
    """softmax(q k^T / sqrt(hd) + key_bias) v per head; q,k,v [B, L, W]; key_bias [B, L] additive or None.
    ViT: nn.MultiheadAttention inside clip/model.py:231,245-251 (no mask); BERT: modeling_bert.py:144-165.
    prob_keep [B, heads, L, L] (bool) applies nn.Dropout(p_drop) to the probabilities with a GIVEN mask (modeling_bert.py:158)."""
    B, L, W = q.shape
    hd = W // heads
    qh = q.view(B, L, heads, hd).transpose(1, 2)
    kh = k.view(B, k.shape[1], heads, hd).transpose(1, 2)
    vh = v.view(B, v.shape[1], heads, hd).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(hd)
    if key_bias is not None:
        s = s + key_bias[:, None, None, :]
    p = torch.softmax(s, dim=-1)
    if prob_keep is not None:
        p = p * prob_keep.to(p.dtype) / (1.0 - p_drop)
    return (p @ vh).transpose(1, 2).reshape(B, L, W)


# ----------------------------------------------------------------------------------------------------------------
# ViT  (clip/model.py:275-335; block :227-256)
# ----------------------------------------------------------------------------------------------------------------
def vit_block(sd, pfx, x, heads):
    W = x.shape[-1]
    h = layer_norm(x, sd[pfx + "ln_1.weight"], sd[pfx + "ln_1.bias"], 1e-5)
    qkv = h @ sd[pfx + "attn.in_proj_weight"].t() + sd[pfx + "attn.in_proj_bias"]
    a = mha(qkv[..., :W], qkv[..., W : 2 * W], qkv[..., 2 * W :], heads)
    x = x + a @ sd[pfx + "attn.out_proj.weight"].t() + sd[pfx + "attn.out_proj.bias"]
    h = layer_norm(x, sd[pfx + "ln_2.weight"], sd[pfx + "ln_2.bias"], 1e-5)
    u = h @ sd[pfx + "mlp.c_fc.weight"].t() + sd[pfx + "mlp.c_fc.bias"]
    x = x + quick_gelu(u) @ sd[pfx + "mlp.c_proj.weight"].t() + sd[pfx + "mlp.c_proj.bias"]
    return x


def patchify(image, p):
    """[B,3,H,W] -> [B, (H/p)*(W/p), 3*p*p]; patch order row-major over the grid, feature order (c, dy, dx) — the
    order in which nn.Conv2d(3, W, p, p).weight.view(W, -1) contracts (clip/model.py:289-295,310-312)."""
    B, C, H, Wd = image.shape
    g = H // p
    x = image.view(B, C, g, p, Wd // p, p).permute(0, 2, 4, 1, 3, 5)
    return x.reshape(B, g * (Wd // p), C * p * p)


def vit_forward(sd, image, heads, pfx="", n_layers=None):
    """VisionTransformer.forward, clip/model.py:309-335. Returns [B, E]."""
    w = sd[pfx + "conv1.weight"]
    width, p = w.shape[0], w.shape[-1]
    x = patchify(image, p) @ w.reshape(width, -1).t()
    cls = sd[pfx + "class_embedding"].expand(x.shape[0], 1, width)
    x = torch.cat([cls, x], dim=1) + sd[pfx + "positional_embedding"]
    x = layer_norm(x, sd[pfx + "ln_pre.weight"], sd[pfx + "ln_pre.bias"], 1e-5)
    if n_layers is None:
        n_layers = 1 + max(int(k[len(pfx) :].split(".")[2]) for k in sd if k.startswith(pfx + "transformer.resblocks."))
    for i in range(n_layers):
        x = vit_block(sd, f"{pfx}transformer.resblocks.{i}.", x, heads)
    x = layer_norm(x[:, 0, :], sd[pfx + "ln_post.weight"], sd[pfx + "ln_post.bias"], 1e-5)
    return x @ sd[pfx + "proj"]


# ----------------------------------------------------------------------------------------------------------------
# BERT  (clip/modeling_bert.py)
# ----------------------------------------------------------------------------------------------------------------
def _drop(x, keep, p_drop):
    """nn.Dropout(p_drop) in training mode with a GIVEN keep mask (same shape as x, bool): x * keep / (1 - p)."""
    return x if keep is None else x * keep.to(x.dtype) / (1.0 - p_drop)


def bert_embeddings(sd, pfx, input_ids, token_type_ids=None, eps=1e-12, inputs_embeds=None, keep=None, p_drop=0.0):
    """BertEmbeddings.forward, modeling_bert.py:86-103 (inputs_embeds variant: prj/base_vtp/.../clip_text_encoder.py:36-60);
    `keep` = the mask of `self.dropout` on the LayerNorm output (:101)."""
    B, L = input_ids.shape if inputs_embeds is None else inputs_embeds.shape[:2]
    we = sd[pfx + "word_embeddings.weight"][input_ids] if inputs_embeds is None else inputs_embeds
    pe = sd[pfx + "position_embeddings.weight"][:L][None]
    tt = torch.zeros(B, L, dtype=torch.long) if token_type_ids is None else token_type_ids
    te = sd[pfx + "token_type_embeddings.weight"][tt]
    return _drop(layer_norm(we + pe + te, sd[pfx + "LayerNorm.weight"], sd[pfx + "LayerNorm.bias"], eps), keep, p_drop)


def bert_layer(sd, pfx, x, heads, key_bias, eps=1e-12, masks=None, p_hidden=0.0, p_attn=0.0):
    """BertLayer.forward, modeling_bert.py:260-270 (post-LN). Training-mode dropout with GIVEN masks (the reference draws them from torch's
    RNG; the oracle takes them as inputs so that a counter-based mask can be checked): masks = {"attn": [B, heads, L, L],
    "self_out": [B, L, H], "out": [B, L, H]} for BertSelfAttention.dropout (:158), BertSelfOutput.dropout (:180), BertOutput.dropout (:232)."""
    masks = masks or {}
    q = x @ sd[pfx + "attention.self.query.weight"].t() + sd[pfx + "attention.self.query.bias"]
    k = x @ sd[pfx + "attention.self.key.weight"].t() + sd[pfx + "attention.self.key.bias"]
    v = x @ sd[pfx + "attention.self.value.weight"].t() + sd[pfx + "attention.self.value.bias"]
    c = mha(q, k, v, heads, key_bias, masks.get("attn"), p_attn)
    a = _drop(c @ sd[pfx + "attention.output.dense.weight"].t() + sd[pfx + "attention.output.dense.bias"], masks.get("self_out"), p_hidden)
    x1 = layer_norm(a + x, sd[pfx + "attention.output.LayerNorm.weight"], sd[pfx + "attention.output.LayerNorm.bias"], eps)
    i = gelu_erf(x1 @ sd[pfx + "intermediate.dense.weight"].t() + sd[pfx + "intermediate.dense.bias"])
    o = _drop(i @ sd[pfx + "output.dense.weight"].t() + sd[pfx + "output.dense.bias"], masks.get("out"), p_hidden)
    return layer_norm(o + x1, sd[pfx + "output.LayerNorm.weight"], sd[pfx + "output.LayerNorm.bias"], eps)


def bert_encoder(sd, pfx, x, heads, attention_mask, eps=1e-12, n_layers=None, masks=None, p_hidden=0.0, p_attn=0.0):
    """BertEncoder.forward (modeling_bert.py:283-314) with the additive mask of BertModel.forward (:487-497):
    (1 - mask) * -10000 on the key axis."""
    key_bias = (1.0 - attention_mask.to(x.dtype)) * -10000.0
    if n_layers is None:
        n_layers = 1 + max(int(k[len(pfx) :].split(".")[1]) for k in sd if k.startswith(pfx + "layer."))
    for i in range(n_layers):
        x = bert_layer(sd, f"{pfx}layer.{i}.", x, heads, key_bias, eps, masks[i] if masks else None, p_hidden, p_attn)
    return x


def bert_forward(sd, input_ids, attention_mask, heads, pfx="", eps=1e-12, masks=None, p_hidden=0.0, p_attn=0.0):
    """BertModel.forward, modeling_bert.py:469-534. Returns sequence_output [B, L, H]. masks = {"emb": [B, L, H], "layers": [per-layer dict]}
    switches the training-mode dropouts on with given masks (see bert_layer)."""
    masks = masks or {}
    x = bert_embeddings(sd, pfx + "embeddings.", input_ids, eps=eps, keep=masks.get("emb"), p_drop=p_hidden)
    return bert_encoder(sd, pfx + "encoder.", x, heads, attention_mask, eps, masks=masks.get("layers"), p_hidden=p_hidden, p_attn=p_attn)


# ----------------------------------------------------------------------------------------------------------------
# counter-based dropout masks — CPU restatement of hash32 / drop_stream_key / drop_keep (csrc/common.cuh), bit-exact.
# The reference's nn.Dropout draws from torch's stateful RNG (modeling_bert.py:84,124,158,180,232); b200mm replaces the generator, not
# the arithmetic: y = x * keep / (1 - p). The tests feed the SAME masks to this oracle and compare everything downstream.
# ----------------------------------------------------------------------------------------------------------------
def _hash32(x):
    import numpy as np

    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def _stream_key(seed, stream):
    import numpy as np

    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    stream = stream.astype(np.uint64)
    lo = (stream & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hi = (stream >> np.uint64(32)).astype(np.uint32)
    a = _hash32(np.uint32(seed & 0xFFFFFFFF) ^ (lo * np.uint32(0x9E3779B9)))
    return _hash32(a + np.uint32(seed >> 32) + hi * np.uint32(0x85EBCA6B))


def drop_threshold(p):
    t = float(torch.tensor(p, dtype=torch.float32)) * 4294967296.0  # the C side receives p as a float
    return 4294967295 if t >= 4294967295.0 else int(t)


def dropout_keep(seed, rows, cols, p, row0=0):
    """keep[r, c] of b200mm_dropout / the GEMM epilogue for rows row0 .. row0+rows-1: stream = row, index = column. bool [rows, cols]."""
    import numpy as np

    with np.errstate(over="ignore"):
        key = _stream_key(seed, np.arange(row0, row0 + rows, dtype=np.uint64))[:, None]
        idx = np.arange(cols, dtype=np.uint32)[None, :] * np.uint32(0x9E3779B9)
        keep = _hash32(key ^ idx) >= np.uint32(drop_threshold(p))
    return torch.from_numpy(keep)


def attention_dropout_keep(seed, B, H, L, p):
    """keep[b, h, q, k] of the attention kernels: stream = b * H + h, index = q * L + k. bool [B, H, L, L]."""
    import numpy as np

    with np.errstate(over="ignore"):
        key = _stream_key(seed, np.arange(B * H, dtype=np.uint64))[:, None]
        idx = np.arange(L * L, dtype=np.uint32)[None, :] * np.uint32(0x9E3779B9)
        keep = _hash32(key ^ idx) >= np.uint32(drop_threshold(p))
    return torch.from_numpy(keep).view(B, H, L, L)


# ----------------------------------------------------------------------------------------------------------------
# CN-CLIP dual encoder  (clip/cn_model.py:201-226)
# ----------------------------------------------------------------------------------------------------------------
def cnclip_forward(sd, image, text, vision_heads, text_heads, pad_id=0):
    """CNCLIP.forward: returns (image_features, text_features, logits_per_image, logits_per_text)."""
    img = vit_forward(sd, image, vision_heads, pfx="visual.")
    mask = text.ne(pad_id)
    seq = bert_forward(sd, text, mask, text_heads, pfx="bert.")
    txt = seq[:, 0, :] @ sd["text_projection"]
    img = img / img.norm(dim=-1, keepdim=True)
    txt = txt / txt.norm(dim=-1, keepdim=True)
    logits = sd["logit_scale"].exp() * img @ txt.t()
    return img, txt, logits, logits.t()


# ----------------------------------------------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------------------------------------------
def symmetric_info_nce(logits_per_image):
    """0.5*(CE(S, arange) + CE(S^T, arange)). The reference ships the one-direction form as CrossEn
    (prj/dmae_vtp/roi_univl/univl/model/dmae_utils.py:528-537) applied both ways (univl_video_ret.py (dmae) :464-469)."""
    S = logits_per_image
    d = torch.diagonal(S)
    return 0.5 * ((torch.logsumexp(S, 1) - d).mean() + (torch.logsumexp(S, 0) - d).mean())


def mil_nce_n1(sim_t2v):
    """UnivlForVideoTextRetrieval.get_mil_nce_loss (prj/base_vtp/roi_univl/univl/model/univl_video_ret.py:146-197)
    for n_clips == 1: sim_t2v[i, j] = text_i · video_j.
    loss = mean_j( LSE( {S[i, j] for all i} ∪ {S[j, k] for k != j} ) - S[j, j] )."""
    S = sim_t2v
    B = S.shape[0]
    eye = torch.eye(B, dtype=torch.bool)
    row = S.masked_fill(eye, float("-inf"))  # text_j against all other videos
    both = torch.cat([S.t(), row], dim=1)  # video_j against all texts, then the masked row
    return (torch.logsumexp(both, dim=1) - torch.diagonal(S)).mean()


def mil_nce_clips(sim3):
    """get_mil_nce_loss as driven by forward_stage1 for n_clips >= 1 (univl_video_ret.py:146-197 and :357-387).
    sim3[i, j, c] = text_i · clip c of video_j  ([B, B, n], the output of get_l1_simi_matrix). forward_stage1 repeats every text
    row n times, so in the row of video j / clip c the column block "all texts" holds each S[i, j, c] n times; the block "this text
    against all clips" has the n clips of its own video masked out. Only the middle clip c = n // 2 of each video is scored:
        loss = mean_j( log( n * sum_i e^{S[i,j,c]} + sum_{k != j, c'} e^{S[j,k,c']} ) - S[j,j,c] - ln n )."""
    B, _, n = sim3.shape
    c = n // 2
    from_video = sim3[:, :, c].t()                                    # [j, i] = S[i, j, c]
    own = torch.eye(B, dtype=torch.bool)[:, :, None].expand(B, B, n)
    from_text = sim3.masked_fill(own, float("-inf")).reshape(B, B * n)  # [j, (k, c')] without video j's own clips
    lse_v = torch.logsumexp(from_video, dim=1) + torch.log(torch.tensor(float(n)))
    total = torch.logaddexp(lse_v, torch.logsumexp(from_text, dim=1))
    pos = from_video[torch.arange(B), torch.arange(B)]
    return (total - pos - torch.log(torch.tensor(float(n)))).mean()


def moco_nce(pos, neg, T):
    """MocoUtils.moco_loss, prj/base_vtp/roi_univl/univl/model/moco_utils.py:71-81:
    mean( LSE([pos, neg]/T) - LSE(pos/T) ); pos [N, P], neg [N, K]."""
    allv = torch.cat([pos, neg], dim=1) / T
    return (torch.logsumexp(allv, dim=1) - torch.logsumexp(pos / T, dim=1)).mean()


def l1_simi_matrix(text, video, n_clips=1):
    """get_l1_simi_matrix (univl_video_ret.py:199-226) with cal_cross=True: [B_t, B_v, n_clips]."""
    E = text.shape[-1]
    return torch.matmul(video.view(-1, n_clips, E), text.t()).permute(2, 0, 1)


def retrieval_ranks(sim):
    """0-based rank of the positive (diagonal) of every row of `sim` in descending order.
    Restates _compute_retrieval_metrics (antmmf/modules/metrics/global_retrieval_recall.py:12-27) and cal_ret_metric
    (prj/base_vtp/roi_univl/univl/model/univl_video_pretrain.py:294-312): the reference sorts -sim and looks up where the diagonal
    value lands; that position is the number of strictly larger entries, and — faithfully — every entry EQUAL to the positive
    yields one more listed position (so ties make the output longer than the number of rows)."""
    import numpy as np

    x = np.asarray(sim.detach().cpu().numpy() if isinstance(sim, torch.Tensor) else sim)
    out = []
    for i in range(x.shape[0]):
        greater = int((x[i] > x[i, i]).sum())
        equal = int((x[i] == x[i, i]).sum())
        out.extend(range(greater, greater + equal))
    return np.asarray(out, dtype=np.int64)


def recall_from_ranks(ranks, eps=1e-10):
    """mr / r@1 / r@5 / r@10 of _cal_recall (global_retrieval_recall.py:91-103): median rank is 1-based, recalls are fractions."""
    import numpy as np

    ranks = np.asarray(ranks)
    rec = lambda k: float((ranks < k).sum() / (len(ranks) + eps))  # noqa: E731
    return {"mr": float(np.median(ranks) + 1), "r@1": rec(1), "r@5": rec(5), "r@10": rec(10)}


def sym_recall(sim, t2v, v2t):
    """Both retrieval directions with several ground truths per query (_cal_sym_recall, global_retrieval_recall.py:30-88):
    rows = texts, columns = visuals; a query's rank is the best rank among its ground-truth ids; r@k counts queries whose best
    ground truth is inside the top k; mr = median best rank + 1. (No ties assumed: the reference's argsort order of ties is arbitrary.)"""
    import numpy as np

    x = np.asarray(sim.detach().cpu().numpy() if isinstance(sim, torch.Tensor) else sim)

    def one_direction(mat, gts, tag):
        best = np.asarray([min(int((mat[i] > mat[i, g]).sum()) for g in set(gts[i])) for i in range(mat.shape[0])])
        r = {k: float((best < k).sum()) / mat.shape[0] for k in (1, 5, 10)}
        return {f"{tag}-mean_recall": (r[1] + r[5] + r[10]) / 3.0, f"{tag}-r@1": r[1], f"{tag}-r@5": r[5], f"{tag}-r@10": r[10],
                f"{tag}-mr": float(np.median(best) + 1)}

    out = one_direction(x, t2v, "t2v")
    out.update(one_direction(x.T, v2t, "v2t"))
    return out


def to_dtype(sd, dtype):
    return {k: (v.to(dtype) if torch.is_floating_point(v) else v) for k, v in sd.items()}


# ----------------------------------------------------------------------------------------------------------------
# M²-Encoder (BEiT-3 multiway transformer)  — prj/M2_Encoder/vlmo  
# ----------------------------------------------------------------------------------------------------------------
def xpos(x, scale_base=512, downscale=False):
    """XPOS.forward for offset 0 (vlmo/torchscale/component/xpos_relative_position.py:41-61) on x [N, L, hd]: per-position, per-pair scale
    zeta_i^(pos/scale_base) with zeta_i = (2i + 0.4 hd)/(1.4 hd) and pos in [-(L//2 rounded up), ...), sin/cos of (l * 10000^(-i/(hd/2)))
    for l = 0..L-1 (:9-13), "rotate every two" pairing (:17-21); k uses 1/scale."""
    N, L, hd = x.shape
    base = (torch.arange(0, hd, 2) + 0.4 * hd) / (1.4 * hd)
    min_pos = -(L // 2) if L % 2 == 0 else -((L + 1) // 2)  # python: -(L) // 2 floors
    pos = torch.arange(min_pos, min_pos + L).to(base)
    scale = base[None, :] ** (pos[:, None] / scale_base)
    inv_freq = 1.0 / (10000 ** (torch.arange(0, hd // 2) / (hd // 2)))
    ang = torch.arange(L, dtype=torch.float)[:, None] * inv_freq[None, :]
    if downscale:
        scale = 1 / scale
    cs = (torch.cos(ang) * scale).repeat_interleave(2, dim=1).to(x)
    sn = (torch.sin(ang) * scale).repeat_interleave(2, dim=1).to(x)
    rot = torch.stack((-x[..., 1::2], x[..., ::2]), dim=-1).flatten(-2)
    return x * cs + rot * sn


def m2_attention(sd, pfx, x, heads, way, key_pad=None, eps=1e-5, xpos_scale_base=None):
    """MultiheadAttention.forward, vlmo/torchscale/component/multihead_attention.py:66-154 with the multiway expert
    `way` ("A" vision / "B" language, multiway_network.py:33-45): q scaled by hd^-0.5 BEFORE q·k^T (:95), key padding
    → -inf (:129-135), softmax in fp32 (:141), sub-LN `inner_attn_ln` on the merged heads (:148-149), out_proj (:151)."""
    B, L, W = x.shape
    hd = W // heads
    lin = lambda n, t: t @ sd[f"{pfx}{n}.{way}.weight"].t() + sd[f"{pfx}{n}.{way}.bias"]  # noqa: E731
    q = (lin("q_proj", x) * hd ** -0.5).view(B, L, heads, hd).transpose(1, 2)
    k = lin("k_proj", x).view(B, L, heads, hd).transpose(1, 2)
    v = lin("v_proj", x).view(B, L, heads, hd).transpose(1, 2)
    if xpos_scale_base is not None:  # multihead_attention.py:112-118 (offset 0)
        k = xpos(k.reshape(B * heads, L, hd), xpos_scale_base, downscale=True).view(B, heads, L, hd)
        q = xpos(q.reshape(B * heads, L, hd), xpos_scale_base, downscale=False).view(B, heads, L, hd)
    s = q @ k.transpose(-1, -2)
    if key_pad is not None:
        s = s.masked_fill(key_pad.bool()[:, None, None, :], float("-inf"))
    a = (torch.softmax(s.float(), dim=-1).to(s.dtype) @ v).transpose(1, 2).reshape(B, L, W)
    a = layer_norm(a, sd[f"{pfx}inner_attn_ln.{way}.weight"], sd[f"{pfx}inner_attn_ln.{way}.bias"], eps)
    return lin("out_proj", a)


def m2_ffn(sd, pfx, x, way, eps=1e-5):
    """FeedForwardNetwork.forward, vlmo/torchscale/component/feedforward_network.py:117-128: fc2(ffn_layernorm(gelu(fc1 x)))
    (activation "gelu" = exact erf GELU, :97-103; sub-LN over the ffn width)."""
    u = x @ sd[f"{pfx}{way}.fc1.weight"].t() + sd[f"{pfx}{way}.fc1.bias"]
    g = layer_norm(F.gelu(u), sd[f"{pfx}{way}.ffn_layernorm.weight"], sd[f"{pfx}{way}.ffn_layernorm.bias"], eps)
    return g @ sd[f"{pfx}{way}.fc2.weight"].t() + sd[f"{pfx}{way}.fc2.bias"]


def m2_encoder_layer(sd, pfx, x, heads, way, key_pad=None, eps=1e-5, xpos_scale_base=None):
    """EncoderLayer.forward, vlmo/torchscale/architecture/encoder.py:113-168 with normalize_before=True, subln=True,
    alpha = 1 (no deepnorm), dropout / drop_path 0, no MoE."""
    h = layer_norm(x, sd[f"{pfx}self_attn_layer_norm.{way}.weight"], sd[f"{pfx}self_attn_layer_norm.{way}.bias"], eps)
    x = x + m2_attention(sd, pfx + "self_attn.", h, heads, way, key_pad, eps, xpos_scale_base)
    h = layer_norm(x, sd[f"{pfx}final_layer_norm.{way}.weight"], sd[f"{pfx}final_layer_norm.{way}.bias"], eps)
    return x + m2_ffn(sd, pfx + "ffn.", h, way, eps)


def m2_encoder(sd, pfx, x, heads, way, key_pad=None, eps=1e-5, xpos_scale_base=None):
    """Encoder.forward, architecture/encoder.py:388-482, from token embeddings `x` that already carry positions:
    zero the padded rows (:440), the layers, final `layer_norm` (:469-470; normalize_output=True)."""
    if key_pad is not None:
        x = x * (1 - key_pad.unsqueeze(-1).to(x.dtype))
    n_layers = 1 + max(int(k[len(pfx) :].split(".")[1]) for k in sd if k.startswith(pfx + "layers."))
    for i in range(n_layers):
        x = m2_encoder_layer(sd, f"{pfx}layers.{i}.", x, heads, way, key_pad, eps, xpos_scale_base)
    return layer_norm(x, sd[f"{pfx}layer_norm.{way}.weight"], sd[f"{pfx}layer_norm.{way}.bias"], eps)


def m2_vision_embed(sd, pfx, image):
    """VisionEmbedding.forward, vlmo/torchscale/component/embedding.py:67-83 (conv with bias, CLS prepended; no mask
    positions) + PositionalEmbedding expert A starting at position 2 (embedding.py:93-110, encoder.py:353-363)."""
    w = sd[pfx + "vision_embed.proj.weight"]
    width, p = w.shape[0], w.shape[-1]
    x = patchify(image, p) @ w.reshape(width, -1).t() + sd[pfx + "vision_embed.proj.bias"]
    x = torch.cat([sd[pfx + "vision_embed.cls_token"].expand(x.shape[0], 1, width), x], dim=1)
    L = x.shape[1]
    return x + sd[pfx + "encoder.embed_positions.A.weight"][2 : L + 2]


def m2_text_embed(sd, pfx, ids):
    """TextEmbedding lookup (embedding.py:86-90) + PositionalEmbedding expert B, positions 2..L+1."""
    L = ids.shape[1]
    return sd[pfx + "text_embed.weight"][ids] + sd[pfx + "encoder.embed_positions.B.weight"][2 : L + 2]


def m2_infer_image(sd, image, heads, xpos_scale_base=None):
    """VLMo.infer_image, vlmo/modules/vlmo_module.py:364-405 (the caller passes the already inception-normalised
    tensor, :385): backbone with expert A, backbone_vl with expert A (split −1), ITC heads on the CLS rows, L2-norm."""
    h = m2_encoder(sd, "backbone.encoder.", m2_vision_embed(sd, "backbone.", image), heads, "A", xpos_scale_base=xpos_scale_base)
    hv = m2_encoder(sd, "backbone_vl.", h, heads, "A", xpos_scale_base=xpos_scale_base)
    f = h[:, 0] @ sd["itc_image_proj.fc.weight"].t()
    fv = hv[:, 0] @ sd["itc_vl_image_proj.fc.weight"].t()
    return h, f / f.norm(dim=-1, keepdim=True), fv / fv.norm(dim=-1, keepdim=True)


def m2_infer_text(sd, ids, masks, heads, xpos_scale_base=None):
    """VLMo.infer_text, vlmo_module.py:323-362: backbone with expert B and key padding = 1 − text_masks; backbone_vl
    with expert A (split −1, :343) on the language hiddens; ITC heads on CLS, L2-norm."""
    pad = 1 - masks
    h = m2_encoder(sd, "backbone.encoder.", m2_text_embed(sd, "backbone.", ids), heads, "B", pad, xpos_scale_base=xpos_scale_base)
    hv = m2_encoder(sd, "backbone_vl.", h, heads, "A", pad, xpos_scale_base=xpos_scale_base)
    f = h[:, 0] @ sd["itc_text_proj.fc.weight"].t()
    fv = hv[:, 0] @ sd["itc_vl_text_proj.fc.weight"].t()
    return h, f / f.norm(dim=-1, keepdim=True), fv / fv.norm(dim=-1, keepdim=True)


def m2_itc_loss(sd, img_f, txt_f, img_fv, txt_fv):
    """Symmetric InfoNCE on both ITC head pairs with their own temperatures: logits = exp(logit_scale)·I·Tᵀ as in
    prj/M2_Encoder/m2_encoder.py:92-95 (`logit_vl_scale` for the VL-layer heads, vlmo_module.py:193-194)."""
    l1 = symmetric_info_nce(sd["logit_scale"].exp() * img_f @ txt_f.t())
    l2 = symmetric_info_nce(sd["logit_vl_scale"].exp() * img_fv @ txt_fv.t())
    return l1 + l2


# ----------------------------------------------------------------------------------------------------------------
# Stage-2 cross-modal retrieval: blockwise N x M cross-encoder scoring and hard-negative mining
# prj/base_vtp/roi_univl/univl/model/univl_video_ret.py:33-144, :389-443  
# State-dict names: "cross_encoder.layer.{i}.…" (BertEncoder), "text_projection" [H, E],
# "similarity_dense.0.{weight,bias}" [2E, E], "similarity_dense.2.{weight,bias}" [1, 2E].
# ----------------------------------------------------------------------------------------------------------------
def cross_pair_logits(sd, text_seq, text_mask, vis_seq, vis_mask, heads):
    """One cross-encoder score per ALIGNED pair p: get_cross_output (univl_video_base.py:229-271, arch 'clip', n_clips 1) —
    concat [text ; visual] tokens and masks, additive key mask (1−m)·−10000, the BERT layers, CLS @ text_projection — followed by
    similarity_dense = Linear → ReLU → Linear(·, 1) (univl_video_ret.py:24-28, dropout p = 0). Returns [P]."""
    embed = torch.cat([text_seq, vis_seq], dim=1)
    mask = torch.cat([text_mask, vis_mask], dim=1)
    seq = bert_encoder(sd, "cross_encoder.", embed, heads, mask.to(embed.dtype))
    pooled = seq[:, 0, :] @ sd["text_projection"]
    h = torch.relu(pooled @ sd["similarity_dense.0.weight"].t() + sd["similarity_dense.0.bias"])
    return (h @ sd["similarity_dense.2.weight"].t() + sd["similarity_dense.2.bias"]).squeeze(-1)


def cross_similarity(sd, text_seq, text_mask, vis_seq, vis_mask, heads):
    """_cross_similarity (univl_video_ret.py:33-89): every text against every video; the reference walks the texts in blocks of 5
    (memory only) — the result is the [B_text, B_video] matrix of pair scores."""
    Bt, Bv = text_seq.shape[0], vis_seq.shape[0]
    ti = torch.arange(Bt).repeat_interleave(Bv)
    vi = torch.arange(Bv).repeat(Bt)
    return cross_pair_logits(sd, text_seq[ti], text_mask[ti], vis_seq[vi], vis_mask[vi], heads).view(Bt, Bv)


def hard_mining_indices(l1_simi, beg_idx, bsz, method="top_k"):
    """Negative selection of _cross_similarity_hard_mining (univl_video_ret.py:107-134), row by row with the same torch.topk
    calls (sorted=False: the slot order is whatever topk returns on this device, exactly as in the reference). Row i of the
    result lists the bsz global video indices scored against local text i; slot i is overwritten with the positive beg_idx + i.
    `l1_simi` is not modified (the reference subtracts 100 from the diagonal of its detached clone in place)."""
    out = torch.empty((bsz, bsz), dtype=torch.long, device=l1_simi.device)
    for i in range(bsz):
        raw = beg_idx + i
        row = l1_simi[raw].clone()
        if method == "top_k":
            row[raw] -= 100.0
            _, chosen = torch.topk(row, bsz, sorted=False)
        elif method == "nearliest":
            row = (row - row[raw]).abs()
            row[raw] = 100.0
            _, chosen = torch.topk(row, bsz, sorted=False, largest=False)
        else:
            raise ValueError(method)
        chosen = chosen.clone()
        chosen[i] = raw
        out[i] = chosen
    return out


def cross_similarity_hard_mining(sd, text_seq, text_mask, vis_all, vis_mask_all, chosen, heads):
    """_cross_similarity_hard_mining (univl_video_ret.py:91-144) given the selection matrix `chosen` [bsz, bsz]
    (hard_mining_indices): entry [i, j] scores local text i against gathered video chosen[i, j]."""
    bsz = text_seq.shape[0]
    ti = torch.arange(bsz).repeat_interleave(bsz)
    vi = chosen.reshape(-1)
    return cross_pair_logits(sd, text_seq[ti], text_mask[ti], vis_all[vi], vis_mask_all[vi], heads).view(bsz, bsz)


def hard_mining_weights(l1_diag, method="top_k"):
    """Row re-weighting of forward_stage2 with re_weight_method == 'median' (univl_video_ret.py:414-430): rows whose level-1
    positive score is above the batch MEAN (named 'median' in the reference) get max((mean − min)/(d − min), 0.2), others 1.
    With re_sample_method 'top_k' the reference reads the diagonal AFTER _cross_similarity_hard_mining subtracted 100 from it in
    place (:113, through the row view of the detached clone); the formula is shift-invariant up to fp32 rounding, and the shift
    is reproduced here so that the weights are bit-identical."""
    if method == "top_k":
        l1_diag = l1_diag - 100.0
    mean, mn = l1_diag.mean(), l1_diag.min()
    w = torch.ones_like(l1_diag)
    above = l1_diag > mean
    w[above] = torch.clamp_min((mean - mn) / (l1_diag[above] - mn), 0.2)
    return w


def mil_nce_matrix(S, weight=None):
    """get_mil_nce_loss (univl_video_ret.py:146-197) on an explicit square score matrix (n_pair 1) with the optional row weights:
    mean_j w_j · ( LSE( {S[i, j] ∀i} ∪ {S[j, k], k ≠ j} ) − S[j, j] )."""
    B = S.shape[0]
    eye = torch.eye(B, dtype=torch.bool, device=S.device)
    both = torch.cat([S.t(), S.masked_fill(eye, float("-inf"))], dim=1)
    per_row = torch.logsumexp(both, dim=1) - torch.diagonal(S)
    return (per_row * weight).mean() if weight is not None else per_row.mean()


def _multiway(fn, x, split):
    """MultiwayNetwork.forward with 0 < split_position < L (vlmo/torchscale/component/multiway_network.py:38-45): expert A on the first
    `split` tokens, expert B on the rest, concatenated along the sequence axis. fn(way, x_part) -> y_part."""
    return torch.cat([fn("A", x[:, :split]), fn("B", x[:, split:])], dim=1)


def m2_encoder_layer_mixed(sd, pfx, x, heads, split, key_pad=None, eps=1e-5, xpos_scale_base=None):
    """EncoderLayer.forward (architecture/encoder.py:113-168) for a fused vision + language sequence: every per-token module is a
    MultiwayNetwork split at `split`; attention (multihead_attention.py:66-154) runs over the joint sequence."""
    B, L, W = x.shape
    hd = W // heads
    ap = pfx + "self_attn."
    ln = lambda name: (lambda way, t: layer_norm(t, sd[f"{name}.{way}.weight"], sd[f"{name}.{way}.bias"], eps))  # noqa: E731
    lin = lambda name: (lambda way, t: t @ sd[f"{name}.{way}.weight"].t() + sd[f"{name}.{way}.bias"])  # noqa: E731
    h = _multiway(ln(pfx + "self_attn_layer_norm"), x, split)
    q = (_multiway(lin(ap + "q_proj"), h, split) * hd ** -0.5).view(B, L, heads, hd).transpose(1, 2)
    k = _multiway(lin(ap + "k_proj"), h, split).view(B, L, heads, hd).transpose(1, 2)
    v = _multiway(lin(ap + "v_proj"), h, split).view(B, L, heads, hd).transpose(1, 2)
    if xpos_scale_base is not None:
        k = xpos(k.reshape(B * heads, L, hd), xpos_scale_base, downscale=True).view(B, heads, L, hd)
        q = xpos(q.reshape(B * heads, L, hd), xpos_scale_base, downscale=False).view(B, heads, L, hd)
    s = q @ k.transpose(-1, -2)
    if key_pad is not None:
        s = s.masked_fill(key_pad.bool()[:, None, None, :], float("-inf"))
    a = (torch.softmax(s.float(), dim=-1).to(s.dtype) @ v).transpose(1, 2).reshape(B, L, W)
    a = _multiway(ln(ap + "inner_attn_ln"), a, split)
    x = x + _multiway(lin(ap + "out_proj"), a, split)
    h = _multiway(ln(pfx + "final_layer_norm"), x, split)
    return x + _multiway(lambda way, t: m2_ffn(sd, pfx + "ffn.", t, way, eps), h, split)


def m2_fused_forward(sd, image, ids, masks, heads, xpos_scale_base=None, eps=1e-5):
    """BEiT3.forward with both modalities (vlmo/torchscale/model/BEiT3.py:68-96) + Encoder.forward (architecture/encoder.py:388-482):
    [vision tokens ; text tokens], split position = number of vision tokens, padding mask = [zeros ; 1 - text_masks], padded rows
    zeroed, the layers, the final multiway layer_norm. Returns [B, Lv + Lt, W]."""
    xv = m2_vision_embed(sd, "backbone.", image)
    xt = m2_text_embed(sd, "backbone.", ids)
    split = xv.shape[1]
    x = torch.cat([xv, xt], dim=1)
    pad = torch.cat([torch.zeros(xv.shape[:2], dtype=torch.long), 1 - masks], dim=1)
    x = x * (1 - pad.unsqueeze(-1).to(x.dtype))
    pfx = "backbone.encoder."
    n_layers = 1 + max(int(k[len(pfx):].split(".")[1]) for k in sd if k.startswith(pfx + "layers."))
    for i in range(n_layers):
        x = m2_encoder_layer_mixed(sd, f"{pfx}layers.{i}.", x, heads, split, pad, eps, xpos_scale_base)
    return _multiway(lambda way, t: layer_norm(t, sd[f"{pfx}layer_norm.{way}.weight"], sd[f"{pfx}layer_norm.{way}.bias"], eps), x, split)
