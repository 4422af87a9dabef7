"""The "kernel to beat" of SURVEY.md §8d: the SAME arithmetic as the hot path (ViT-L/14 + BERT-base towers, symmetric InfoNCE, fwd + bwd)
run by eager PyTorch in bf16 on the same B200 — torch.nn.functional only (F.layer_norm, F.linear, F.scaled_dot_product_attention =
torch's fused attention, erf-GELU / QuickGELU, cross_entropy on materialised logits), i.e. what the reference's modules do after
`.cuda().bfloat16()` with torch 2.11's fast paths. Stand-alone (does not import the product or the oracle); weights are random with the
reference's shapes. Prints one JSON line with pairs/s so that it can be set next to bench.py's `value`.

  python tools/eager_gpu_baseline.py [--batch 256] [--steps 5] [--warmup 2]
"""
import argparse
import json
import math

import torch
import torch.nn.functional as F

BF = torch.bfloat16


def params_vit(W, layers, patch, res, out_dim, dev):
    g = lambda *s: (torch.randn(*s, device=dev) * 0.02).to(BF).requires_grad_()  # noqa: E731
    one = lambda n: torch.ones(n, device=dev, dtype=BF, requires_grad=True)  # noqa: E731
    zero = lambda n: torch.zeros(n, device=dev, dtype=BF, requires_grad=True)  # noqa: E731
    L = (res // patch) ** 2 + 1
    p = dict(conv=g(W, 3, patch, patch), cls=g(W), pos=g(L, W), ln_pre=(one(W), zero(W)), ln_post=(one(W), zero(W)), proj=g(W, out_dim), blocks=[])
    for _ in range(layers):
        p["blocks"].append(dict(ln1=(one(W), zero(W)), in_w=g(3 * W, W), in_b=zero(3 * W), out_w=g(W, W), out_b=zero(W), ln2=(one(W), zero(W)),
                                fc_w=g(4 * W, W), fc_b=zero(4 * W), pj_w=g(W, 4 * W), pj_b=zero(W)))
    return p


def vit(p, image, heads):
    W = p["cls"].shape[0]
    x = F.conv2d(image, p["conv"], stride=p["conv"].shape[-1]).flatten(2).transpose(1, 2)
    x = torch.cat([p["cls"].expand(x.shape[0], 1, W), x], 1) + p["pos"]
    x = F.layer_norm(x, (W,), *p["ln_pre"])
    B, L, _ = x.shape
    for b in p["blocks"]:
        h = F.layer_norm(x, (W,), *b["ln1"])
        q, k, v = F.linear(h, b["in_w"], b["in_b"]).view(B, L, 3, heads, W // heads).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, L, W)
        x = x + F.linear(a, b["out_w"], b["out_b"])
        h = F.layer_norm(x, (W,), *b["ln2"])
        u = F.linear(h, b["fc_w"], b["fc_b"])
        x = x + F.linear(u * torch.sigmoid(1.702 * u), b["pj_w"], b["pj_b"])
    return F.layer_norm(x[:, 0], (W,), *p["ln_post"]) @ p["proj"]


def params_bert(Hd, layers, inter, vocab, max_pos, out_dim, dev):
    g = lambda *s: (torch.randn(*s, device=dev) * 0.02).to(BF).requires_grad_()  # noqa: E731
    one = lambda n: torch.ones(n, device=dev, dtype=BF, requires_grad=True)  # noqa: E731
    zero = lambda n: torch.zeros(n, device=dev, dtype=BF, requires_grad=True)  # noqa: E731
    p = dict(word=g(vocab, Hd), pos=g(max_pos, Hd), typ=g(2, Hd), ln=(one(Hd), zero(Hd)), proj=g(Hd, out_dim), layers=[])
    for _ in range(layers):
        p["layers"].append(dict(qkv_w=g(3 * Hd, Hd), qkv_b=zero(3 * Hd), o_w=g(Hd, Hd), o_b=zero(Hd), ln1=(one(Hd), zero(Hd)), i_w=g(inter, Hd),
                                i_b=zero(inter), d_w=g(Hd, inter), d_b=zero(Hd), ln2=(one(Hd), zero(Hd))))
    return p


def bert(p, ids, heads):
    B, L = ids.shape
    Hd = p["word"].shape[1]
    x = F.layer_norm(F.embedding(ids, p["word"]) + p["pos"][:L] + p["typ"][0], (Hd,), *p["ln"], eps=1e-12)
    mask = (ids != 0)[:, None, None, :]
    for l in p["layers"]:
        q, k, v = F.linear(x, l["qkv_w"], l["qkv_b"]).view(B, L, 3, heads, Hd // heads).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v, attn_mask=mask).transpose(1, 2).reshape(B, L, Hd)
        x = F.layer_norm(F.linear(a, l["o_w"], l["o_b"]) + x, (Hd,), *l["ln1"], eps=1e-12)
        x = F.layer_norm(F.linear(F.gelu(F.linear(x, l["i_w"], l["i_b"])), l["d_w"], l["d_b"]) + x, (Hd,), *l["ln2"], eps=1e-12)
    return x[:, 0] @ p["proj"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    dev = "cuda"
    torch.manual_seed(0)
    pv = params_vit(1024, 24, 14, 224, 768, dev)
    pb = params_bert(768, 12, 3072, 21128, 512, 768, dev)
    ls = torch.tensor(math.log(1 / 0.07), device=dev, requires_grad=True)
    image = torch.randn(a.batch, 3, 224, 224, device=dev).to(BF)
    ids = torch.randint(1, 21128, (a.batch, 77), device=dev)
    ids[:, 0] = 101
    ids[:, 60:] = 0
    leaves = [t for t in ([ls] + [v for v in pv.values() if isinstance(v, torch.Tensor)] + [v for v in pb.values() if isinstance(v, torch.Tensor)])]

    def step():
        i = F.normalize(vit(pv, image, 16).float(), dim=-1)
        t = F.normalize(bert(pb, ids, 12).float(), dim=-1)
        logits = ls.exp() * i @ t.t()
        lab = torch.arange(a.batch, device=dev)
        loss = 0.5 * (F.cross_entropy(logits, lab) + F.cross_entropy(logits.t(), lab))
        loss.backward()
        return loss

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps({"what": "eager PyTorch bf16 (F.* + SDPA) ViT-L/14 + BERT-base fwd+bwd on this GPU", "batch": a.batch, "ms_per_step": round(ms, 2),
                      "pairs_per_s": round(a.batch / ms * 1e3, 1), "peak_mem_gib": round(torch.cuda.max_memory_allocated() / 2**30, 1),
                      "n_leaves": len(leaves)}))


if __name__ == "__main__":
    main()
