"""Dropout (p > 0) — CPU coverage: the oracle's restatement of the counter-based mask (statistics, determinism), and the HOST logic of the
training-mode BERT tower over the emulated kernels: seeds drawn per site, the same masks in forward / backward / checkpoint recompute /
GradCache pass 2, eval mode and p = 0 untouched. Reference semantics: nn.Dropout in BertEmbeddings (modeling_bert.py:84,101),
BertSelfAttention (:124,158), BertSelfOutput (:180), BertOutput (:232) — y = x * keep / (1 - p); only the generator differs."""
import torch

from oracle import restated
from tests import emulated_ops

BF = torch.bfloat16


def test_mask_statistics_and_determinism():
    for p in (0.1, 0.5):
        k = restated.dropout_keep(0x1234_5678_9ABC_DEF0, 2048, 768, p)
        assert k.dtype == torch.bool and k.shape == (2048, 768)
        rate = k.float().mean().item()
        assert abs(rate - (1 - p)) < 4 * (p * (1 - p) / k.numel()) ** 0.5 + 1e-4, rate
        c = k.float() - (1 - p)
        var = p * (1 - p)
        # neighbouring rows / columns / diagonal are uncorrelated (|rho| within 5 sigma of 0 for ~1.5 M samples)
        for prod in (c[:-1] * c[1:], c[:, :-1] * c[:, 1:], c[:-1, :-1] * c[1:, 1:]):
            assert abs(prod.mean().item() / var) < 5 / prod.numel() ** 0.5
        # per-row and per-column keep rates stay binomial (no dead rows / columns)
        assert (k.float().mean(1) - (1 - p)).abs().max().item() < 6 * (var / 768) ** 0.5
        assert (k.float().mean(0) - (1 - p)).abs().max().item() < 6 * (var / 2048) ** 0.5
    a = restated.dropout_keep(7, 64, 64, 0.1)
    assert torch.equal(a, restated.dropout_keep(7, 64, 64, 0.1))
    assert not torch.equal(a, restated.dropout_keep(8, 64, 64, 0.1))
    # a row's decisions depend on (seed, row, col) only: a window of rows reproduces the same bits
    assert torch.equal(restated.dropout_keep(7, 16, 64, 0.1, row0=32), a[32:48])
    m = restated.attention_dropout_keep(99, 3, 4, 77, 0.1)
    assert m.shape == (3, 4, 77, 77) and abs(m.float().mean().item() - 0.9) < 3e-3
    assert restated.drop_threshold(0.0) == 0 and restated.dropout_keep(1, 8, 8, 0.0).all()


def _tiny_bert(p_hidden, p_attn, layers=2):
    from b200mm.modules.bert import BertConfig, BertModel

    torch.manual_seed(0)
    cfg = BertConfig(vocab_size_or_config_json_file=200, hidden_size=64, num_hidden_layers=layers, num_attention_heads=2, intermediate_size=128,
                     hidden_dropout_prob=p_hidden, attention_probs_dropout_prob=p_attn, max_position_embeddings=32)
    return BertModel(cfg), cfg


def _expected_masks(seed, B, L, Hd, heads, layers, p_hidden, p_attn):
    """The masks the module must have used: seeds are drawn embeddings first, then (attention, self-output, output) per layer."""
    from b200mm import ops

    ops.manual_seed(seed)
    emb = restated.dropout_keep(ops.next_dropout_seed(), B * L, Hd, p_hidden).view(B, L, Hd) if p_hidden > 0 else None
    per_layer = []
    for _ in range(layers):
        s_a, s_so, s_o = (ops.next_dropout_seed() for _ in range(3))
        per_layer.append({
            "attn": restated.attention_dropout_keep(s_a, B, heads, L, p_attn) if p_attn > 0 else None,
            "self_out": restated.dropout_keep(s_so, B * L, Hd, p_hidden).view(B, L, Hd) if p_hidden > 0 else None,
            "out": restated.dropout_keep(s_o, B * L, Hd, p_hidden).view(B, L, Hd) if p_hidden > 0 else None,
        })
    return {"emb": emb, "layers": per_layer}


def _probe(shape):
    """Fixed random cotangent (sum of squares of a LayerNorm output is constant, so it would make every gradient vanish)."""
    return torch.randn(shape, generator=torch.Generator().manual_seed(123))


def _run(model, ids, mask, seed, checkpoint=False):
    from b200mm import ops

    model.set_grad_checkpointing(checkpoint)
    for p_ in model.parameters():
        p_.grad = None
    ops.manual_seed(seed)
    out = model(ids, attention_mask=mask)[0]
    (out.float() * _probe(out.shape)).sum().backward()
    return out.detach().float(), {n: p_.grad.detach().float().clone() for n, p_ in model.named_parameters()}


def test_bert_training_dropout_host_logic_matches_oracle_with_same_masks():
    B, L, p_h, p_a = 3, 12, 0.1, 0.2
    model, cfg = _tiny_bert(p_h, p_a)
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(1, 200, (B, L), generator=g)
    mask = torch.ones(B, L, dtype=torch.long)
    mask[1, 8:] = 0
    sd16 = {k: v.to(BF).float().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    with emulated_ops.patched():
        model = model.to(BF).train()
        out, grads = _run(model, ids, mask, seed=77)
        out_ck, grads_ck = _run(model, ids, mask, seed=77, checkpoint=True)   # recompute in backward regenerates the same masks
        out_other, _ = _run(model, ids, mask, seed=78)
        model.eval()
        out_eval = model(ids, attention_mask=mask)[0].float()
    masks = _expected_masks(77, B, L, 64, 2, 2, p_h, p_a)
    ref = restated.bert_forward(sd16, ids, mask, 2, masks=masks, p_hidden=p_h, p_attn=p_a)
    (ref * _probe(ref.shape)).sum().backward()
    assert (out - ref).norm() / ref.norm() < 2e-2
    assert torch.equal(out, out_ck)
    assert (out - out_other).abs().max() > 1e-2, "a different seed must give different masks"
    ref_eval = restated.bert_forward({k: v.detach() for k, v in sd16.items()}, ids, mask, 2)
    assert (out_eval - ref_eval).norm() / ref_eval.norm() < 2e-2, "eval mode applies no dropout"
    for n, gr in grads.items():
        r = sd16[n].grad
        if r is None or r.abs().max() < 1e-6:
            continue
        assert (gr - r).norm() / r.norm() < 5e-2, (n, float((gr - r).norm() / r.norm()))
        assert (grads_ck[n] - gr).norm() / r.norm() < 1e-2, n


def test_gradcache_replays_dropout_masks():
    """Pass 2 of the GradCache driver must rebuild each micro-batch with the masks of pass 1: the result equals the single-pass step."""
    from b200mm import ops
    from b200mm.gradcache import GradCache

    B, L = 8, 10
    model, _ = _tiny_bert(0.1, 0.1, layers=1)
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(1, 200, (B, L), generator=g)
    with emulated_ops.patched():
        model = model.to(BF).train()

        def enc(x):
            return model(x)[0][:, 0, :].float()

        def loss_fn(e):
            return (e @ e.t()).square().mean()

        # single pass, micro-batched by hand with the same seed sequence: micro-batch i starts from the state pass 1 leaves it in
        ops.manual_seed(5)
        for p_ in model.parameters():
            p_.grad = None
        full = torch.cat([enc(ids[s:s + 4]) for s in (0, 4)])
        loss_fn(full).backward()
        want = {n: p_.grad.float().clone() for n, p_ in model.named_parameters()}
        ops.manual_seed(5)
        for p_ in model.parameters():
            p_.grad = None
        GradCache([enc], loss_fn, micro_batch=4).step(ids)
        state_after = ops.get_dropout_state()
        for n, p_ in model.named_parameters():
            if want[n].abs().max() < 1e-7:
                continue
            assert (p_.grad.float() - want[n]).norm() / want[n].norm() < 2e-2, n
        assert state_after[1] == 2 * (1 + 3), "the seed sequence continues after the step as if one pass had run"
