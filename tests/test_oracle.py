"""Pins the CPU oracle (oracle/restated.py): against the committed golden vectors produced by the real reference
(oracle/make_golden.py) and, when the reference tree is present, against the reference modules directly."""
import os
import types

import pytest
import torch
import torch.nn.functional as F

from oracle import ref_loader, restated


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


@pytest.mark.parametrize("name", ["cnclip_tiny.pt", "cnclip_tiny_h80.pt"])
def test_restated_cnclip_matches_golden_forward_backward(golden_dir, name):
    fx = _load(golden_dir, name)
    cfg = fx["config"]
    sd = {k: v.clone().requires_grad_(torch.is_floating_point(v)) for k, v in fx["state_dict"].items()}
    vh = cfg["vision_width"] // cfg["vision_head_width"]
    img, txt, lpi, lpt = restated.cnclip_forward(sd, fx["image"], fx["text"], vh, cfg["text_num_attention_heads"])
    torch.testing.assert_close(img, fx["image_features"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(txt, fx["text_features"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(lpi, fx["logits_per_image"], rtol=1e-4, atol=1e-4)
    loss = restated.symmetric_info_nce(lpi)
    torch.testing.assert_close(loss, fx["loss"], rtol=1e-5, atol=1e-6)
    loss.backward()
    for n, g in fx["grads"].items():
        got = sd[n].grad
        assert got is not None, n
        denom = g.abs().max().clamp_min(1e-6)  # key.bias grads are analytically 0 (softmax shift invariance)
        assert float((got - g).abs().max() / denom) < 2e-4, n


def test_restated_losses_match_golden(golden_dir):
    fx = _load(golden_dir, "losses.pt")
    for key in ["mil_b4", "mil_b37"]:
        c = fx[key]
        t = c["t"].clone().requires_grad_()
        v = c["v"].clone().requires_grad_()
        sim = restated.l1_simi_matrix(t, v, 1).view(t.shape[0], v.shape[0])
        loss = restated.mil_nce_n1(sim)
        torch.testing.assert_close(loss, c["loss"], rtol=1e-5, atol=1e-6)
        loss.backward()
        torch.testing.assert_close(t.grad, c["dt"], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(v.grad, c["dv"], rtol=1e-4, atol=1e-6)
    # n_clips > 1 (forward_stage1's repeat + get_mil_nce_loss of the unmodified reference)
    for key in ["mil_b3_n2", "mil_b6_n3", "mil_b9_n4"]:
        c = fx[key]
        t = c["t"].clone().requires_grad_()
        v = c["v"].clone().requires_grad_()
        loss = restated.mil_nce_clips(restated.l1_simi_matrix(t, v, c["n"]))
        torch.testing.assert_close(loss, c["loss"], rtol=1e-5, atol=1e-6)
        loss.backward()
        torch.testing.assert_close(t.grad, c["dt"], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(v.grad, c["dv"], rtol=1e-4, atol=1e-6)
    # n = 1 is the same function
    c = fx["mil_b37"]
    torch.testing.assert_close(restated.mil_nce_clips(restated.l1_simi_matrix(c["t"], c["v"], 1)), c["loss"], rtol=1e-5, atol=1e-6)
    # SURVEY.md §8c known answer (1)
    assert abs(float(fx["mil_b4"]["loss"]) - 2.0541925) < 1e-6
    assert abs(float(fx["mil_b4"]["dt"].abs().sum()) - 2.4046431) < 1e-5
    assert abs(float(fx["mil_b4"]["dv"].abs().sum()) - 2.0970497) < 1e-5
    m = fx["moco"]
    torch.testing.assert_close(restated.moco_nce(m["pos"], m["neg"], m["T"]), m["loss"], rtol=1e-5, atol=1e-5)


def test_patchify_equals_conv():
    torch.manual_seed(0)
    img = torch.randn(3, 3, 32, 48)
    w = torch.randn(10, 3, 8, 8)
    ref = F.conv2d(img, w, stride=8).flatten(2).transpose(1, 2)
    got = restated.patchify(img, 8) @ w.reshape(10, -1).t()
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
def test_restated_matches_live_reference():
    ns = ref_loader.load()
    cfg = dict(ns.cn_model.CONFIGS["ViT-B-16"])
    cfg.update(vision_layers=1, text_num_hidden_layers=1, image_resolution=32, vocab_size=256,
               text_hidden_dropout_prob=0.0, text_attention_probs_dropout_prob=0.0)
    model = ref_loader.build_cnclip(cfg, seed=3)
    torch.manual_seed(5)
    image = torch.randn(3, 3, 32, 32)
    text = torch.randint(1, 256, (3, 9))
    text[:, 0] = 101
    text[1, 5:] = 0
    a = model(image, text)
    sd = {k: v for k, v in model.state_dict().items()}
    b = restated.cnclip_forward(sd, image, text, 12, 12)
    for x, y in zip(a, b):
        torch.testing.assert_close(y, x.detach(), rtol=1e-4, atol=1e-4)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
def test_reference_loss_functions_extract():
    fn = ref_loader.load_loss_functions()
    torch.manual_seed(0)
    t = F.normalize(torch.randn(4, 8))
    v = F.normalize(torch.randn(4, 8))
    sim = fn.get_l1_simi_matrix(None, t, v, 1, True).view(4, 4)
    assert abs(float(fn.get_mil_nce_loss(None, sim, 4, 1)) - 2.0541925) < 1e-6
    pos, neg = torch.randn(5, 1), torch.randn(5, 7)
    torch.testing.assert_close(fn.moco_loss(types.SimpleNamespace(T=0.05), pos, neg), restated.moco_nce(pos, neg, 0.05))


def _np_sim(t, v):
    return (t @ v.t()).numpy()


def test_restated_retrieval_metrics_match_golden(golden_dir):
    """oracle retrieval ranks / recalls against the vectors produced by the unmodified reference (make_golden.make_retrieval)"""
    import numpy as np

    fx = _load(golden_dir, "retrieval.pt")
    sq = fx["square"]
    ranks = restated.retrieval_ranks(_np_sim(sq["t"], sq["v"]))
    assert np.array_equal(ranks, sq["ranks"].numpy())
    got = restated.recall_from_ranks(ranks)
    for k, v in sq["recall"].items():
        assert abs(got[k] - v) < 1e-12, k
    mg = fx["multi_gt"]
    got = restated.sym_recall(_np_sim(mg["t"], mg["v"]), mg["t2v"], mg["v2t"])
    assert set(got) == set(mg["metrics"])
    for k, v in mg["metrics"].items():
        assert abs(got[k] - v) < 1e-12, k
    # ties: every entry equal to the positive is listed (the reference's np.where(sorted == diag) behaviour)
    tie = np.array([[1.0, 1.0, 0.0], [0.5, 0.2, 0.9], [0.0, 0.0, 0.0]])
    assert restated.retrieval_ranks(tie).tolist() == [0, 1, 2, 0, 1, 2]


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
def test_restated_retrieval_metrics_match_live_reference():
    import numpy as np

    fn = ref_loader.load_retrieval_metric_functions()
    rng = np.random.default_rng(0)
    for n in (1, 7, 33):
        x = rng.standard_normal((n, n)).astype(np.float32)
        assert np.array_equal(restated.retrieval_ranks(x), fn._compute_retrieval_metrics(x))
        a, b = restated.recall_from_ranks(restated.retrieval_ranks(x)), fn._cal_recall(x)
        assert all(abs(a[k] - float(b[k])) < 1e-12 for k in b)
    tie = np.array([[1.0, 1.0, 0.0], [0.5, 0.2, 0.9], [0.0, 0.0, 0.0]], dtype=np.float32)
    assert np.array_equal(restated.retrieval_ranks(tie), fn._compute_retrieval_metrics(tie))
    x = rng.standard_normal((20, 6)).astype(np.float32)
    t2v = [[int(rng.integers(0, 6))] for _ in range(20)]
    v2t = [sorted(set(int(i) for i in rng.integers(0, 20, size=3))) for _ in range(6)]
    a, b = restated.sym_recall(x, t2v, v2t), fn._cal_sym_recall(x, t2v, v2t)
    assert all(abs(a[k] - float(b[k])) < 1e-12 for k in b)


@pytest.mark.parametrize("name", ["m2_tiny.pt", "m2_tiny_xpos.pt"])
def test_restated_m2_encoder_matches_golden_forward_backward(golden_dir, name):
    """M²-Encoder (BEiT-3 multiway) restatement vs the unmodified reference classes (oracle/make_golden.py::make_m2); the second fixture
    was produced with args.xpos_rel_pos = True (XPOS rotary embedding on q / k, odd sequence lengths 17 and 11)."""
    fx = _load(golden_dir, name)
    heads = fx["config"]["heads"]
    xp = 512 if fx["config"].get("xpos") else None
    sd = {k: v.clone().requires_grad_(torch.is_floating_point(v)) for k, v in fx["state_dict"].items()}
    h_i, img_f, img_fv = restated.m2_infer_image(sd, fx["image"], heads, xp)
    h_t, txt_f, txt_fv = restated.m2_infer_text(sd, fx["ids"], fx["masks"], heads, xp)
    torch.testing.assert_close(h_i, fx["image_hidden"], rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(h_t, fx["text_hidden"], rtol=1e-4, atol=2e-5)
    for got, key in [(img_f, "img_f"), (txt_f, "txt_f"), (img_fv, "img_fv"), (txt_fv, "txt_fv")]:
        torch.testing.assert_close(got, fx[key], rtol=1e-4, atol=1e-5)
    loss = restated.m2_itc_loss(sd, img_f, txt_f, img_fv, txt_fv)
    torch.testing.assert_close(loss, fx["loss"], rtol=1e-5, atol=1e-6)
    loss.backward()
    assert len(fx["grads"]) > 70
    for n, g in fx["grads"].items():
        got = sd[n].grad
        assert got is not None, n
        denom = g.abs().max().clamp_min(1e-4)  # k_proj.bias grads are analytically 0 (softmax shift invariance)
        assert float((got - g).abs().max() / denom) < 3e-4, n
    # parameters the reference leaves without gradient on this path stay without gradient in the restatement too
    for n, v in sd.items():
        if torch.is_floating_point(v) and n not in fx["grads"]:
            assert v.grad is None or float(v.grad.abs().max()) == 0.0, n


def test_restated_stage2_cross_similarity_and_hard_mining_match_golden(golden_dir):
    """Stage-2 restatements vs the unmodified reference functions (oracle/make_golden.py::make_stage2): blockwise N x M scores,
    hard-negative selection (index-exact), the 'median' row weights, the weighted MIL-NCE on the score matrix, and all gradients."""
    fx = _load(golden_dir, "stage2.pt")
    heads = fx["config"]["heads"]

    def fresh():
        return {k: v.clone().requires_grad_(True) for k, v in fx["state_dict"].items()}

    c = fx["cross"]
    sd = fresh()
    seq, vis = c["seq"].clone().requires_grad_(), c["vis"].clone().requires_grad_()
    logits = restated.cross_similarity(sd, seq, c["am"], vis, c["vm"], heads)
    torch.testing.assert_close(logits, c["logits"], rtol=1e-4, atol=1e-5)
    logits.square().sum().backward()
    torch.testing.assert_close(seq.grad, c["d_seq"], rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(vis.grad, c["d_vis"], rtol=1e-3, atol=1e-6)
    for n, g in c["grads"].items():
        assert float((sd[n].grad - g).abs().max() / g.abs().max().clamp_min(1e-4)) < 3e-4, n  # key.bias grads are analytically 0

    for method in ["top_k", "nearliest"]:
        h = fx["hard_" + method]
        B = h["seq"].shape[0]
        sd = fresh()
        seq, vis = h["seq"].clone().requires_grad_(), h["vis"].clone().requires_grad_()
        l1_before = h["l1"].clone()
        chosen = restated.hard_mining_indices(h["l1"], 0, B, method)
        assert torch.equal(h["l1"], l1_before)  # not modified
        assert torch.equal(torch.diagonal(chosen), torch.arange(B))  # slot i = the positive
        l2 = restated.cross_similarity_hard_mining(sd, seq, h["am"], vis, h["vm"], chosen, heads)
        torch.testing.assert_close(l2, h["logits"], rtol=1e-4, atol=1e-5)
        w = restated.hard_mining_weights(torch.diagonal(h["l1"]), method)
        assert torch.equal(w, h["weights"])
        torch.testing.assert_close(restated.mil_nce_matrix(l2), h["loss_unweighted"], rtol=1e-5, atol=1e-6)
        loss = restated.mil_nce_matrix(l2, w)
        torch.testing.assert_close(loss, h["loss"], rtol=1e-5, atol=1e-6)
        loss.backward()
        torch.testing.assert_close(seq.grad, h["d_seq"], rtol=1e-3, atol=1e-7)
        torch.testing.assert_close(vis.grad, h["d_vis"], rtol=1e-3, atol=1e-7)
        for n, g in h["grads"].items():
            assert float((sd[n].grad - g).abs().max() / g.abs().max().clamp_min(1e-3)) < 3e-4, (method, n)  # the last bias has an analytically zero gradient (shift invariance)


@pytest.mark.parametrize("name", ["m2_tiny.pt", "m2_tiny_xpos.pt"])
def test_restated_m2_fused_input_matches_golden(golden_dir, name):
    """Multiway split INSIDE the sequence (fused vision + language input, BEiT3.forward with both modalities): restatement vs the
    unmodified reference — joint hidden states and every parameter gradient of a fixed random projection of the valid rows."""
    fx = _load(golden_dir, name)
    heads = fx["config"]["heads"]
    xp = 512 if fx["config"].get("xpos") else None
    sd = {k: v.clone().requires_grad_(torch.is_floating_point(v)) for k, v in fx["state_dict"].items()}
    out = restated.m2_fused_forward(sd, fx["image"], fx["ids"], fx["masks"], heads, xp)
    torch.testing.assert_close(out, fx["fused_hidden"], rtol=1e-4, atol=3e-5)
    Lv = out.shape[1] - fx["ids"].shape[1]
    valid = torch.cat([torch.ones(out.shape[0], Lv, dtype=torch.long), fx["masks"]], 1).unsqueeze(-1)
    (out * fx["fused_proj"] * valid).sum().backward()
    assert len(fx["fused_grads"]) > 40
    for n, g in fx["fused_grads"].items():
        got = sd[n].grad
        assert got is not None, n
        assert float((got - g).abs().max() / g.abs().max().clamp_min(1e-4)) < 3e-4, n
    # both experts of the backbone layers are exercised, the vl encoder and the heads are not
    assert "backbone.encoder.layers.0.ffn.A.fc1.weight" in fx["fused_grads"] and "backbone.encoder.layers.0.ffn.B.fc1.weight" in fx["fused_grads"]
    assert not any(k.startswith(("backbone_vl.", "itc_")) for k in fx["fused_grads"])


def test_restated_bert_dropout_matches_reference_with_preset_masks(golden_dir):
    """Training-mode dropout: tests/golden/bert_dropout.pt holds the UNMODIFIED reference BertModel run with F.dropout replaced by preset
    masks (make_golden.make_bert_dropout). The oracle, given the masks its restated hash generates from the recorded seeds, must reproduce the
    output and every gradient — this pins WHERE each dropout sits (modeling_bert.py:101,158,180,232) and the y = x keep / (1 - p) rule."""
    fx = torch.load(os.path.join(golden_dir, "bert_dropout.pt"), weights_only=False)
    c = fx["config"]
    B, L = fx["ids"].shape
    seeds = fx["seeds"]
    masks = {"emb": restated.dropout_keep(seeds[0], B * L, c["hidden"], c["p_hidden"]).view(B, L, c["hidden"]), "layers": []}
    for i in range(c["layers"]):
        s_a, s_so, s_o = seeds[1 + 3 * i: 4 + 3 * i]
        masks["layers"].append({"attn": restated.attention_dropout_keep(s_a, B, c["heads"], L, c["p_attn"]),
                                "self_out": restated.dropout_keep(s_so, B * L, c["hidden"], c["p_hidden"]).view(B, L, c["hidden"]),
                                "out": restated.dropout_keep(s_o, B * L, c["hidden"], c["p_hidden"]).view(B, L, c["hidden"])})
    # the call order the reference made: embeddings, then (probabilities, self-output, output) per layer
    assert [len(sh) for sh, _ in fx["calls"]] == [3] + [4, 3, 3] * c["layers"]
    sd = {k: v.clone().requires_grad_(torch.is_floating_point(v)) for k, v in fx["state_dict"].items()}
    out = restated.bert_forward(sd, fx["ids"], fx["mask"], c["heads"], masks=masks, p_hidden=c["p_hidden"], p_attn=c["p_attn"])
    torch.testing.assert_close(out, fx["out"], rtol=1e-4, atol=1e-5)
    (out * fx["probe"]).sum().backward()
    for n, g in fx["grads"].items():
        got = sd[n].grad
        if n == "embeddings.word_embeddings.weight":  # nn.Embedding(padding_idx=0): the reference gives row 0 no gradient (modeling_bert.py:71-73)
            assert float(g[0].abs().max()) == 0.0
            got, g = got[1:], g[1:]
        torch.testing.assert_close(got, g, rtol=2e-3, atol=1e-5, msg=lambda m, n=n: f"{n}: {m}")
    # without the masks the output differs (the fixture really ran with dropout on)
    plain = restated.bert_forward({k: v.detach() for k, v in sd.items()}, fx["ids"], fx["mask"], c["heads"])
    assert float((plain - fx["out"]).abs().max()) > 0.1
