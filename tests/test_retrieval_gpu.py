"""Retrieval evaluation through the counting epilogue (b200mm_contrast_rank) against the oracle's sort-based restatement of
global_retrieval_recall.py — the committed golden vectors of the unmodified reference, and seeded larger cases."""
import os

import numpy as np
import pytest
import torch

from oracle import restated

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _bf(x):
    return x.to(BF).float()


def test_ranks_and_recall_match_reference_golden(golden_dir):
    from b200mm import retrieval

    fx = torch.load(os.path.join(golden_dir, "retrieval.pt"), weights_only=False)
    sq = fx["square"]
    t, v = sq["t"].cuda(), sq["v"].cuda()
    ranks = retrieval.positive_ranks(t, v)
    assert ranks.dtype == torch.int32 and ranks.is_cuda
    # the kernels see bf16-rounded embeddings: the oracle gets the same rounded inputs (bit-exact integer ranks expected,
    # fp32 accumulation order cannot flip a comparison unless two logits agree to ~1e-7)
    sim16 = (_bf(sq["t"]) @ _bf(sq["v"]).t()).numpy()
    assert np.array_equal(ranks.cpu().numpy(), restated.retrieval_ranks(sim16))
    got, want = retrieval.cal_recall(t, v), restated.recall_from_ranks(restated.retrieval_ranks(sim16))
    assert all(abs(got[k] - want[k]) < 1e-12 for k in want)
    # and the fp32 golden of the real reference agrees on this fixture except where bf16 rounding reorders near-equal scores
    assert (ranks.cpu() - sq["ranks"]).abs().float().mean() < 0.5
    mr, r1, r5, r10 = retrieval.cal_ret_metric(t, v)
    assert int(mr) == int(np.sort(restated.retrieval_ranks(sim16))[(len(sim16) - 1) // 2]) + 1 and abs(float(r1) - want["r@1"]) < 1e-6

    mg = fx["multi_gt"]
    sim16 = (_bf(mg["t"]) @ _bf(mg["v"]).t()).numpy()
    got = retrieval.cal_sym_recall(mg["t"].cuda(), mg["v"].cuda(), mg["t2v"], mg["v2t"])
    want = restated.sym_recall(sim16, mg["t2v"], mg["v2t"])
    assert set(got) == set(mg["metrics"])
    assert all(abs(got[k] - want[k]) < 1e-12 for k in want), (got, want)


@pytest.mark.parametrize("M,N,E", [(1000, 3000, 64), (257, 129, 40), (4096, 4096, 768)])
def test_ranks_general_shapes(M, N, E):
    from b200mm import ops, retrieval

    g = torch.Generator().manual_seed(M + N)
    q = torch.nn.functional.normalize(torch.randn(M, E, generator=g), dim=-1)
    k = torch.nn.functional.normalize(torch.randn(N, E, generator=g), dim=-1)
    gt = torch.randint(0, N, (M,), generator=g)
    k[gt[: M // 2]] = 0.7 * k[gt[: M // 2]] + 0.3 * q[: M // 2] * k[gt[: M // 2]].norm(dim=-1, keepdim=True)  # make half the positives strong
    qd, kd = q.cuda(), k.cuda()
    ranks = retrieval.positive_ranks(qd, kd, gt.cuda()).cpu().numpy()
    sim = (_bf(q).double() @ _bf(k).double().t()).numpy()
    ref = sim[np.arange(M), gt.numpy()]
    want = (sim > ref[:, None]).sum(1)
    # rows whose positive is within fp32 accumulation noise of another score are ambiguous: exclude them from the exact check
    margin = np.abs(sim - ref[:, None])
    margin[np.arange(M), gt.numpy()] = 1.0
    clear = margin.min(1) > 2e-6
    assert clear.mean() > 0.9
    assert np.array_equal(ranks[clear], want[clear])
    assert np.abs(ranks[~clear] - want[~clear]).max(initial=0) <= 2
    if M == N:
        r0 = retrieval.positive_ranks(qd, kd).cpu().numpy()
        d = np.diag(sim)
        w0 = (sim > d[:, None]).sum(1)
        m0 = np.abs(sim - d[:, None]) + np.eye(M)
        ok = m0.min(1) > 2e-6
        assert np.array_equal(r0[ok], w0[ok])


def test_metric_object_summarize(golden_dir):
    from b200mm.retrieval import B200GlobalRetrievalRecall

    fx = torch.load(os.path.join(golden_dir, "retrieval.pt"), weights_only=False)["multi_gt"]
    t, v = fx["t"].cuda(), fx["v"].cuda()
    m = B200GlobalRetrievalRecall(simi_logit_key=["l1_simi"])
    # three text batches of 20, two visual batches of 6 — like the block-wise collect of retrieval_trainer.py:200-260
    for i in range(3):
        m.collect(i, 0, text_emb=t[20 * i: 20 * i + 20], t2v=fx["t2v"][20 * i: 20 * i + 20])
    for j in range(2):
        m.collect(0, j, visual_emb=v[6 * j: 6 * j + 6], v2t=fx["v2t"][6 * j: 6 * j + 6])
    out = m.summarize()
    sim16 = (_bf(fx["t"]) @ _bf(fx["v"]).t()).numpy()
    want = restated.sym_recall(sim16, fx["t2v"], fx["v2t"])
    assert set(out) == {f"l1_simi_{k}" for k in want}
    assert all(abs(float(out[f"l1_simi_{k}"]) - want[k]) < 1e-12 and out[f"l1_simi_{k}"].dtype == torch.float64 for k in want)
    sq = m.calculate(t[:12], v)
    assert set(sq) == {"l1_simi_mr", "l1_simi_r@1", "l1_simi_r@5", "l1_simi_r@10"}
    assert float(m.calculate(t[:7], v)["l1_simi_mr"]) == 0.0
