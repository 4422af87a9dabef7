"""Host logic of the retrieval evaluation (b200mm.retrieval: positive ranks from the counting epilogue, recall / median rank, multiple
ground truths, the metric object's block-wise collect) without a GPU, over the torch stand-in of b200mm_contrast_rank — against the
golden vectors of the unmodified reference (global_retrieval_recall.py) and the oracle's sort-based restatement."""
import os

import numpy as np
import torch

from oracle import restated
from tests import emulated_ops

BF = torch.bfloat16


def _bf(x):
    return x.to(BF).float()


def test_ranks_recall_and_metric_object_on_emulated_kernel(golden_dir, monkeypatch):
    from b200mm import retrieval
    from b200mm.retrieval import B200GlobalRetrievalRecall

    def rows_cpu(x):  # the product refuses CPU tensors; the host-logic test lifts exactly that check
        x = x.to(BF).contiguous()
        pad = (-x.shape[1]) % 8
        return torch.cat([x, x.new_zeros(x.shape[0], pad)], dim=1) if pad else x

    monkeypatch.setattr(retrieval, "_as_bf16_rows", rows_cpu)

    fx = torch.load(os.path.join(golden_dir, "retrieval.pt"), weights_only=False)
    sq, mg = fx["square"], fx["multi_gt"]
    with emulated_ops.patched():
        ranks = retrieval.positive_ranks(sq["t"], sq["v"])
        sim16 = (_bf(sq["t"]) @ _bf(sq["v"]).t()).numpy()
        assert ranks.dtype == torch.int32 and np.array_equal(ranks.numpy(), restated.retrieval_ranks(sim16))
        got, want = retrieval.cal_recall(sq["t"], sq["v"]), restated.recall_from_ranks(restated.retrieval_ranks(sim16))
        assert all(abs(got[k] - want[k]) < 1e-12 for k in want)
        assert (ranks - sq["ranks"]).abs().float().mean() < 0.5  # vs the fp32 run of the real reference (bf16 reorders near-ties only)
        simm = (_bf(mg["t"]) @ _bf(mg["v"]).t()).numpy()
        got = retrieval.cal_sym_recall(mg["t"], mg["v"], mg["t2v"], mg["v2t"])
        want = restated.sym_recall(simm, mg["t2v"], mg["v2t"])
        assert set(got) == set(mg["metrics"]) and all(abs(got[k] - want[k]) < 1e-12 for k in want)
        m = B200GlobalRetrievalRecall(simi_logit_key=["l1_simi"])
        for i in range(3):
            m.collect(i, 0, text_emb=mg["t"][20 * i: 20 * i + 20], t2v=mg["t2v"][20 * i: 20 * i + 20])
        for j in range(2):
            m.collect(0, j, visual_emb=mg["v"][6 * j: 6 * j + 6], v2t=mg["v2t"][6 * j: 6 * j + 6])
        out = m.summarize()
        assert all(abs(float(out[f"l1_simi_{k}"]) - want[k]) < 1e-12 for k in want)
