"""Retrieval evaluation on the fused similarity kernels — SURVEY.md §8(f) rank 4.

Reference: `_compute_retrieval_metrics` / `_cal_recall` / `_cal_sym_recall` and the `GlobalRetrievalRecall` metric
(antmmf/modules/metrics/global_retrieval_recall.py:12-180) and the in-model `cal_ret_metric`
(prj/base_vtp/roi_univl/univl/model/univl_video_pretrain.py:294-312). The reference materialises the [N_text, N_visual] similarity,
moves it to the host and sorts every row (numpy) to find where the ground truth lands. Here the rank of a ground truth is COUNTED:

    rank(m, g) = #{ n != g : <q_m, k_n> > <q_m, k_g> }

by an epilogue of the tcgen05 similarity GEMM (`b200mm_contrast_rank`), so the similarity matrix exists only tile by tile in TMEM and
nothing but one int32 per query leaves the GPU. With several ground truths per query (5 captions per image) the best rank counts, as
in `_cal_sym_recall`. Ties: the count is of STRICTLY larger scores (the reference's behaviour under exact ties is an artefact of
`np.where(sorted == diag)` / an unstable argsort).
"""
from typing import Dict, List, Optional, Sequence

import torch

from . import ops

BF16 = torch.bfloat16


def _as_bf16_rows(x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise ops._lib.B200mmError("b200mm.retrieval: embeddings must be CUDA tensors (no CPU fallback exists)")
    x = x.to(BF16).contiguous()
    pad = (-x.shape[1]) % 8
    if pad:  # TMA rows are 16-byte granular
        x = torch.cat([x, x.new_zeros(x.shape[0], pad)], dim=1)
    return x


def positive_ranks(queries: torch.Tensor, keys: torch.Tensor, gt: Optional[torch.Tensor] = None) -> torch.Tensor:
    """0-based rank (int32 [M]) of key `gt[m]` (default: m) among all keys for query m, by dot-product similarity."""
    q, k = _as_bf16_rows(queries), _as_bf16_rows(keys)
    M = q.shape[0]
    if gt is None:
        # one pass computes the diagonal logit with exactly the arithmetic of the counting pass
        _, _, diag = ops.contrast_lse_partials(q, k, 1.0, 0)
        return ops.contrast_rank(q, k, 1.0, diag, diag_off=0)
    gt = gt.to(device=q.device, dtype=torch.int32).contiguous()
    if gt.shape != (M,) or int(gt.min()) < 0 or int(gt.max()) >= k.shape[0]:
        raise ValueError("positive_ranks: gt must hold one valid key index per query")
    ref = ops.rowdot(q, k[gt.long()].contiguous(), 1.0)
    return ops.contrast_rank(q, k, 1.0, ref, gt_col=gt)


def recall_from_ranks(ranks: torch.Tensor, eps: float = 1e-10) -> Dict[str, float]:
    """{"mr", "r@1", "r@5", "r@10"} with the reference's conventions (_cal_recall :91-103): 1-based median rank, numpy median
    (mean of the two middle values), recall = count / (n + eps)."""
    r = ranks.to(torch.float64)
    n = r.numel()
    return {"mr": float(torch.quantile(r, 0.5)) + 1.0, "r@1": int((r < 1).sum()) / (n + eps), "r@5": int((r < 5).sum()) / (n + eps),
            "r@10": int((r < 10).sum()) / (n + eps)}


def cal_recall(text_emb: torch.Tensor, visual_emb: torch.Tensor) -> Dict[str, float]:
    """`_cal_recall(text_emb @ visual_emb.T)` for the 1:1 case without the matrix."""
    return recall_from_ranks(positive_ranks(text_emb, visual_emb))


def _best_rank(q, k, gts: Sequence[Sequence[int]]) -> torch.Tensor:
    slots = max(len(set(g)) for g in gts)
    best = None
    for s in range(slots):
        # queries with fewer ground truths repeat their last one: the minimum is unchanged
        col = torch.tensor([sorted(set(g))[min(s, len(set(g)) - 1)] for g in gts], dtype=torch.int32)
        r = positive_ranks(q, k, col)
        best = r if best is None else torch.minimum(best, r)
    return best


def cal_sym_recall(text_emb: torch.Tensor, visual_emb: torch.Tensor, t2v: List[List[int]], v2t: List[List[int]]) -> Dict[str, float]:
    """`_cal_sym_recall(text_emb @ visual_emb.T, t2v, v2t)` (:30-88): same keys, same conventions."""
    out = {}
    for tag, q, k, gts in (("t2v", text_emb, visual_emb, t2v), ("v2t", visual_emb, text_emb, v2t)):
        best = _best_rank(q, k, gts).to(torch.float64)
        n = best.numel()
        r1, r5, r10 = (int((best < kk).sum()) / n for kk in (1, 5, 10))
        out.update({f"{tag}-mean_recall": (r1 + r5 + r10) / 3.0, f"{tag}-r@1": r1, f"{tag}-r@5": r5, f"{tag}-r@10": r10,
                    f"{tag}-mr": float(torch.quantile(best, 0.5)) + 1.0})
    return out


def cal_ret_metric(text_emb: torch.Tensor, visual_emb: torch.Tensor):
    """(mr, r@1, r@5, r@10) as device tensors — the in-model metric of univl_video_pretrain.py:294-312 (torch.median = lower middle)."""
    ranks = positive_ranks(text_emb, visual_emb)
    n = ranks.numel()
    mr = torch.median(ranks) + 1
    return (mr,) + tuple((ranks < kk).sum() / (n + 1e-10) for kk in (1, 5, 10))


class B200GlobalRetrievalRecall:
    """Embedding-collecting counterpart of GlobalRetrievalRecall (:105-180): `collect` keeps the (tiny) embeddings of each batch on the
    GPU instead of similarity blocks on the host; `summarize` returns the same `<key>_<metric>` dictionary of float64 tensors."""

    def __init__(self, name: str = "b200_global_retrieval_recall", simi_logit_key: Sequence[str] = ("l1_simi",)):
        self.name = name
        self._keys = list(simi_logit_key)
        self.reset()

    def reset(self):
        self._text, self._vis, self.gt_t2v, self.gt_v2t = {}, {}, {}, {}

    def collect(self, idx_t: int, idx_v: int, text_emb: Optional[torch.Tensor] = None, visual_emb: Optional[torch.Tensor] = None, t2v=None, v2t=None):
        if text_emb is not None and idx_t not in self._text:
            self._text[idx_t] = text_emb.detach()
        if visual_emb is not None and idx_v not in self._vis:
            self._vis[idx_v] = visual_emb.detach()
        if t2v is not None and idx_t not in self.gt_t2v:
            self.gt_t2v[idx_t] = t2v
        if v2t is not None and idx_v not in self.gt_v2t:
            self.gt_v2t[idx_v] = v2t

    def calculate(self, text_emb: torch.Tensor, visual_emb: torch.Tensor) -> Dict[str, torch.Tensor]:
        """batch-wise metric of a square batch (GlobalRetrievalRecall.calculate :150-176)."""
        if text_emb.shape[0] != visual_emb.shape[0]:
            vals = {"mr": 0.0, "r@1": 0.0, "r@5": 0.0, "r@10": 0.0}
        else:
            vals = cal_recall(text_emb, visual_emb)
        return {f"{k}_{n}": torch.tensor(v, dtype=torch.float64) for k in self._keys for n, v in vals.items()}

    def summarize(self) -> Dict[str, torch.Tensor]:
        text = torch.cat([self._text[i] for i in sorted(self._text)])
        vis = torch.cat([self._vis[i] for i in sorted(self._vis)])
        t2v = [a for _, x in sorted(self.gt_t2v.items()) for a in x]
        v2t = [a for _, x in sorted(self.gt_v2t.items()) for a in x]
        vals = cal_sym_recall(text, vis, t2v, v2t)
        return {f"{k}_{n}": torch.tensor(v, dtype=torch.float64) for k in self._keys for n, v in vals.items()}
