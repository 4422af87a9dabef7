"""MoCo-queue contrastive learning on the b200mm kernels — SURVEY.md §8 row a16.

Reference: MocoUtils (prj/base_vtp/roi_univl/univl/model/moco_utils.py:13-108) and its use in
UnivlForVideoTextRetrieval.get_simi_logits (univl_video_ret.py:263-312):
  * key encoders = momentum copies of the query encoders, updated by  p_k = m p_k + (1-m) p_q   (fused EMA kernel; the copies
    are kept in fp32 because (1-m) = 1e-4 is below bf16 resolution, the modules cast to bf16 per call),
  * queues [dim, K] of negative keys with a ring pointer (same buffer names / shapes as the reference: img_queue, img_queue_ptr,
    txt_queue, txt_queue_ptr), enqueue after an all-gather of the keys,
  * loss  mean( LSE([pos, neg]/T) - LSE(pos/T) ),  pos = <q_i, k_i>,  neg = q · queue  — computed WITHOUT materialising the
    [N, K] logits: the tcgen05 GEMM reads the queue in its native [dim, K] layout (MN-major B operand) and reduces each
    logit tile to (max, sum-exp) in the epilogue; backward recomputes the tiles into the softmax gradient (bf16 [N, K]).
"""
import copy

import torch
from torch import nn
from torch.autograd import Function

from . import ops
from .distributed import gather_tensor

BF16 = torch.bfloat16


class _MocoNceFn(Function):
    """loss = mean_i( log(e^{pos_i/T} + sum_k e^{<q_i, queue_k>/T}) - pos_i/T ),  pos_i = <q_i, kpos_i>  (one positive per row)."""

    @staticmethod
    def forward(ctx, q, k_pos, queue, T):
        N, E = q.shape
        inv_t = 1.0 / T
        pos = ops.rowdot(q, k_pos, inv_t)                                     # [N] f32, already / T
        parts = ops.contrast_lse_partials(q, queue, inv_t, ops.NO_DIAG, b_mn=True)
        loss_sum = torch.zeros(1, device=q.device, dtype=torch.float32)
        lse = ops.contrast_lse_merge(parts[:2], None, pos, -1, loss_sum)      # adds the positive logit to the row's LSE
        ctx.save_for_backward(q, k_pos, queue, lse, pos)
        ctx.meta = (inv_t, N)
        return (loss_sum / N).reshape(())

    @staticmethod
    def backward(ctx, gout):
        q, k_pos, queue, lse, pos = ctx.saved_tensors
        inv_t, N = ctx.meta
        coef = float(gout) / N
        K = queue.shape[1]
        G = ops.contrast_softgrad(q, queue, K, inv_t, ops.NO_DIAG, lse, coef, 0.0, False, None, b_mn=True)  # dL/d<q_i, queue_k>
        dq = ops.gemm(G, queue, out_f32=True)                                   # [N, E] = G [N, K] · queue[E, K]^T
        dpos = (torch.exp(pos - lse) - 1.0) * (coef * inv_t)                    # [N]: d loss / d <q_i, kpos_i>
        dq = torch.addcmul(dq, dpos[:, None], k_pos.float())
        return dq.to(BF16), None, None, None


def moco_nce(q, k_pos, queue, T):
    """Fused MoCo NCE: q, k_pos [N, E] bf16 (k_pos carries no gradient), queue [E, K] bf16."""
    return _MocoNceFn.apply(q.contiguous(), k_pos.detach().contiguous(), queue, T)


class B200MocoUtils(nn.Module):
    def __init__(self, config, img_encoder=None, txt_encoder=None):
        assert img_encoder is not None or txt_encoder is not None
        super().__init__()
        self.config = config
        get = (lambda k, d: config.get(k, d)) if hasattr(config, "get") else (lambda k, d: getattr(config, k, d))
        self.dim = get("hidden_size", None)
        self.txt_K = get("K", 16384)
        self.img_K = 16384
        self.m = get("M", 0.9999)
        self.T = get("T", 0.05)
        self.img_encoder_q = self.txt_encoder_q = None
        if img_encoder is not None:
            self.img_encoder_q = img_encoder
            self.img_encoder_k = self._momentum_copy(img_encoder)
            self.register_buffer("img_queue", torch.nn.functional.normalize(torch.randn(self.dim, self.img_K), dim=0))
            self.register_buffer("img_queue_ptr", torch.zeros(1, dtype=torch.long))
        if txt_encoder is not None:
            self.txt_encoder_q = txt_encoder
            self.txt_encoder_k = self._momentum_copy(txt_encoder)
            self.register_buffer("txt_queue", torch.nn.functional.normalize(torch.randn(self.dim, self.txt_K), dim=0))
            self.register_buffer("txt_queue_ptr", torch.zeros(1, dtype=torch.long))
        self._shadow = {}

    @staticmethod
    def _momentum_copy(enc):
        k = copy.deepcopy(enc).float()  # fp32 master copy (see module docstring)
        for p in k.parameters():
            p.requires_grad = False
        return k

    @torch.no_grad()
    def momentum_update_key_encoder(self):
        for enc_q, enc_k in ((self.img_encoder_q, getattr(self, "img_encoder_k", None)), (self.txt_encoder_q, getattr(self, "txt_encoder_k", None))):
            if enc_q is None:
                continue
            for pq, pk in zip(enc_q.parameters(), enc_k.parameters()):
                ops.ema_update(pk.data, pq.data.contiguous(), self.m)

    def _queue_bf16(self, name):
        """bf16 image of a queue buffer for the kernels (refreshed at enqueue time)."""
        q = getattr(self, name)
        sh = self._shadow.get(name)
        if sh is None or sh.device != q.device:
            sh = q.to(BF16).contiguous()
            self._shadow[name] = sh
        return sh

    def moco_loss(self, pos, neg, mining_top_K=None):
        raise NotImplementedError("b200mm: use moco_loss_fused(q, k_pos, which) — the [N, K] negatives are never materialised")

    def moco_loss_fused(self, q, k_pos, which):
        """which = 'txt' (queries against the text queue) or 'img'."""
        return moco_nce(q, k_pos, self._queue_bf16(which + "_queue"), self.T)

    @torch.no_grad()
    def dequeue_and_enqueue(self, vis_keys, txt_keys):
        def _one(keys, ptr_buf, name, K):
            keys = gather_tensor(keys, method="cat", back_gradient=False, pad_tensors=True)
            if torch.isnan(keys).any().item():
                return
            queue = getattr(self, name)
            ptr = int(ptr_buf)
            end = min(ptr + keys.shape[0], K)
            start = end - keys.shape[0]
            queue[:, start:end] = keys.T.to(queue.dtype)
            # the bf16 image may be held by a pending backward (the reference clones the queue per step, univl_video_ret.py:286,299):
            # never write into it — drop it, the next loss call re-casts the updated queue
            self._shadow.pop(name, None)
            ptr_buf[0] = end % K

        if self.img_encoder_q is not None and vis_keys is not None:
            _one(vis_keys, self.img_queue_ptr, "img_queue", self.img_K)
        if self.txt_encoder_q is not None and txt_keys is not None:
            _one(txt_keys, self.txt_queue_ptr, "txt_queue", self.txt_K)
