"""Two-pass micro-batching of the dual-encoder contrastive step (GradCache form) — SURVEY.md §8(f) rank 1.

The reference reaches large contrastive batches only through `gradient_accumulation_steps` (antmmf/trainers/base_trainer.py:392-395),
which shrinks the set of negatives to the micro-batch, or through `grad_checkpointing` (clip/modeling_bert.py:461-467). This driver
keeps the FULL batch of negatives at micro-batch activation memory:

  pass 1   for every micro-batch: embeddings under no_grad                                   -> [B, E] per tower (bf16, tiny)
  loss     loss_fn(*embeddings) on leaf copies; backward gives d loss / d embedding           -> [B, E] per tower
           (the all-gather / reduce-scatter of the fused contrastive kernels lives inside loss_fn, so the negatives are global)
  pass 2   for every micro-batch: re-encode with autograd, backward(embedding-gradient slice)  -> parameter gradients accumulate

The result is the exact gradient of the full-batch loss (the towers are per-sample functions: no batch statistics; dropout masks are replayed in pass 2),
at the price of one extra forward. The driver itself is tower-agnostic and runs on any device (CPU tests use toy towers); the B200
path enters through the encoders and the loss handed to it.
"""
import contextlib
from typing import Callable, Sequence

import torch


def _chunks(n: int, size: int):
    return [(s, min(s + size, n)) for s in range(0, n, size)]


class GradCache:
    """encoders[i](inputs[i][s:e]) -> [e-s, E_i];   loss_fn(emb_0, emb_1, ...) -> scalar.

    `sync_modules`: DistributedDataParallel wrappers whose gradient all-reduce must fire only once — every backward of pass 2 except
    the last runs under their `no_sync()`.
    """

    def __init__(self, encoders: Sequence[Callable], loss_fn: Callable, micro_batch: int, sync_modules: Sequence = ()):
        if micro_batch <= 0:
            raise ValueError("GradCache: micro_batch must be positive")
        self.encoders = list(encoders)
        self.loss_fn = loss_fn
        self.micro_batch = int(micro_batch)
        self.sync_modules = [m for m in sync_modules if hasattr(m, "no_sync")]

    def _no_sync(self):
        stack = contextlib.ExitStack()
        for m in self.sync_modules:
            stack.enter_context(m.no_sync())
        return stack

    def step(self, *inputs, loss_scale: float = 1.0) -> torch.Tensor:
        """Accumulates d(loss_scale * loss)/d(theta) into the `.grad`s of everything the encoders and loss_fn touch; returns the
        detached (unscaled) loss."""
        if len(inputs) != len(self.encoders):
            raise ValueError(f"GradCache: {len(self.encoders)} encoders but {len(inputs)} inputs")
        n = inputs[0].shape[0]
        if any(x.shape[0] != n for x in inputs):
            raise ValueError("GradCache: all towers must see the same number of samples")
        spans = _chunks(n, self.micro_batch)
        # pass 1: embeddings only
        # (dropout: each micro-batch forward of pass 1 records the seed-sequence position it started from, pass 2 replays it, so both
        # passes see the same masks — the counter-based analogue of GradCache's RNG-state snapshots)
        from . import ops

        drop_states = {}
        with torch.no_grad():
            embs = []
            for t, (enc, x) in enumerate(zip(self.encoders, inputs)):
                outs = []
                for s, e in spans:
                    drop_states[(t, s)] = ops.get_dropout_state()
                    outs.append(enc(x[s:e]))
                embs.append(torch.cat(outs))
        drop_end = ops.get_dropout_state()
        leaves = [e.detach().requires_grad_() for e in embs]
        # loss on the full batch; parameters used directly by the loss (logit_scale) get their gradient here
        with self._no_sync():
            loss = self.loss_fn(*leaves)
            (loss * loss_scale).backward()
        d_embs = [leaf.grad for leaf in leaves]
        # pass 2: rebuild each micro-batch graph and push the cached embedding gradient through it
        jobs = [(t, s, e) for t in range(len(self.encoders)) for s, e in spans]
        for i, (t, s, e) in enumerate(jobs):
            ctx = contextlib.nullcontext() if i == len(jobs) - 1 else self._no_sync()
            with ctx:
                ops.set_dropout_state(drop_states[(t, s)])
                out = self.encoders[t](inputs[t][s:e])
                out.backward(d_embs[t][s:e].to(out.dtype))
        ops.set_dropout_state(drop_end)
        return loss.detach()


def allreduce_grads(params, group=None, bucket_bytes: int = 256 << 20) -> None:
    """Average the `.grad`s over the ranks of `group` in flat buckets (what DDP's reducer does; used here because pass 2 calls the
    towers' methods directly, below any DistributedDataParallel wrapper)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    world = dist.get_world_size(group)
    grads = [p.grad for p in params if p.grad is not None]
    by_dtype = {}
    for g in grads:
        by_dtype.setdefault(g.dtype, []).append(g)
    for gs in by_dtype.values():
        bucket, size = [], 0
        for g in gs + [None]:
            if g is not None and (size == 0 or size + g.numel() * g.element_size() <= bucket_bytes):
                bucket.append(g)
                size += g.numel() * g.element_size()
                continue
            flat = torch.cat([b.reshape(-1) for b in bucket])
            dist.all_reduce(flat, group=group)
            flat.div_(world)
            off = 0
            for b in bucket:
                b.copy_(flat[off:off + b.numel()].view_as(b))
                off += b.numel()
            bucket, size = ([g], g.numel() * g.element_size()) if g is not None else ([], 0)


def cnclip_gradcache_step(model, image: torch.Tensor, text: torch.Tensor, micro_batch: int, group=None, loss_scale: float = 1.0,
                          sync_grads: bool = True) -> torch.Tensor:
    """One full-batch-negatives training step of a b200mm CNCLIP at micro-batch memory (cfg 1/2/4 of BASELINE.json).
    `model` is the bare CNCLIP (not a DDP wrapper): with more than one rank the parameter gradients are averaged here."""
    from . import functional as Fn
    from .contrastive import clip_contrastive_loss

    gc = GradCache(
        [lambda x: Fn.RowNormFn.apply(model.encode_image(x)), lambda t: Fn.RowNormFn.apply(model.encode_text(t))],
        lambda i, t: clip_contrastive_loss(i, t, model.logit_scale, group),
        micro_batch,
    )
    loss = gc.step(image, text, loss_scale=loss_scale)
    if sync_grads:
        allreduce_grads(list(model.parameters()), group)
    return loss
