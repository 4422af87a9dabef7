"""gather_tensor — drop-in for antmmf/utils/distributed_utils.py:122-189 on the contrastive path.

One `all_gather_into_tensor` into a contiguous [W*B, ...] buffer (rank order = row order, identical to the reference's
torch.cat of the per-rank list) and, with back_gradient=True, ONE reduce-scatter in backward (the reference issues W
asynchronous reduces, :104-116).

pad_tensors=True (distributed_utils.py:145-176: ranks may hold different numbers of rows, e.g. the last batch with
drop_last=False) exchanges the row counts with one int64 all-gather, pads every rank to the maximum, gathers, and trims
the padding again — same result and same host synchronisation as the reference (the counts are needed on the host to
shape the result); the gradient of the padded rows is dropped in backward. `gathered_sizes(...)` exposes the counts so
that callers can derive row offsets (rank r's rows start at sum(sizes[:r]), not r * B).
"""
import torch
import torch.distributed as dist


def get_rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def get_world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class _AllGatherRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        world = get_world_size()
        x = x.contiguous()
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)
        dist.all_gather_into_tensor(out, x)
        ctx.rows = x.shape[0]
        return out

    @staticmethod
    def backward(ctx, g):
        out = torch.empty((ctx.rows,) + tuple(g.shape[1:]), device=g.device, dtype=g.dtype)
        dist.reduce_scatter_tensor(out, g.contiguous(), op=dist.ReduceOp.SUM)
        return out


class _AllGatherRowsPadded(torch.autograd.Function):
    """Ranks hold different row counts `sizes` (host ints): pad to max, gather, trim; backward = reduce-scatter of the re-padded gradient."""

    @staticmethod
    def forward(ctx, x, sizes):
        world, rank = get_world_size(), get_rank()
        mx = max(sizes)
        x = x.contiguous()
        padded = x.new_zeros((mx,) + tuple(x.shape[1:]))
        padded[: x.shape[0]] = x
        out = torch.empty((world * mx,) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)
        dist.all_gather_into_tensor(out, padded)
        ctx.sizes, ctx.rank = list(sizes), rank
        out = out.view((world, mx) + tuple(x.shape[1:]))
        return torch.cat([out[r, : sizes[r]] for r in range(world)], dim=0)

    @staticmethod
    def backward(ctx, g):
        sizes, rank = ctx.sizes, ctx.rank
        world, mx = len(sizes), max(sizes)
        full = g.new_zeros((world, mx) + tuple(g.shape[1:]))
        o = 0
        for r, n in enumerate(sizes):
            full[r, :n] = g[o : o + n]
            o += n
        out = torch.empty((mx,) + tuple(g.shape[1:]), device=g.device, dtype=g.dtype)
        dist.reduce_scatter_tensor(out, full.view((world * mx,) + tuple(g.shape[1:])), op=dist.ReduceOp.SUM)
        return out[: sizes[rank]], None


def gathered_sizes(tensor):
    """Row count of every rank's `tensor` (list of ints; one int64 all-gather + one host read, as distributed_utils.py:150-158)."""
    world = get_world_size()
    if world < 2:
        return [int(tensor.shape[0])]
    mine = torch.tensor([tensor.shape[0]], dtype=torch.int64, device=tensor.device)
    out = torch.empty(world, dtype=torch.int64, device=tensor.device)
    dist.all_gather_into_tensor(out, mine)
    return [int(v) for v in out.tolist()]


def gather_tensor(tensor, method="stack", back_gradient=False, pad_tensors=False):
    world = get_world_size()
    if world < 2:
        return tensor
    if tensor.ndim == 0:
        if method != "stack":
            raise ValueError('gather_tensor not support 0-dim tensor with method is not "stack"')
        if pad_tensors:
            raise ValueError("gather_tensor not support 0-dim tensor with padding_tensors is True")
        out = torch.empty(world, device=tensor.device, dtype=tensor.dtype)
        dist.all_gather_into_tensor(out, tensor.reshape(1))
        return out
    if pad_tensors:
        sizes = gathered_sizes(tensor)
        if len(set(sizes)) > 1:
            if method == "stack":
                # the reference fails in torch.stack here as well (distributed_utils.py:185-186)
                raise RuntimeError(f"gather_tensor(method='stack'): ranks hold different row counts {sizes}")
            if back_gradient:
                return _AllGatherRowsPadded.apply(tensor, sizes)
            with torch.no_grad():
                return _AllGatherRowsPadded.apply(tensor, sizes)
    if back_gradient:
        flat = _AllGatherRows.apply(tensor)
    else:
        with torch.no_grad():
            flat = _AllGatherRows.apply(tensor)
    return flat.view((world,) + tuple(tensor.shape)) if method == "stack" else flat
