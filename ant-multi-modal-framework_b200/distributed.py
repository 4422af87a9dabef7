"""gather_tensor — drop-in for antmmf/utils/distributed_utils.py:122-189 on the contrastive path.

One `all_gather_into_tensor` into a contiguous [W*B, ...] buffer (rank order = row order, identical to the reference's
torch.cat of the per-rank list) and, with back_gradient=True, ONE reduce-scatter in backward (the reference issues W
asynchronous reduces, :104-116). Shapes are static on this path, so the size exchange + host sync of the reference's
pad_tensors branch (:145-164) is replaced by a check that all ranks agree.
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


def gather_tensor(tensor, method="stack", back_gradient=False, pad_tensors=False):
    world = get_world_size()
    if world < 2:
        return tensor
    if tensor.ndim == 0:
        if method != "stack" or pad_tensors:
            raise ValueError("gather_tensor: 0-dim tensors only support method='stack' without padding")
        out = torch.empty(world, device=tensor.device, dtype=tensor.dtype)
        dist.all_gather_into_tensor(out, tensor.reshape(1))
        return out
    if back_gradient:
        flat = _AllGatherRows.apply(tensor)
    else:
        with torch.no_grad():
            flat = _AllGatherRows.apply(tensor)
    return flat.view((world,) + tuple(tensor.shape)) if method == "stack" else flat
