"""Image-parallel sharding of the hot path over the GPUs of one node.

The path has no cross-image state at inference (BatchNorm in eval mode), so each rank runs the
whole pipeline on a contiguous slice of the batch with replicated weights; the only collective is
one all_gather of the outputs the caller consumes (para: 916 B / image)."""
import torch
import torch.distributed as dist


def shard_bounds(n, world, rank):
    """Contiguous slice [lo, hi) of n items for `rank`; remainder goes to the low ranks."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t, world=None, rank=None):
    world = dist.get_world_size() if world is None else world
    rank = dist.get_rank() if rank is None else rank
    lo, hi = shard_bounds(t.shape[0], world, rank)
    return t[lo:hi]


def gather_outputs(local, n_total, group=None):
    """All-gather variable-length shards (padded to the largest shard) back into [n_total, ...]."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    per = (n_total + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    pieces = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, world, r)
        pieces.append(out[r * per:r * per + (hi - lo)])
    return torch.cat(pieces, 0)


def infer_sharded(model, images, group=None, infer=None):
    """Every rank passes the same [N,3,224,224] batch (or its own view of it); returns para [N,229]
    on every rank.  `infer` (images -> para) defaults to model.infer_net; an empty shard contributes an empty
    tensor on the MODEL's device (the collective needs one device type on every rank)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(images.shape[0], world, rank)
    local = images[lo:hi]
    if infer is None:
        infer = lambda x: model.infer_net(x)["para"]
    if hi > lo:
        para = infer(local)
    else:
        dev = next(model.parameters()).device if model is not None else images.device
        para = torch.zeros(0, 229, device=dev)
    return gather_outputs(para, images.shape[0], group)
