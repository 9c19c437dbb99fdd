"""One-process-per-GPU sharding of the two phases of the path.

* render / pose loop: the N hypotheses split contiguously over ranks; ``z_obj`` and weights are
  replicated; hypotheses are independent (reference pose/estimation.py:582-594) so there is NO
  per-iteration data-path collective — only ``merge_rankings`` (a tiny all_gather of [n] losses and
  camera parameters) when a global ranking is wanted.
* reconstruction: the V reference views split over ranks (the axis the reference's single-process
  ``MyDataParallel`` scatters, recon/models.py:248-251).  Associative fusers (pool:mean / pool:max) reduce
  locally and ``all_reduce``; the others (GRU/LSTM recurrences, median, abs_max) ``all_gather`` the
  per-view cubes and fuse replicated.  Either way every rank ends with the fused latent volume in place
  for its render shard.

The collectives go through ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced [lo, hi) slice of n items for `rank` (first n % world ranks get one extra)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_ragged(local, sizes, group=None):
    """Concatenate per-rank tensors whose dim-0 sizes are `sizes` (known on every rank)."""
    world = len(sizes)
    if world == 1:
        return local
    pad = max(sizes)
    buf = local.new_zeros((pad, *local.shape[1:]))
    buf[:local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf.contiguous(), group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


def fuse_views_sharded(local_views, pool_type, num_views, rank, world, group=None):
    """View-axis pooling when each rank holds `local_views` [v_local, ...] of `num_views` total.
    mean / max reduce locally then all_reduce; anything else gathers the views first."""
    if world == 1:
        gathered = local_views
    elif pool_type == 'mean':
        part = local_views.sum(dim=0, keepdim=True)
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
        return part / num_views
    elif pool_type == 'max':
        part = (local_views.max(dim=0, keepdim=True)[0] if local_views.shape[0]
                else local_views.new_full((1, *local_views.shape[1:]), float('-inf')))
        dist.all_reduce(part, op=dist.ReduceOp.MAX, group=group)
        return part
    else:
        sizes = [shard_range(num_views, r, world)[1] - shard_range(num_views, r, world)[0] for r in range(world)]
        gathered = all_gather_ragged(local_views, sizes, group)
    return gathered


def merge_rankings(local_losses, local_params, ranking_size, group=None):
    """Global top-`ranking_size` over all ranks' hypotheses: returns (losses [k], params [k, P], owner rank [k])."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        order = torch.argsort(local_losses)[:ranking_size]
        return local_losses[order], local_params[order], torch.zeros_like(order)
    n = torch.tensor([local_losses.shape[0]], device=local_losses.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    sizes = [int(c.item()) for c in counts]
    losses = all_gather_ragged(local_losses.reshape(-1, 1), sizes, group).reshape(-1)
    params = all_gather_ragged(local_params, sizes, group)
    owner = torch.cat([torch.full((s,), r, dtype=torch.long, device=losses.device) for r, s in enumerate(sizes)])
    order = torch.argsort(losses)[:ranking_size]
    return losses[order], params[order], owner[order]


class _AllGatherViews(torch.autograd.Function):
    """Differentiable all-gather along `dim` of equally sized shards: forward concatenates the ranks' shards in rank
    order, backward hands every rank the SUM over ranks of the gradient slice that belongs to its shard (each rank
    back-propagates its own share of the loss through a replica of whatever consumed the gathered tensor)."""

    @staticmethod
    def forward(ctx, local, dim, group):
        world = dist.get_world_size(group)
        ctx.dim, ctx.group, ctx.world, ctx.rank = dim, group, world, dist.get_rank(group)
        x = local.movedim(dim, 0).contiguous()
        out = x.new_empty((world * x.shape[0], *x.shape[1:]))
        dist.all_gather_into_tensor(out, x, group=group)
        ctx.n = x.shape[0]
        return out.movedim(0, dim)

    @staticmethod
    def backward(ctx, g):
        g = g.movedim(ctx.dim, 0).contiguous()
        if dist.get_backend(ctx.group) == 'nccl':
            part = g.new_empty((ctx.n, *g.shape[1:]))
            dist.reduce_scatter_tensor(part, g, op=dist.ReduceOp.SUM, group=ctx.group)
        else:                                   # gloo has no reduce-scatter: all-reduce, keep the own slice
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
            part = g[ctx.rank * ctx.n:(ctx.rank + 1) * ctx.n]
        return part.movedim(0, ctx.dim), None, None


def all_gather_views(local, dim=1, group=None):
    """autograd-aware all-gather of per-rank view shards (identity when not distributed / world 1)"""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    return _AllGatherViews.apply(local, dim, group)


def allreduce_gradients(parameters, group=None):
    """DDP-style: SUM every parameter gradient over the ranks in one flat bucket (each rank's loss is already its
    share of the global mean, so the sum is the gradient of the global loss).  Parameters a rank did not touch
    contribute zeros."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    params = [p for p in parameters if p.requires_grad]
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p in params:
        n = p.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n


@torch.no_grad()
def build_latent_object_sharded(model, cameras, color, mask, rank=0, world=1, group=None, depth=None):
    """Sculptor.encode with the reference views sharded over ranks.  `cameras` (V zoomed cameras),
    `color` [1,V,3,H,W] and `mask` [1,V,1,H,W] are the full (replicated) host inputs; each rank encodes its
    slice on its own GPU.  Returns the fused z_obj [1,1,C,S,S,S] on every rank."""
    from .recon import fusion
    from .recon.models import gan_normalize
    dev = model.device
    sc, fuser = model.sculptor, model.fuser
    if isinstance(fuser, fusion.BlendFuser):
        raise NotImplementedError("view-sharded reconstruction does not cover the BlendFuser (it needs the camera-space "
                                  "intermediates of every view); use Sculptor.encode")
    V = color.shape[1]
    lo, hi = shard_range(V, rank, world)
    # input planes exactly as Sculptor.encode assembles them (recon/models.py:226-246)
    planes = []
    if sc.input_color:
        planes.append(color[0, lo:hi])
    if sc.input_depth:
        if depth is None:
            raise ValueError("this Sculptor takes a depth plane (input_depth=True)")
        planes.append(depth[0, lo:hi])
    if sc.input_mask:
        planes.append(gan_normalize(mask[0, lo:hi]))
    x = torch.cat(planes, dim=1).to(dev)
    z_local, _, _ = sc(x, cameras[lo:hi].to(dev))                                 # [v_local, C, S, S, S]
    if isinstance(fuser, fusion.PoolFuser) and fuser.pool_type in ('mean', 'max') and world > 1:
        return fuse_views_sharded(z_local, fuser.pool_type, V, rank, world, group).unsqueeze(0)
    if world > 1 and V % world == 0:
        # equal shards: one all_gather_into_tensor straight into the [V, C, S^3] buffer the recurrence reads
        z_all = z_local.new_empty((V, *z_local.shape[1:]))
        dist.all_gather_into_tensor(z_all, z_local.contiguous(), group=group)
    else:
        z_all = fuse_views_sharded(z_local, 'gather', V, rank, world, group)
    z, _ = fuser(z_all.unsqueeze(0), [], [], cameras.to(dev))
    return z
