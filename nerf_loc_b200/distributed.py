"""Ray sharding across the GPUs of one NVSwitch box.

Rays are independent (only the RayUnet couples samples WITHIN a ray), so rank g renders the contiguous slice
[g*per, (g+1)*per) of the row-major ray list; scene tensors and weights are replicated.  The one exchange step is an
all-gather of the per-ray outputs the matcher / caller needs (`feat [R,192]`, optionally rgb / depth / mask).  One process
per GPU, `torch.distributed` (NCCL on GPUs; gloo works for the host-side logic and the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(R, world, rank):
    """(lo, hi, per): this rank's slice of R rays and the padded per-rank row count used by the all-gather."""
    per = (R + world - 1) // world if R > 0 else 0
    lo = min(R, rank * per)
    hi = min(R, lo + per)
    return lo, hi, per


def all_gather_rows(x, R, group=None):
    """x: this rank's rows [hi-lo, ...] -> all R rows on every rank (one collective)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return x
    rank = dist.get_rank(group)
    lo, hi, per = shard_bounds(R, world, rank)
    assert x.shape[0] == hi - lo, "shard size mismatch"
    if x.shape[0] < per:
        x = torch.cat([x, x.new_zeros((per - x.shape[0],) + tuple(x.shape[1:]))])
    x = x.contiguous()
    out = x.new_empty((per * world,) + tuple(x.shape[1:]))
    dist.all_gather_into_tensor(out, x, group=group)
    return out[:R]


def render_rays_sharded(model, data, rays, gather=("feat",), group=None):
    """`ConditionalNeRF.render_rays` over this rank's slice of `rays`, then one all-gather per requested output.
    Returns (local_outputs, gathered_outputs)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    R = rays["rays_o"].shape[0]
    lo, hi, _ = shard_bounds(R, world, rank)
    local = dict(rays)
    local["rays_o"], local["rays_d"] = rays["rays_o"][lo:hi], rays["rays_d"][lo:hi]
    if "pixel_coordinates" in rays:
        local["pixel_coordinates"] = rays["pixel_coordinates"][lo:hi]
    out = model.render_rays(data, local)
    gathered = {}
    for k in gather:
        v = out[k]
        g = all_gather_rows(v.to(torch.uint8) if v.dtype == torch.bool else v, R, group)
        gathered[k] = g.bool() if v.dtype == torch.bool else g
    return out, gathered
