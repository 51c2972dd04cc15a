"""Ray sharding across the GPUs of one NVSwitch box.

Rays are independent (only the RayUnet couples samples WITHIN a ray), so rank g renders the contiguous slice
[g*per, (g+1)*per) of the row-major ray list; scene tensors and weights are replicated.  The one exchange step is an
all-gather of the per-ray outputs the matcher / caller needs (`feat [R,192]`, optionally rgb / depth / mask).  One process
per GPU, `torch.distributed` (NCCL on GPUs; gloo works for the host-side logic and the CPU tests).

On GPUs the exchange of `feat` is FUSED into the render: `FeatExchange` allocates the gathered [R,192] matrix in symmetric
memory (every rank's copy is mapped into every other rank), and the ray kernel's epilogue stores each rendered feature row
straight into all copies with peer stores over NVLink (`nlb_render_rays_gather`).  What is left of the collective is one
device-side barrier.  `all_gather_rows` (one NCCL / gloo all-gather) remains for the other outputs and for the CPU tests.
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


class FeatExchange:
    """Gathered rendered-feature matrix [R,192] in symmetric memory + the peer pointers the ray kernels store into.

    Two copies, used alternately (`begin_frame`): the peers' ray kernels of frame n+1 store into the copy frame n did NOT use, so
    a rank that is still reading frame n's gathered matrix (its matcher) is never overwritten by a faster peer.  Frame n+2
    reuses frame n's copy, and a peer can only get there through the barrier of frame n+1, which this rank joins on the same
    stream AFTER its frame-n consumers.  Contract: per frame `begin_frame()` -> render with the returned pointers -> `barrier()`
    -> read `gathered()` on the stream the barrier ran on; the view is valid until the second `begin_frame()` after it."""

    def __init__(self, R, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.R = R
        self.rows = max(R, 1)
        self.buf = symm_mem.empty(2 * self.rows, 192, dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        self.base = [int(p) for p in self.handle.buffer_ptrs]
        if len(self.base) > 8:
            raise RuntimeError("FeatExchange: at most 8 ranks (one NVSwitch box)")
        self.parity = 1
        self.ptrs = None

    def begin_frame(self):
        """Flips to the other copy; returns the peer pointers the ray kernels of this frame store into."""
        self.parity ^= 1
        off = self.parity * self.rows * 192 * 4
        self.ptrs = [p + off for p in self.base]
        return self.ptrs

    def barrier(self):
        """All ranks' peer stores queued before this point (on the current stream) are visible after it."""
        self.handle.barrier()

    def gathered(self):
        lo = self.parity * self.rows
        return self.buf[lo:lo + self.R]


def render_rays_sharded(model, data, rays, gather=("feat",), group=None, exchange=None):
    """`ConditionalNeRF.render_rays` over this rank's slice of `rays`, then one all-gather per requested output.
    With `exchange` (a FeatExchange over all R rays) the gather of `feat` is done by the ray kernel itself (peer stores) and
    only a barrier follows.  Returns (local_outputs, gathered_outputs)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    R = rays["rays_o"].shape[0]
    lo, hi, _ = shard_bounds(R, world, rank)
    local = dict(rays)
    local["rays_o"], local["rays_d"] = rays["rays_o"][lo:hi], rays["rays_d"][lo:hi]
    if "pixel_coordinates" in rays:
        local["pixel_coordinates"] = rays["pixel_coordinates"][lo:hi]
    if exchange is not None:
        out = model.render_rays(data, local, _feat_peers=(exchange.begin_frame(), lo))
        exchange.barrier()
    else:
        out = model.render_rays(data, local)
    gathered = {}
    for k in gather:
        if k == "feat" and exchange is not None:
            gathered[k] = exchange.gathered()
            continue
        v = out[k]
        g = all_gather_rows(v.to(torch.uint8) if v.dtype == torch.bool else v, R, group)
        gathered[k] = g.bool() if v.dtype == torch.bool else g
    return out, gathered
