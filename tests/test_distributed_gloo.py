"""World-size-2 gloo test (CPU) of the ray-sharding / all-gather logic used by bench.py --gpus N."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nerf_loc_b200.distributed import all_gather_rows, render_rays_sharded, shard_bounds


def test_shard_bounds_cover_everything():
    for R in (0, 1, 5, 8, 307200, 307201):
        for world in (1, 2, 3, 8):
            covered = []
            for r in range(world):
                lo, hi, per = shard_bounds(R, world, r)
                assert 0 <= lo <= hi <= R and hi - lo <= per
                covered += list(range(lo, hi)) if R < 100 else [(lo, hi)]
            if R < 100:
                assert covered == list(range(R))
            else:
                assert covered[0][0] == 0 and covered[-1][1] == R
                assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))


class _FakeModel:
    """Stands in for ConditionalNeRF.render_rays: outputs are a known function of the ray origin."""

    def render_rays(self, data, rays):
        o = rays["rays_o"]
        return {"feat": o[:, :1].repeat(1, 192) * 2.0, "depth": o[:, 0] + 1.0, "mask": o[:, 0] > 2.5}


def _worker(rank, world, port, R, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ro = torch.arange(R, dtype=torch.float32)[:, None].repeat(1, 3)
        rays = {"rays_o": ro, "rays_d": torch.zeros(R, 3)}
        local, full = render_rays_sharded(_FakeModel(), {}, rays, gather=("feat", "depth", "mask"))
        lo, hi, _ = shard_bounds(R, world, rank)
        ok = (local["feat"].shape[0] == hi - lo
              and torch.equal(full["feat"], ro[:, :1].repeat(1, 192) * 2.0)
              and torch.equal(full["depth"], ro[:, 0] + 1.0)
              and torch.equal(full["mask"], ro[:, 0] > 2.5)
              and torch.equal(all_gather_rows(ro[lo:hi], R), ro))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("R", [7, 64])
def test_sharded_render_gathers_all_rays_gloo(R):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, R, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=10) for _ in range(2))
    assert res == {0: True, 1: True}
