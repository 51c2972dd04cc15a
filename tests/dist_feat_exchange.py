"""Run under torchrun (1..8 ranks, one GPU each): the fused gather of the rendered features (ray kernel epilogue -> peer stores
into symmetric memory, nerf_loc_b200/distributed.py::FeatExchange) must equal one NCCL all-gather of the local `feat` rows,
bit for bit, and leave every other output unchanged.  Exit code 0 = pass."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerf_loc_b200.distributed import FeatExchange, all_gather_rows, render_rays_sharded, shard_bounds  # noqa: E402
from tests.common import RENDER_CASES, render_inputs  # noqa: E402
from tests.gpu_common import cuda_model, setup_frame  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    name = list(RENDER_CASES)[0]
    S, sd_cpu, sc, scene, ro, rd = render_inputs(name)
    model, sd = cuda_model(S, RENDER_CASES[name][5])
    data = setup_frame(model, sc)
    rays = {"rays_o": ro.cuda(), "rays_d": rd.cuda(), "depth_range": data["depth_range"][0]}
    R = ro.shape[0]
    ex = FeatExchange(R, dev)
    ex.buf.fill_(float("nan"))
    ex.barrier()
    # several frames with a slow consumer on the odd ranks: a faster peer's next frame must not overwrite the gathered matrix this
    # rank is still reading (FeatExchange alternates between two copies)
    frames_ok = True
    subs = [{"rays_o": rays["rays_o"].roll(f, 0), "rays_d": rays["rays_d"].roll(f, 0), "depth_range": rays["depth_range"]} for f in range(4)]
    got = []
    for sub in subs:                           # fused frames back to back: no host sync, no other collective in between
        _, g_fused = render_rays_sharded(model, data, sub, gather=("feat",), exchange=ex)
        if rank % 2 == 1:
            torch.cuda._sleep(200_000_000)     # ~0.1 s on the stream before this rank's consumer reads
        got.append(g_fused["feat"].clone())
    for sub, g in zip(subs, got):
        _, g_ref = render_rays_sharded(model, data, sub, gather=("feat",))
        frames_ok = frames_ok and torch.equal(g, g_ref["feat"])
    torch.cuda.synchronize()
    out_a, g_a = render_rays_sharded(model, data, rays, gather=("feat",))                  # NCCL all-gather
    out_b, g_b = render_rays_sharded(model, data, rays, gather=("feat",), exchange=ex)     # fused peer stores
    torch.cuda.synchronize()
    ok = frames_ok and torch.equal(g_a["feat"], g_b["feat"])
    for k in ("rgb", "depth", "weights", "feat", "depth_uncertainty", "mask"):
        ok = ok and torch.equal(out_a[k], out_b[k])
    lo, hi, _ = shard_bounds(R, world, rank)
    ok = ok and torch.equal(g_b["feat"][lo:hi], out_b["feat"]) and bool(torch.isfinite(g_b["feat"]).all())
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("feat exchange", "ok" if int(flag) == 1 else "MISMATCH", "world", world, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
