"""Ray-count / sample-count sweep of `ConditionalNeRF.render_rays` on one GPU (BASELINE.json configs[4]) and the
Cambridge-shape frame of configs[3] (1920x1056 query, 192 samples/ray, 8 views).

    python tools/sweep.py [--rays 14,16,18] [--samples 64,128,256] [--cambridge RAYS]

Prints one JSON line per point: rays/s on the device (CUDA events, inputs resident, 2 untimed warm-up passes).  Rays are
drawn with replacement from the pixel grid of the synthetic 640x480 query of SURVEY.md 8(d); one model per S (the RayUnet
LayerNorm shapes depend on S).  The Cambridge point renders the first RAYS rays of the 1920x1056 frame spread evenly over the frame; the number is a per-ray
throughput, extrapolated to the 2,027,520-ray frame)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerf_loc_b200 import params, synthetic as syn  # noqa: E402
from nerf_loc_b200.conditional_nerf import ConditionalNeRF  # noqa: E402
from nerf_loc_b200.config import default_args  # noqa: E402


def setup(H, W, V, S, dev):
    sc = syn.make_scene(H, W, V, seed=1234)
    sd = syn.synthetic_state_dict(params.conditional_nerf_shapes(S), 1234)
    model = ConditionalNeRF(default_args(S)).eval()
    model.load_state_dict(sd, strict=False)
    model = model.to(dev)
    data = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc.items() if k != "vis_featmaps"}
    data["scene"], data["filename"] = "synthetic", "sweep"
    model.support_neural_points = None
    model.multiview_aggregator.vis_featmaps = sc["vis_featmaps"].to(dev)
    model.build_support_neural_points(data)
    return sc, model, data


def timed(model, data, ro, rd, reps=2):
    rays = {"rays_o": ro, "rays_d": rd, "depth_range": data["depth_range"][0]}
    for _ in range(2):
        model.render_rays(data, rays)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        model.render_rays(data, rays)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", default="14,16,18,20")
    ap.add_argument("--samples", default="64,128,256")
    ap.add_argument("--cambridge", type=int, default=37888, help="rays of the 1920x1056 / S=192 frame to render (0 = skip)")
    args = ap.parse_args()
    dev = torch.device("cuda")
    for S in [int(s) for s in args.samples.split(",") if s]:
        sc, model, data = setup(480, 640, 8, S, dev)
        px_all = syn.all_pixels(480, 640)
        for lg in [int(x) for x in args.rays.split(",") if x]:
            R = 1 << lg
            g = torch.Generator().manual_seed(lg)
            px = px_all[torch.randint(0, px_all.shape[0], (R,), generator=g)]
            ro, rd = syn.pixel_rays(sc["K"], sc["pose"], px)
            ms = timed(model, data, ro.to(dev), rd.to(dev))
            print(json.dumps({"config": "sweep", "query": "640x480", "views": 8, "samples_per_ray": S, "rays": R,
                              "ms": ms, "rays_per_s": R / (ms * 1e-3), "samples_per_s": R * S / (ms * 1e-3)}), flush=True)
        del model, data
        torch.cuda.empty_cache()
    if args.cambridge > 0:
        H, W, S = 1056, 1920, 192
        sc, model, data = setup(H, W, 8, S, dev)
        px = syn.all_pixels(H, W)
        px = px[::max(1, px.shape[0] // args.cambridge)][:args.cambridge].contiguous()   # evenly spread over the frame
        ro, rd = syn.pixel_rays(sc["K"], sc["pose"], px)
        ms = timed(model, data, ro.to(dev), rd.to(dev), reps=1)
        M = int(model.support_neural_points["fine"]["xyz"].shape[0])
        print(json.dumps({"config": "cambridge-shape", "query": "1920x1056", "views": 8, "samples_per_ray": S,
                          "support_points": M, "rays": int(px.shape[0]), "ms": ms, "rays_per_s": px.shape[0] / (ms * 1e-3),
                          "frame_s_extrapolated": H * W / (px.shape[0] / (ms * 1e-3))}), flush=True)


if __name__ == "__main__":
    main()
