#!/usr/bin/env python
"""Benchmark of the NeRF-Loc render hot path on B200 (BASELINE.json: rays/sec, 640x480, 128 samples/ray, 8 views).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm's CPU port (oracle/) on host cores

A step = one pass of `ConditionalNeRF.render_rays` over every pixel ray of the synthetic 640x480 query frame
(configs[1]).  With N GPUs the rays of the SAME frame are split into N contiguous slices (strong scaling) and the
rendered per-ray features are gathered on every rank once per step, as the matcher would need them - by the ray kernel's own
epilogue (peer stores over NVLink into symmetric memory, nerf_loc_b200/distributed.py::FeatExchange) plus one barrier.
Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, V, S = 480, 640, 8, 128
F_SAMPLE = 2880640 + 38912 * V          # algorithmic matmul+conv FLOP per sample point (SURVEY.md 8d / BASELINE.md 3)
# split of F_SAMPLE over the kernels, per sample (BASELINE.md section 3 breakdown)
F_KERNEL = {"aggregate": 201e3 + 176e3 * V / 8, "neighbor": 1108e3 + 1081e3 + 264e3, "ray": 270e3 + 82e3 + 9e3, "knn": 0.0}


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture (profiles/), bytes;
# None until a capture of the current build exists
# profiles/r1i_ncu_metrics.json: one launch each on 37,888 rays; kept per ray and scaled to the rays one launch of the timed run covers
TRAFFIC_RAYS = 37888
TRAFFIC = {"aggregate": 0.041086e9 + 8.021647e9, "neighbor": 2.832573e9 + 2.443556e9, "ray": 10.319115e9 + 0.053280e9,
           "knn": 0.004504e9 + 0.320801e9}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["bf16_tflops_sustained"], d["hbm_gbs"], "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in o.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx) if mx else None, "reasons": reasons}


def build_frame():
    """Synthetic frame of SURVEY.md section 8(d) on the CPU (seed 1234) + synthetic weights."""
    from nerf_loc_b200 import params, synthetic as syn
    sc = syn.make_scene(H, W, V, seed=1234)
    sd = syn.synthetic_state_dict(params.conditional_nerf_shapes(S), 1234)
    px = syn.all_pixels(H, W)
    ro, rd = syn.pixel_rays(sc["K"], sc["pose"], px)
    return sc, sd, ro, rd


def oracle_setup(sc, sd):
    from oracle import nerfloc_oracle as O
    scene = dict(Ks=sc["topk_Ks"], c2ws=sc["topk_poses"], images=sc["topk_images"], vis_maps=sc["vis_featmaps"],
                 depth_range=sc["depth_range"][0])
    with torch.no_grad():
        # only the fine level is needed for rendering; confidence runs the aggregator on every support point
        ff, xf, nf, df = O.backproject_support_frame(sc["topk_images"], sc["feat_fine_src"], sc["topk_depths"],
                                                     sc["topk_Ks"], sc["topk_poses"], 4)
        conf = []
        for s in range(0, xf.shape[0], 16384):
            agg, _, _ = O.aggregator_forward(sd, "multiview_aggregator", xf[s:s + 16384], scene["Ks"], scene["c2ws"],
                                             scene["images"], sc["feat_fine_src"].permute(0, 3, 1, 2),
                                             scene["vis_maps"], scene["depth_range"])
            h = torch.nn.functional.leaky_relu(torch.nn.functional.linear(agg, sd["confidence_mlp.0.weight"], sd["confidence_mlp.0.bias"]), 0.01)
            conf.append(torch.sigmoid(torch.nn.functional.linear(h, sd["confidence_mlp.2.weight"], sd["confidence_mlp.2.bias"])))
    return scene, {"xyz": xf, "feature": ff, "confidence": torch.cat(conf), "direction": df}


def oracle_render(sc, sd, scene, sup, ro, rd):
    from oracle import knn_oracle as KO
    from oracle import nerfloc_oracle as O
    knn = lambda a, b, K: KO.knn_c(a, b, K)  # exact, all host threads
    with torch.no_grad():
        return O.render_rays(sd, scene, sup, sc["feat_fine_src"].permute(0, 3, 1, 2), ro, rd, sc["pose"], S, knn=knn)


def cpu_sample(ro, rd, n):
    idx = torch.linspace(0, ro.shape[0] - 1, n).long()
    return ro[idx].contiguous(), rd[idx].contiguous(), idx


def matcher_frame(N3=4096, hc=60, wc=80, seed=1234):
    """Synthetic matching inputs of SURVEY.md section 8(d): 4096 3D descriptors vs the 60x80 coarse / 120x160 fine maps of
    a 640x480 query, 2048 planted pairs."""
    import math
    g = torch.Generator().manual_seed(seed)
    Mc = hc * wc
    d2 = torch.randn(Mc, 192, generator=g)
    d3 = torch.randn(N3, 192, generator=g)
    P = min(N3 // 2, Mc)
    cells = torch.randperm(Mc, generator=g)[:P]
    d3[:P] = d2[cells] + 0.05 * torch.randn(P, 192, generator=g)
    kps3d = torch.rand(N3, 3, generator=g) * 2
    freqs = 2.0 ** torch.linspace(0.0, 31.0, steps=32)
    pe3 = torch.cat([f(kps3d * fr) for fr in freqs for f in (torch.sin, torch.cos)], -1)
    ys = (torch.arange(1, hc + 1, dtype=torch.float32) - 0.5) / (hc + 1e-6)
    xs = (torch.arange(1, wc + 1, dtype=torch.float32) - 0.5) / (wc + 1e-6)
    p = torch.stack([xs[None, :].expand(hc, wc), ys[:, None].expand(hc, wc)], -1)
    bases = [i + 1 for i in range(48)]
    pe2 = torch.cat([torch.sin(i * math.pi * p) for i in bases] + [torch.cos(i * math.pi * p) for i in bases], -1).reshape(Mc, -1)
    gy, gx = torch.meshgrid(torch.arange(hc), torch.arange(wc), indexing="ij")
    return dict(desc_3d=d3, pos_emd_3d=pe3, desc_2d_coarse=d2, pos_emd_2d=pe2, kps3d=kps3d,
                kps2d=torch.stack([gx, gy], -1).view(-1, 2).float(),
                feat_fine=torch.randn(1, hc * 2, wc * 2, 192, generator=g), feat_coarse=torch.randn(1, hc, wc, 192, generator=g),
                desc_3d_fine=torch.randn(N3, 192, generator=g), stride_coarse=8, stride_fine=4)


def bench_matcher(dev, cpu_n3):
    """Match ms/frame (BASELINE.json metric, second half): Matcher.forward at configs[2] size on the device, and the CPU port on a
    bounded sample (fewer 3D points, same 2D maps)."""
    from nerf_loc_b200 import params, synthetic as syn
    from nerf_loc_b200.config import default_args
    from nerf_loc_b200.matcher import Matcher
    sd = syn.synthetic_state_dict(params.matcher_shapes(), 99)
    m = Matcher(default_args(), 192, 192, 192).eval()
    m.load_state_dict(sd)
    m = m.to(dev)
    data = matcher_frame()
    dd = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in data.items()}
    with torch.no_grad():
        for _ in range(2):
            out = m(dict(dd))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            out = m(dict(dd))
        e1.record()
        torch.cuda.synchronize()
        # the S2D kernel alone
        d3t, d2t = m.coarse_transformer(dd["desc_3d"][None], dd["pos_emd_3d"][None], dd["desc_2d_coarse"][None], dd["pos_emd_2d"][None])
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(3):
            m.s2d(d3t[0], d2t[0], 0.2)
        f1.record()
        torch.cuda.synchronize()
    res = {"ms_per_frame": e0.elapsed_time(e1) / 3, "s2d_ms": f0.elapsed_time(f1) / 3, "n3": 4096, "mc": 4800,
           "matches": int(out["i_ids"].numel()),
           "s2d_tflops_algorithmic": 82368.0 * 4096 * 4800 / (f0.elapsed_time(f1) / 3 * 1e-3) / 1e12}
    if cpu_n3 > 0:
        from oracle import matcher_oracle as MO
        small = matcher_frame(N3=cpu_n3)
        with torch.no_grad():
            t0 = time.perf_counter()
            ref = MO.matcher_forward(sd, small)
            dt = time.perf_counter() - t0
            got = m({k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in small.items()})
        res["cpu_port"] = {"ms": dt * 1e3, "n3": cpu_n3, "mc": 4800, "cores": torch.get_num_threads(),
                           "sample": f"oracle/ matcher_forward with {cpu_n3} of the 4096 3D points, full 60x80 / 120x160 maps"}
        res["parity_on_sample"] = {"score_matrix": float((got["score_matrix"].cpu() - ref["score_matrix"]).abs().max() / ref["score_matrix"].abs().max())}
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    sc, sd, ro, rd = build_frame()
    scene, sup = oracle_setup(sc, sd)
    n = args.cpu_rays
    ro_s, rd_s, _ = cpu_sample(ro, rd, n)
    for _ in range(args.warmup):
        oracle_render(sc, sd, scene, sup, ro_s[:32], rd_s[:32])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_render(sc, sd, scene, sup, ro_s, rd_s)
    dt = (time.perf_counter() - t0) / args.steps
    val = n / dt
    line = {
        "impl": "reference", "metric": "rays/sec", "value": val, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "640x480 query, 128 samples/ray, 8 ref views, full conditional render (configs[1])",
                   "rays_per_step": n},
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{n} evenly spaced rays of the frame x {S} samples, oracle/ render_rays incl. exact KNN"},
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch.distributed as dist
    from nerf_loc_b200 import _lib
    from nerf_loc_b200.conditional_nerf import ConditionalNeRF
    from nerf_loc_b200.config import default_args

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (nerf_loc_b200 has no CPU path); use --impl reference for the CPU port")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.load()

    sc, sd, ro, rd = build_frame()
    R_total = ro.shape[0] if args.rays <= 0 else min(args.rays, ro.shape[0])
    ro, rd = ro[:R_total], rd[:R_total]
    per = (R_total + world - 1) // world
    lo, hi = rank * per, min(R_total, (rank + 1) * per)
    model = ConditionalNeRF(default_args(S)).eval()
    model.load_state_dict(sd, strict=False)
    model = model.to(dev)
    model.chunk_rays = args.chunk
    data = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc.items() if k != "vis_featmaps"}
    data["scene"], data["filename"] = "synthetic", "bench"
    model.support_neural_points = None
    model.multiview_aggregator.vis_featmaps = sc["vis_featmaps"].to(dev)
    t_setup0 = time.perf_counter()
    model.build_support_neural_points(data)
    model._level_scene(data, "fine", query_pose=data["pose"])
    torch.cuda.synchronize()
    setup_ms = (time.perf_counter() - t_setup0) * 1e3

    ro_h, rd_h = ro[lo:hi].contiguous().pin_memory(), rd[lo:hi].contiguous().pin_memory()
    ro_d, rd_d = ro_h.to(dev), rd_h.to(dev)
    Rl = hi - lo
    from nerf_loc_b200.distributed import FeatExchange
    # the one exchange step, fused into the render: every rank's ray kernel stores its feature rows into all
    # ranks' gathered [R,192] matrix (symmetric memory, peer stores over NVLink); a device-side barrier closes the step
    exch, exch_note = None, ""
    if world > 1:
        # agree on the path BEFORE the collective allocation: a rank that failed alone inside the rendezvous would hang the rest
        try:
            __import__("torch.distributed._symmetric_memory")
            ok = torch.ones(1, device=dev)
        except Exception as e:  # symmetric memory not available in this build: one NCCL all-gather per step instead
            ok, exch_note = torch.zeros(1, device=dev), f"{type(e).__name__}"
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)   # all ranks take the same path
        if float(ok.item()) != 0.0:
            exch = FeatExchange(R_total, dev)
    from nerf_loc_b200.distributed import all_gather_rows

    def step_device():
        rays = {"rays_o": ro_d, "rays_d": rd_d, "depth_range": data["depth_range"][0]}
        if exch is None:
            out = model.render_rays(data, rays)
            if world > 1:
                out["feat_all"] = all_gather_rows(out["feat"], R_total)
            return out
        out = model.render_rays(data, rays, _feat_peers=(exch.begin_frame(), lo))
        exch.barrier()
        out["feat_all"] = exch.gathered()
        return out

    host_out = {}

    def step_e2e():
        rays = {"rays_o": ro_h.to(dev, non_blocking=True), "rays_d": rd_h.to(dev, non_blocking=True),
                "depth_range": data["depth_range"][0]}
        if exch is None:
            out = model.render_rays(data, rays)
            if world > 1:
                all_gather_rows(out["feat"], R_total)
        else:
            out = model.render_rays(data, rays, _feat_peers=(exch.begin_frame(), lo))
            exch.barrier()
        nbytes = 0
        for k, v in out.items():
            if k not in host_out:
                host_out[k] = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
            host_out[k].copy_(v, non_blocking=True)
            nbytes += v.numel() * v.element_size()
        torch.cuda.current_stream().synchronize()
        return nbytes

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    sampler.start()
    total_ms = timed(step_device, args.steps)
    sampler.stop_flag = True
    # end-to-end through the public API with host buffers
    step_e2e()
    e2e_ms = timed(step_e2e, max(1, min(args.steps, 3)))
    e2e_steps = max(1, min(args.steps, 3))
    d2h = sum(v.numel() * v.element_size() for v in host_out.values())
    # per-kernel device time (one extra, untimed-for-value pass with events around every launch)
    L.nlb_profile_enable(1)
    step_device()
    torch.cuda.synchronize()
    ms4 = (ctypes.c_double * 4)()
    n4 = (ctypes.c_int64 * 4)()
    L.nlb_profile_read(ms4, n4, 4)
    L.nlb_profile_enable(0)
    names = ["knn", "aggregate", "neighbor", "ray"]
    kern = {n: {"ms": ms4[i], "launches": int(n4[i])} for i, n in enumerate(names)}
    dom = max(names, key=lambda n: kern[n]["ms"])
    tf_peak, hbm_peak, which = peaks()
    samples_rank = Rl * S
    ach = F_KERNEL[dom] * samples_rank / (kern[dom]["ms"] * 1e-3) / 1e12 if kern[dom]["ms"] > 0 else 0.0

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_step = total_ms / args.steps
    value = R_total / (ms_step * 1e-3)
    line = {
        "metric": "rays/sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "640x480 query, 128 samples/ray, 8 ref views, full conditional render (configs[1])",
                   "rays_per_step": R_total, "samples_per_ray": S, "views": V, "support_points": int(model.support_neural_points["fine"]["xyz"].shape[0]),
                   "chunk_rays": args.chunk,
                   "mma_mode": "3xTF32 tcgen05 (neighbour MLP with the A operand in tensor memory, RayUnet, feat/blend layers) + 3xTF32 mma.sync (visibility decoder, q / q~ projections) + fp32 FFMA2 (rest of the aggregator, small per-sample GEMMs)",
                   "l2": "working set per step (scene 294 MB + >1 GB of per-chunk intermediates) exceeds the 126 MB L2",
                   "parallelism": f"ray-shard x{world}" + ("" if world == 1 else (" + all-gather of feat[R,192] fused into the ray kernel epilogue (peer stores over NVLink, symmetric memory)" if exch is not None else " + NCCL all-gather of feat[R,192] (symmetric memory unavailable: " + exch_note + ")")),
                   "per_frame_setup_ms": setup_ms},
        "e2e": {"value": R_total / (e2e_ms / e2e_steps * 1e-3), "unit": "rays/s",
                "h2d_bytes_per_step": int(ro_h.numel() * 4 * 2), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(L.nlb_render_launch_count(Rl, args.chunk)) * args.steps,
        "clocks": sampler.summary(),
        "kernels_ms_per_step": {n: kern[n]["ms"] for n in names},
        "roofline": {"bound": "tensor", "kernel": dom + "_kernel", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s",
                     "frac": ach / tf_peak,
                     "traffic": TRAFFIC[dom] / TRAFFIC_RAYS * Rl / max(1, kern[dom]["launches"]), "peak_source": which + " bf16 sustained",
                     "whole_step_achieved": F_SAMPLE * R_total * S / (ms_step * 1e-3) / 1e12,
                     "per_kernel": {n: {"achieved": (F_KERNEL[n] * samples_rank / (kern[n]["ms"] * 1e-3) / 1e12) if kern[n]["ms"] > 0 else 0.0,
                                        "frac": (F_KERNEL[n] * samples_rank / (kern[n]["ms"] * 1e-3) / 1e12 / tf_peak) if kern[n]["ms"] > 0 else 0.0}
                                    for n in names if F_KERNEL[n] > 0}},
    }
    if args.cpu_rays > 0 and world == 1:
        torch.set_num_threads(os.cpu_count() or 1)
        scene, sup = oracle_setup(sc, sd)
        ro_s, rd_s, idx = cpu_sample(ro, rd, args.cpu_rays)
        oracle_render(sc, sd, scene, sup, ro_s[:16], rd_s[:16])
        t0 = time.perf_counter()
        ref = oracle_render(sc, sd, scene, sup, ro_s, rd_s)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": args.cpu_rays / dt, "unit": "rays/s", "cores": torch.get_num_threads(),
                                "kind": "port",
                                "sample": f"{args.cpu_rays} evenly spaced rays of the frame x {S} samples (oracle/ render_rays, exact KNN on all host threads)"}
        # parity of the benchmarked frame on that sample (checker only).  Both paths get the SAME per-frame inputs:
        # the support points are injected from the CPU setup, because a 1-ulp difference between a GPU and a CPU
        # back-projection flips near-tied nearest neighbours (the reference itself is discontinuous there).
        model.support_neural_points = {"fine": {k: v.to(dev) for k, v in sup.items()}, "coarse": None}
        out = step_device()
        err = {k: float((out[k][idx.to(dev)].cpu() - ref[k]).abs().max() / ref[k].abs().max()) for k in ("rgb", "depth", "feat", "weights")}
        line["parity_on_sample"] = err
        line["match"] = bench_matcher(dev, args.cpu_match_n3)
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rays", type=int, default=0, help="debug: render only the first N rays of the frame")
    ap.add_argument("--chunk", type=int, default=37888, help="rays per kernel wave (148 SMs x 256)")
    ap.add_argument("--cpu-match-n3", type=int, default=256, help="3D points in the CPU matcher sample (0 = skip)")
    ap.add_argument("--cpu-rays", type=int, default=256, help="rays in the CPU-baseline sample (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
