#!/usr/bin/env python
"""Benchmark of the NeRF-Loc render-and-match hot path on B200 (BASELINE.json: rays/sec, 640x480, 128 samples/ray, 8 views).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm's CPU port (oracle/) on host cores
    python bench.py --config 3 | --config sweep              # the other BASELINE.json configurations (not the headline)

A step = one pass of `ConditionalNeRF.render_rays` over every pixel ray of the synthetic query frame.  --config 1 (default) is
configs[1], the frame the metric is quoted on (640x480, 128 samples per ray, 8 reference views); its line also carries the
configs[2] pipeline figure (render + coarse-to-fine matching + PnP, ms per frame).  --config 3 is the Cambridge-shape frame
(1920x1056, 192 samples per ray), --config sweep the ray-count sweep 2^14 .. 2^22 at 64 / 128 / 256 samples per ray (one JSON
line per point).  With N GPUs the rays of the SAME frame are split into N contiguous slices (strong scaling) and the rendered
per-ray features are gathered on every rank once per step, as the matcher needs them - by the ray kernel's own epilogue (peer
stores over NVLink into symmetric memory, nerf_loc_b200/distributed.py::FeatExchange) plus one barrier.  Rank 0 prints one JSON
line per measured point.
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
# stdout carries the JSON line(s) and nothing else: native libraries that write to file descriptor 1 (the "NCCL version ..."
# banner, for one) are sent to stderr, the result lines go through a private copy of the original descriptor
_OUT = None


def claim_stdout():
    """Called once by main(): from here on file descriptor 1 is stderr for everybody but emit()."""
    global _OUT
    if _OUT is None:
        sys.stdout.flush()
        _OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _OUT if _OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


CONFIGS = {
    # name: (H, W, V, S, workload)
    "1": (480, 640, 8, 128, "640x480 query, 128 samples/ray, 8 ref views, full conditional render (configs[1])"),
    "3": (1056, 1920, 8, 192, "Cambridge-shape 1920x1056 query, 192 samples/ray, 8 ref views, rays sharded over the GPUs (configs[3])"),
}


def f_sample(V):
    """algorithmic matmul + conv FLOP per sample point (SURVEY.md 8d / BASELINE.md section 3)"""
    return 2880640 + 38912 * V


def f_kernel(V):
    """F_SAMPLE attributed to the kernels that execute it now, FLOP per sample point"""
    return {"knn_query_rays": 0.0,
            "visibility": 176e3 * V / 8,                        # visibility decoder, per view
            "aggregate": 201e3 - 66.7e3,                        # per-view blend layer, mean / variance
            "fc_tail": 66.7e3 + 32.8e3,                         # out_fc 393 -> 64 -> 128, query projection
            "qproj": 32.8e3,
            "neighbor2": 1108e3 + 1081e3 + 264e3 - 2 * 32.8e3,  # base_mlp, key / value projections, attention, weights
            "neighbor": 1108e3 + 1081e3 + 264e3,
            "attn_tail": 32.8e3,                                # output projection of the attention + LayerNorm
            "ray": 270e3 + 82e3 + 9e3, "ray_long": 270e3 + 82e3 + 9e3}


# dram__bytes_read.sum + dram__bytes_write.sum per RAY and kernel from the committed `ncu --set full` capture of this build
# (profiles/traffic_per_ray.json, written by tools/ncu_traffic.py); scaled to the rays one launch of the timed run covers
TRAFFIC_PER_RAY = {}
try:
    TRAFFIC_PER_RAY = json.load(open(os.path.join(ROOT, "profiles", "traffic_per_ray.json")))["bytes_per_ray"]
except Exception:
    pass


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


def build_frame(H=480, W=640, V=8, S=128):
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


def oracle_render(sc, sd, scene, sup, ro, rd, S, chunk=2048):
    """The CPU port through the reference's own chunk loop (model.py:616-633: 2048 rays per chunk)."""
    from oracle import knn_oracle as KO
    from oracle import nerfloc_oracle as O
    knn = lambda a, b, K: KO.knn_c(a, b, K)  # exact, all host threads (KNN variant (ii) of BASELINE.md section 4)
    outs = []
    with torch.no_grad():
        for s in range(0, ro.shape[0], chunk):
            outs.append(O.render_rays(sd, scene, sup, sc["feat_fine_src"].permute(0, 3, 1, 2), ro[s:s + chunk], rd[s:s + chunk],
                                      sc["pose"], S, knn=knn))
    return {k: torch.cat([o[k] for o in outs]) for k in outs[0]}


def cpu_sample(ro, rd, n):
    idx = torch.linspace(0, ro.shape[0] - 1, n).long()
    return ro[idx].contiguous(), rd[idx].contiguous(), idx


def faithful_knn_probe(sc, sup, ro, rd, S, n_rays=4):
    """KNN variant (i) of BASELINE.md section 4: the reference's OWN knn_cpu.cpp (single-threaded; oracle/_ref/knn_ref.so, built
    from /root/reference in the build container) on a few rays, beside the threaded C port the CPU baseline uses."""
    from oracle import knn_oracle as KO
    from oracle import nerfloc_oracle as O
    if not KO.ref_available():
        return None
    ro_s, rd_s, _ = cpu_sample(ro, rd, n_rays)
    z = O.sample_depths(S, *sc["depth_range"][0])
    q = (ro_s[:, None, :] + rd_s[:, None, :] * z[None, :, None]).reshape(-1, 3)
    t0 = time.perf_counter()
    KO.knn_reference(q, sup["xyz"], 8)
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    KO.knn_c(q, sup["xyz"], 8)
    dt2 = time.perf_counter() - t1
    return {"queries": int(q.shape[0]), "support_points": int(sup["xyz"].shape[0]), "reference_knn_cpu_s": dt,
            "threaded_port_s": dt2, "knn_only_rays_per_s_reference": n_rays / dt, "cores": 1,
            "note": "knn_cpu.cpp compiled from the reference (single-threaded), KNN alone; the CPU baseline uses the threaded port"}

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
    """Match ms/frame (BASELINE.json metric, second half): Matcher.forward at configs[2] size on the device, the S2D kernel alone
    with its tensor roofline, the CPU port on a bounded sample (fewer 3D points, same 2D maps), and the PnP stage."""
    from nerf_loc_b200 import params, pnp, synthetic as syn
    from nerf_loc_b200.config import default_args
    from nerf_loc_b200.matcher import Matcher
    sd = syn.synthetic_state_dict(params.matcher_shapes(), 99)
    m = Matcher(default_args(), 192, 192, 192).eval()
    m.load_state_dict(sd)
    m = m.to(dev)
    data = matcher_frame()
    dd = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in data.items()}
    tf_peak, _, which = peaks()
    with torch.no_grad():
        for _ in range(3):
            out = m(dict(dd))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = m(dict(dd))
        e1.record()
        torch.cuda.synchronize()
        # the S2D kernel alone
        d3t, d2t = m.coarse_transformer(dd["desc_3d"][None], dd["pos_emd_3d"][None], dd["desc_2d_coarse"][None], dd["pos_emd_2d"][None])
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(5):
            m.s2d(d3t[0], d2t[0], 0.2)
        f1.record()
        torch.cuda.synchronize()
    s2d_ms = f0.elapsed_time(f1) / 5
    s2d_tf = 82368.0 * 4096 * 4800 / (s2d_ms * 1e-3) / 1e12
    res = {"ms_per_frame": e0.elapsed_time(e1) / 5, "s2d_ms": s2d_ms, "n3": 4096, "mc": 4800,
           "matches": int(out["i_ids"].numel()),
           "roofline": {"bound": "tensor", "kernel": "s2d_tc_kernel (+ split_cells / colmax / rowmatch / compact)", "achieved": s2d_tf,
                        "peak": tf_peak, "unit": "TFLOP/s", "frac": s2d_tf / tf_peak, "peak_source": which + " bf16 sustained",
                        "mma_mode": "bf16x3 tcgen05 (3 MMAs per product: ceiling = peak / 3)", "frac_of_mode_ceiling": 3 * s2d_tf / tf_peak}}
    # PnP (configs[2]): device P3P-RANSAC + LM on synthetic correspondences.  Parity UNPINNED: COLMAP is absent, validated against
    # the known synthetic pose only
    from oracle import pnp_oracle as P
    p2d, p3d, cam, Rgt, tgt, _ = P.synthetic_correspondences(M=max(64, int(out["i_ids"].numel())), seed=0)
    p2d_d, p3d_d = torch.from_numpy(p2d).to(dev), torch.from_numpy(p3d).to(dev)
    pnp.absolute_pose_estimation(p2d_d, p3d_d, cam, 8.0, iters=2048, seed=0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        ret = pnp.absolute_pose_estimation(p2d_d, p3d_d, cam, 8.0, iters=2048, seed=0)
    torch.cuda.synchronize()
    res["pnp"] = {"ms": (time.perf_counter() - t0) / 3 * 1e3, "matches": int(p2d.shape[0]), "iters": 2048, "success": bool(ret["success"]),
                  "parity": "unpinned (pycolmap absent): validated against the known synthetic pose"}
    if ret["success"]:
        rot, pos = P.pose_error(ret["R"], ret["tvec"], Rgt, tgt)
        res["pnp"]["pose_error_deg_m"] = [float(rot), float(pos)]
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
        res["parity_on_sample"] = {"score_matrix": float((got["score_matrix"].cpu() - ref["score_matrix"]).abs().max() / ref["score_matrix"].abs().max()),
                                   "ids_equal": bool(torch.equal(got["i_ids"].cpu(), ref["i_ids"]) and torch.equal(got["j_ids"].cpu(), ref["j_ids"]))}
    return res


def reference_rays_per_step(steps, warmup):
    """Bounded sample of the frame per step: about four minutes of host time over all steps at ~280 rays/s, between one and
    eight of the reference's own 2048-ray chunks."""
    n = 64000 // max(1, steps)
    return max(2048, min(16384, n // 2048 * 2048))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    H, W, V, S, workload = CONFIGS["1" if args.config == "sweep" else args.config]
    sc, sd, ro, rd = build_frame(H, W, V, S)
    scene, sup = oracle_setup(sc, sd)
    n = args.cpu_rays if args.cpu_rays > 0 else reference_rays_per_step(args.steps, args.warmup)
    ro_s, rd_s, _ = cpu_sample(ro, rd, n)
    for _ in range(args.warmup):
        oracle_render(sc, sd, scene, sup, ro_s[:64], rd_s[:64], S)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_render(sc, sd, scene, sup, ro_s, rd_s, S)
    dt = (time.perf_counter() - t0) / args.steps
    val = n / dt
    line = {
        "impl": "reference", "metric": "rays/sec", "value": val, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "rays_per_step": n, "samples_per_ray": S, "views": V},
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{n} evenly spaced rays of the frame x {S} samples in the reference's own chunks of 2048 rays "
                                   f"(oracle/ render_rays, exact KNN on all host threads); warm-up steps use 64 rays",
                         "faithful_knn": faithful_knn_probe(sc, sup, ro, rd, S)},
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


class Runner:
    """The CUDA path on this rank's slice of a frame: model, scene, symmetric-memory exchange, timed steps."""

    def __init__(self, args, H, W, V, S):
        import torch.distributed as dist
        from nerf_loc_b200 import _lib
        from nerf_loc_b200.conditional_nerf import ConditionalNeRF
        from nerf_loc_b200.config import default_args
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dev = torch.device("cuda", self.local)
        self.L = _lib.load()
        self.H, self.W, self.V, self.S, self.args = H, W, V, S, args
        self.sc, self.sd, self.ro_all, self.rd_all = build_frame(H, W, V, S)
        dev = self.dev
        model = ConditionalNeRF(default_args(S)).eval()
        model.load_state_dict(self.sd, strict=False)
        self.model = model.to(dev)
        self.model.chunk_rays = args.chunk
        self.data = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in self.sc.items() if k != "vis_featmaps"}
        self.data["scene"], self.data["filename"] = "synthetic", "bench"
        self.setup_breakdown = self.frame_setup()
        self.exch, self.exch_note = None, ""

    def frame_setup(self):
        """Per-frame setup with its parts timed: DepthFusionNet visibility maps (cuDNN), back-projection + support
        features + confidence + KNN index, per-frame pre-projections.  The render itself uses the synthetic visibility maps of
        the frame (so that the parity sample sees the inputs the CPU port sees)."""
        m, d, dev = self.model, self.data, self.dev
        out = {}
        torch.cuda.synchronize()
        for rep in range(2):   # first pass: cuDNN autotune / lazy module init
            m.support_neural_points = None
            m.multiview_aggregator.vis_featmaps = None
            t0 = time.perf_counter()
            try:
                m._vis_maps(d)
                torch.cuda.synchronize()
                out["depth_fusion_net_ms"] = (time.perf_counter() - t0) * 1e3
            except Exception as e:   # (the 2D net is outside the parity scope; never fatal for the bench)
                out["depth_fusion_net_ms"] = None
                out["depth_fusion_net_error"] = type(e).__name__
            m.multiview_aggregator.vis_featmaps = self.sc["vis_featmaps"].to(dev)
            t1 = time.perf_counter()
            m.build_support_neural_points(d)
            torch.cuda.synchronize()
            out["support_points_ms"] = (time.perf_counter() - t1) * 1e3
            t2 = time.perf_counter()
            m._level_scene(d, "fine", query_pose=d["pose"])
            torch.cuda.synchronize()
            out["scene_precompute_ms"] = (time.perf_counter() - t2) * 1e3
        out["total_ms"] = sum(v for k, v in out.items() if k.endswith("_ms") and v is not None)
        return out

    def shard(self, R_total):
        per = (R_total + self.world - 1) // self.world
        self.R_total = R_total
        self.lo, self.hi = min(R_total, self.rank * per), min(R_total, (self.rank + 1) * per)
        reps = (R_total + self.ro_all.shape[0] - 1) // self.ro_all.shape[0]
        ro = self.ro_all.repeat(reps, 1)[:R_total] if reps > 1 else self.ro_all[:R_total]
        rd = self.rd_all.repeat(reps, 1)[:R_total] if reps > 1 else self.rd_all[:R_total]
        self.ro_h, self.rd_h = ro[self.lo:self.hi].contiguous().pin_memory(), rd[self.lo:self.hi].contiguous().pin_memory()
        self.ro_d, self.rd_d = self.ro_h.to(self.dev), self.rd_h.to(self.dev)
        self.host_out = {}
        if self.world > 1:
            from nerf_loc_b200.distributed import FeatExchange
            dist = self.dist
            # agree on the path BEFORE the collective allocation: a rank that failed alone inside the rendezvous would hang the rest
            try:
                __import__("torch.distributed._symmetric_memory")
                ok = torch.ones(1, device=self.dev)
            except Exception as e:  # symmetric memory not available in this build: one NCCL all-gather per step instead
                ok, self.exch_note = torch.zeros(1, device=self.dev), f"{type(e).__name__}"
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)   # all ranks take the same path
            self.exch = FeatExchange(R_total, self.dev) if float(ok.item()) != 0.0 else None

    def step_device(self):
        from nerf_loc_b200.distributed import all_gather_rows
        rays = {"rays_o": self.ro_d, "rays_d": self.rd_d, "depth_range": self.data["depth_range"][0]}
        if self.exch is None:
            out = self.model.render_rays(self.data, rays)
            if self.world > 1:
                out["feat_all"] = all_gather_rows(out["feat"], self.R_total)
            return out
        out = self.model.render_rays(self.data, rays, _feat_peers=(self.exch.begin_frame(), self.lo))
        self.exch.barrier()
        out["feat_all"] = self.exch.gathered()
        return out

    def step_e2e(self):
        from nerf_loc_b200.distributed import all_gather_rows
        dev = self.dev
        rays = {"rays_o": self.ro_h.to(dev, non_blocking=True), "rays_d": self.rd_h.to(dev, non_blocking=True),
                "depth_range": self.data["depth_range"][0]}
        if self.exch is None:
            out = self.model.render_rays(self.data, rays)
            if self.world > 1:
                all_gather_rows(out["feat"], self.R_total)
        else:
            out = self.model.render_rays(self.data, rays, _feat_peers=(self.exch.begin_frame(), self.lo))
            self.exch.barrier()
        for k, v in out.items():
            if k not in self.host_out:
                self.host_out[k] = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
            self.host_out[k].copy_(v, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())

    def kernel_profile(self):
        """Per-kernel device time of one extra pass (CUDA events after every launch on the launch stream; same launch order)."""
        L = self.L
        L.nlb_profile_enable(1)
        self.step_device()
        torch.cuda.synchronize()
        buf = ctypes.create_string_buffer(4096)
        L.nlb_profile_report(buf, 4096)
        L.nlb_profile_enable(0)
        kern = {}
        for rec in buf.value.decode().split(";"):
            if rec:
                name, ms, n = rec.split(":")
                kern[name] = {"ms": float(ms), "launches": int(n)}
        return kern

    def measure(self, workload, extra_config=None):
        args, world, S, V = self.args, self.world, self.S, self.V
        Rl = self.hi - self.lo
        for _ in range(args.warmup):
            self.step_device()
        sampler = ClockSampler(self.local)
        sampler.start()
        total_ms = self.timed(self.step_device, args.steps)
        sampler.stop_flag = True
        self.step_e2e()
        e2e_steps = max(1, min(args.steps, 3))
        e2e_ms = self.timed(self.step_e2e, e2e_steps)
        d2h = sum(v.numel() * v.element_size() for v in self.host_out.values())
        kern = self.kernel_profile()
        if self.rank != 0:
            return None
        tf_peak, hbm_peak, which = peaks()
        fk = f_kernel(V)
        samples_rank = Rl * S
        names = list(kern)
        dom = max(names, key=lambda n: kern[n]["ms"])
        per_kernel = {}
        for n in names:
            ach = fk.get(n, 0.0) * samples_rank / (kern[n]["ms"] * 1e-3) / 1e12 if kern[n]["ms"] > 0 else 0.0
            tr = TRAFFIC_PER_RAY.get(n)
            per_kernel[n] = {"ms": kern[n]["ms"], "launches": kern[n]["launches"], "achieved_tflops": ach, "frac": ach / tf_peak,
                             "dram_bytes_per_launch": (tr * Rl / max(1, kern[n]["launches"])) if tr is not None else None}
        ms_step = total_ms / args.steps
        n_sup = int(self.model.support_neural_points["fine"]["xyz"].shape[0])
        compulsory = (V * (3 * self.H * self.W + 224 * (self.H // 4) * (self.W // 4)) * 4 + n_sup * 206 * 4 + 32 * self.R_total
                      + (198 + S) * 4 * self.R_total)
        tr_dom = TRAFFIC_PER_RAY.get(dom)
        whole = f_sample(V) * self.R_total * S / (ms_step * 1e-3) / 1e12
        # products the tensor cores really issue per sample point (each once, not x3), for the kernels that are MMA chains
        exe_flop = {"visibility": V * (32 * 128 + 4 * 32 * 32) * 2, "fc_tail": (416 * 64 + 64 * 128 + 128 * 128) * 2,
                    "neighbor2": 8 * (96 + 4 * 128) * 128 * 2, "attn_tail": 128 * 128 * 2}
        executed = None
        if dom in exe_flop and kern[dom]["ms"] > 0:
            tf = exe_flop[dom] * samples_rank / (kern[dom]["ms"] * 1e-3) / 1e12
            executed = {"kernel": dom, "tflops": tf, "frac_of_mode_ceiling": tf / (tf_peak / 3.0)}
        line = {
            "metric": "rays/sec", "value": self.R_total / (ms_step * 1e-3), "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "rays_per_step": self.R_total, "samples_per_ray": S, "views": V,
                       "support_points": n_sup, "chunk_rays": args.chunk,
                       "mma_mode": "bf16x3 on tcgen05 (operands split into bf16 hi + lo, three MMAs per product, fp32 accumulation in "
                                   "TMEM): visibility decoder, out_fc + query projection, base_mlp + key / value projections, attention "
                                   "output projection, RayUnet / blend / feat layers; fp32 FFMA2 for the gather-side arithmetic"
                                   + ("" if S <= 128 else "; S > 128: RayUnet on the fp32 slab kernel"),
                       "l2": "working set per step (scene 294 MB + >1 GB of per-chunk intermediates) exceeds the 126 MB L2",
                       "parallelism": f"ray-shard x{world}" + ("" if world == 1 else (
                           " + all-gather of feat[R,192] fused into the ray kernel epilogue (peer stores over NVLink, symmetric memory, two "
                           "copies alternating by frame)" if self.exch is not None else
                           " + NCCL all-gather of feat[R,192] (symmetric memory unavailable: " + self.exch_note + ")")),
                       "per_frame_setup_ms": self.setup_breakdown},
            "e2e": {"value": self.R_total / (e2e_ms / e2e_steps * 1e-3), "unit": "rays/s",
                    "h2d_bytes_per_step": int(self.ro_h.numel() * 4 * 2), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(self.L.nlb_render_launch_count(Rl, args.chunk)) * args.steps,
            "clocks": sampler.summary(),
            "kernels_ms_per_step": {n: kern[n]["ms"] for n in names},
            "roofline": {"bound": "tensor", "kernel": dom, "achieved": per_kernel[dom]["achieved_tflops"], "peak": tf_peak,
                         "unit": "TFLOP/s", "frac": per_kernel[dom]["frac"],
                         "traffic": (tr_dom * Rl / max(1, kern[dom]["launches"])) if tr_dom is not None else None,
                         "peak_source": which + " bf16 sustained",
                         "note": "achieved = algorithmic FLOP of the stage (the reference's arithmetic, BASELINE.md section 3) / CUDA-event "
                                 "time of its kernel (one profiled pass, same launch order).  The bf16x3 mode spends three MMAs per product "
                                 "(ceiling = peak / 3), and the exact rewrites of DESIGN.md section 3 remove part of a stage's products, so the "
                                 "algorithmic figure of a stage can exceed that ceiling: `executed` counts the products the MMAs really issue",
                         "executed": executed,
                         "whole_step_achieved": whole, "whole_step_frac": whole / tf_peak,
                         "hbm": {"compulsory_bytes_per_frame": int(compulsory),
                                 "measured_dram_bytes_per_frame": (sum(TRAFFIC_PER_RAY.get(n, 0.0) for n in names) * self.R_total) if TRAFFIC_PER_RAY else None,
                                 "peak_gbs": hbm_peak},
                         "per_kernel": per_kernel},
        }
        if extra_config:
            line["config"].update(extra_config)
        return line


def knn_comparator(r):
    """This repo's ray-sample search beside the REFERENCE's own CUDA KNN (ops/knn/src/knn.cu, compiled for sm_100a into
    oracle/_ref/knn_ref_cuda.so in the build container) on the same box, same queries: comparator only."""
    path = os.path.join(ROOT, "oracle", "_ref", "knn_ref_cuda.so")
    if not os.path.exists(path):
        return {"unavailable": "oracle/_ref/knn_ref_cuda.so not built (make -C oracle ref_cuda)"}
    try:
        from nerf_loc_b200 import _lib
        from oracle import nerfloc_oracle as O
        lib = ctypes.CDLL(path)
        lib.knn_ref_cuda.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        dev, S = r.dev, r.S
        R = 8192
        ro, rd = r.ro_d[:R].contiguous(), r.rd_d[:R].contiguous()
        z = O.sample_depths(S, *r.sc["depth_range"][0]).to(dev).contiguous()
        q = (ro[:, None, :] + rd[:, None, :] * z[None, :, None]).reshape(-1, 3).contiguous()
        sup = r.model.support_neural_points["fine"]["xyz"].contiguous()
        idx = torch.empty(q.shape[0], 8, dtype=torch.int64, device=dev)
        d2 = torch.empty(q.shape[0], 8, device=dev)
        out = {}
        for ver in (2,):   # the version the reference picks for D = 3, K = 8 (knn.cu: ChooseVersion)
            lib.knn_ref_cuda(q.data_ptr(), q.shape[0], sup.data_ptr(), sup.shape[0], 3, 8, ver, idx.data_ptr(), d2.data_ptr())
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            lib.knn_ref_cuda(q.data_ptr(), q.shape[0], sup.data_ptr(), sup.shape[0], 3, 8, ver, idx.data_ptr(), d2.data_ptr())
            torch.cuda.synchronize()
            out[f"reference_v{ver}_ms"] = (time.perf_counter() - t0) * 1e3
        L = _lib.load()
        index = r.model._frame["sup_fine"].index
        geo = torch.zeros(sup.shape[0], 8, device=dev)
        geo[:, :3] = sup
        i32 = torch.empty(q.shape[0], 8, dtype=torch.int32, device=dev)
        d2b = torch.empty(q.shape[0], 8, device=dev)
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _lib.check(L.nlb_debug_knn_rays(_lib.ptr(index.buf), _lib.ptr(ro), _lib.ptr(rd), _lib.ptr(z), 0, _lib.ptr(geo), R, S,
                                            _lib.ptr(i32), _lib.ptr(d2b), _lib.stream()))
            torch.cuda.synchronize()
            out["ours_ms"] = (time.perf_counter() - t0) * 1e3
        # the reference's CUDA kernel accumulates the squared distance with contracted FMAs, its CPU code (the definition this
        # repo reproduces bit for bit, tests/test_gpu_render.py) does not: neighbours can swap at 1-ulp near-ties
        out.update({"queries": int(q.shape[0]), "support_points": int(sup.shape[0]),
                    "index_agreement": float((i32.long() == idx).float().mean()),
                    "max_rel_distance_diff": float(((d2b - d2).abs() / d2.clamp_min(1e-12)).max())})
        return out
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def run_b200(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (nerf_loc_b200 has no CPU path); use --impl reference for the CPU port")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    if args.config == "sweep":
        # ray-count sweep (configs[4]): 2^14 .. 2^22 rays of the 640x480 frame (tiled beyond one frame) at 64 / 128 / 256 samples
        H, W, V, _, _ = CONFIGS["1"]
        for S in (64, 128, 256):
            r = Runner(args, H, W, V, S)
            for e in range(14, 23):
                if e > args.sweep_max:
                    break
                r.shard(2 ** e)
                line = r.measure(f"ray-count sweep (configs[4]): 2^{e} rays x {S} samples/ray, 8 ref views 640x480", {"sweep_point": [2 ** e, S]})
                if rank == 0:
                    line["cpu_baseline"] = None
                    emit(line)
            del r
            torch.cuda.empty_cache()
    else:
        H, W, V, S, workload = CONFIGS[args.config]
        r = Runner(args, H, W, V, S)
        R_total = r.ro_all.shape[0] if args.rays <= 0 else args.rays
        r.shard(R_total)
        line = r.measure(workload)
        if rank == 0:
            if args.cpu_rays > 0 and world == 1:
                torch.set_num_threads(os.cpu_count() or 1)
                scene, sup = oracle_setup(r.sc, r.sd)
                ro_s, rd_s, idx = cpu_sample(r.ro_all[:R_total], r.rd_all[:R_total], args.cpu_rays)
                oracle_render(r.sc, r.sd, scene, sup, ro_s[:16], rd_s[:16], S)
                t0 = time.perf_counter()
                ref = oracle_render(r.sc, r.sd, scene, sup, ro_s, rd_s, S)
                dt = time.perf_counter() - t0
                line["cpu_baseline"] = {"value": args.cpu_rays / dt, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                                        "sample": f"{args.cpu_rays} evenly spaced rays of the frame x {S} samples in chunks of 2048 rays "
                                                  f"(oracle/ render_rays, exact KNN on all host threads)",
                                        "faithful_knn": faithful_knn_probe(r.sc, sup, r.ro_all, r.rd_all, S)}
                # parity of the benchmarked frame on that sample (checker only).  The device frame setup is the model's own:
                # its back-projection reproduces the host operators bit for bit (nlb_backproject_points), so the KNN sees
                # the same support points as the CPU path.  (A cuBLAS back-projection differs by an ulp in places and flips
                # near-tied neighbours - the reference itself is discontinuous there; if the clouds ever differ, the CPU
                # ones are injected and the line says so.)
                own = r.model.support_neural_points["fine"]
                same = all(torch.equal(own[k].cpu(), sup[k]) for k in ("xyz", "feature", "direction"))
                line["support_points_bit_identical"] = bool(same)
                if not same:
                    r.model.support_neural_points = {"fine": {k: v.to(dev) for k, v in sup.items()}, "coarse": None}
                out = r.step_device()
                line["parity_on_sample"] = {k: float((out[k][idx.to(dev)].cpu() - ref[k]).abs().max() / ref[k].abs().max())
                                            for k in ("rgb", "depth", "feat", "weights")}
                if args.config == "1":
                    line["knn_vs_reference_gpu"] = knn_comparator(r)
                    line["match"] = bench_matcher(dev, args.cpu_match_n3)
                    # configs[2]: render + coarse-to-fine matching + PnP as one pipeline figure on this GPU
                    line["pipeline"] = {"workload": "configs[2]: full render + matching (4096 3D pts x 60x80 / 120x160 maps) + PnP-RANSAC",
                                        "render_ms": line["ms_per_step"], "match_ms": line["match"]["ms_per_frame"],
                                        "pnp_ms": line["match"]["pnp"]["ms"],
                                        "ms_per_frame": line["ms_per_step"] + line["match"]["ms_per_frame"] + line["match"]["pnp"]["ms"],
                                        "note": "sum of the three stages measured back to back on one GPU; PnP parity unpinned (COLMAP absent)"}
            else:
                line["cpu_baseline"] = None
            emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="1", choices=["1", "3", "sweep"],
                    help="1: the metric's frame (default; its line also carries configs[2]); 3: Cambridge shape; sweep: configs[4]")
    ap.add_argument("--sweep-max", type=int, default=22, help="largest exponent of the ray-count sweep")
    ap.add_argument("--rays", type=int, default=0, help="debug: render only the first N rays of the frame")
    ap.add_argument("--chunk", type=int, default=76960, help="cap on the rays per kernel wave (148 SMs x 520: a 640x480 frame is four equal waves)")
    ap.add_argument("--cpu-match-n3", type=int, default=256, help="3D points in the CPU matcher sample (0 = skip)")
    ap.add_argument("--cpu-rays", type=int, default=-1,
                    help="rays in the CPU sample: b200 arm default 4096 (about 15 s of host time), 0 = skip; reference arm default: "
                         "a multiple of 2048 sized from --steps")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        args.cpu_rays = max(args.cpu_rays, 0)
        run_reference(args)
    else:
        if args.cpu_rays < 0:
            args.cpu_rays = 4096
        run_b200(args)


if __name__ == "__main__":
    main()
