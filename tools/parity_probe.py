"""Full-size parity probe (GPU box): renders a sample of rays of the 640x480 / V=8 / S=128 bench frame with the CUDA
path and the CPU oracle and prints the max-norm relative error of every intermediate the two expose."""
import sys, os, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from nerf_loc_b200.conditional_nerf import ConditionalNeRF
from nerf_loc_b200.config import default_args
from oracle import nerfloc_oracle as O, knn_oracle as KO

n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
torch.set_num_threads(os.cpu_count())
sc, sd, ro, rd = bench.build_frame()
scene, sup = bench.oracle_setup(sc, sd)
ro_s, rd_s, idx = bench.cpu_sample(ro, rd, n)
S = bench.S
with torch.no_grad():
    ref = O.render_rays(sd, scene, sup, sc["feat_fine_src"].permute(0, 3, 1, 2), ro_s, rd_s, sc["pose"], S,
                        knn=lambda a, b, K: KO.knn_c(a, b, K), return_debug=True)
dev = torch.device("cuda")
model = ConditionalNeRF(default_args(S)).eval()
model.load_state_dict(sd, strict=False)
model = model.to(dev)
data = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc.items() if k != "vis_featmaps"}
data["scene"], data["filename"] = "s", "f"
model.support_neural_points = None
model.multiview_aggregator.vis_featmaps = sc["vis_featmaps"].to(dev)
model.build_support_neural_points(data)
gpu_sup = model.support_neural_points["fine"]
print("support xyz bit-equal between GPU and CPU per-frame setup:", bool(torch.equal(gpu_sup["xyz"].cpu(), sup["xyz"])),
      "max abs diff", float((gpu_sup["xyz"].cpu() - sup["xyz"]).abs().max()))
conf_err = float((model.support_neural_points["fine"]["confidence"].cpu() - sup["confidence"]).abs().max() / sup["confidence"].abs().max())
if "--own-support" not in sys.argv:
    # identical per-frame inputs for both paths: inject the oracle's support points (per-frame setup is not the hot path)
    model.support_neural_points = {"fine": {k: v.to(dev) for k, v in sup.items()}, "coarse": None}
out = model.render_rays(data, {"rays_o": ro_s.to(dev), "rays_d": rd_s.to(dev), "depth_range": data["depth_range"][0]}, _debug=True)
rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())
rep = {"support_conf": conf_err}
for k in ("feature_agg", "sigma", "rgb", "depth", "weights", "depth_uncertainty", "feat"):
    rep[k] = rel(out[k].cpu(), ref[k])
# where do the weights differ?
dw = (out["weights"].cpu() - ref["weights"]).abs()
r, s = divmod(int(dw.argmax()), S)
rep["worst_weight"] = {"ray": r, "sample": s, "cuda": float(out["weights"][r, s]), "ref": float(ref["weights"][r, s]),
                       "sigma_cuda": float(out["sigma"][r, s]), "sigma_ref": float(ref["sigma"][r, s])}
ds = (out["sigma"].cpu() - ref["sigma"]).abs()
rep["sigma_abs_err_max"] = float(ds.max()); rep["sigma_max"] = float(ref["sigma"].max())
rep["sigma_rel_per_ray_max"] = float((ds.max(1)[0] / ref["sigma"].abs().max(1)[0]).max())
dfa = (out["feature_agg"].cpu() - ref["feature_agg"]).abs().view(n, S, -1).amax(-1)
rep["feature_agg_abs_err_max"] = float(dfa.max()); rep["feature_agg_max"] = float(ref["feature_agg"].abs().max())
# query-level intermediates through the public query API
z = O.sample_depths(S, *scene["depth_range"])
xyz = (ro_s[:8, None, :] + rd_s[:8, None, :] * z[None, :, None]).reshape(-1, 3)
dirs = rd_s[:8, None, :].repeat(1, S, 1).reshape(-1, 3)
with torch.no_grad():
    qo = O.query(sd, scene, xyz, sc["feat_fine_src"].permute(0, 3, 1, 2), sup, direction=torch.cat([dirs, torch.zeros(len(dirs), 1)], 1),
                 K=8, knn=lambda a, b, K: KO.knn_c(a, b, K))
qc = model.query(data, xyz.to(dev), direction=torch.cat([dirs, torch.zeros(len(dirs), 1)], 1).to(dev), K=8, _level="fine")
rep["q_knn_equal"] = bool(torch.equal(qc["knn_idx"].long().cpu(), qo["knn_idx"]))
rep["q_mvf"] = rel(qc["multiview_feature"].cpu(), qo["multiview_feature"])
rep["q_vis"] = rel(qc["multiview_visibility"].cpu(), qo["multiview_visibility"])
rep["q_feature_agg"] = rel(qc["feature_agg"].cpu(), qo["feature_agg"])
rep["q_weights"] = rel(qc["weights"].cpu(), qo["weights"])
print(json.dumps(rep, indent=1))
