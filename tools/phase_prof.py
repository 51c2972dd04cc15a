"""Prints the clock64 phase stamps of block 0 of the neighbour kernel (GPU box)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from nerf_loc_b200 import _lib
from nerf_loc_b200.conditional_nerf import ConditionalNeRF
from nerf_loc_b200.config import default_args
sc, sd, ro, rd = bench.build_frame()
dev = torch.device("cuda")
model = ConditionalNeRF(default_args(128)).eval(); model.load_state_dict(sd, strict=False); model = model.to(dev)
data = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc.items() if k != "vis_featmaps"}
data["scene"], data["filename"] = "s", "f"
model.support_neural_points = None
model.multiview_aggregator.vis_featmaps = sc["vis_featmaps"].to(dev)
rays = {"rays_o": ro[150000:150000 + 4736].to(dev), "rays_d": rd[150000:150000 + 4736].to(dev), "depth_range": data["depth_range"][0]}
for _ in range(2):
    model.render_rays(data, rays)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 64)()
_lib.load().nlb_debug_read_prof(buf, 64)
v = list(buf)
if True:
    names = ["P0 (both sub-tiles) + q rows", "wait L1 (sub-tile 0)", "E1 (both)", "E2 (both, incl. waits)", "E3 (both, incl. waits)",
             "wait key projection", "scores + softmax (both)", "wait value projection", "context + weights (both)"]
    for i in range(9):
        print(f"nb2 {names[i]:36s} {v[i+1]-v[i]:8d} clk")
    print("nb2 super-tile (32 samples) total", v[9] - v[0])

# aggregate_kernel, render variant (stamps 16 + i: 0 start, 1 projections done, 4 view weights done, 7 gather + statistics done)
print(f"agg projections (256 rows)            {v[17]-v[16]:8d} clk")
print(f"agg visibility rows + view weights    {v[20]-v[17]:8d} clk")
print(f"agg gather + statistics (4 samples per warp) {v[23]-v[20]:8d} clk")
print("agg tile (32 samples) total", v[23] - v[16])

rn = ["load x (both rays)", "blend weights + wait blend GEMM", "blend MLP + softmax (both rays)", "wait conv1", "epi conv1", "wait conv2",
      "epi conv2", "wait conv3", "epi conv3", "wait tconv3", "epi tconv3", "wait tconv2", "epi tconv2", "wait tconv1",
      "epi tconv1 + reload x", "wait conv_out", "epi conv_out", "composite", "feat"]
rv = v[32:]
for i in range(len(rn)):
    print(f"ray {rn[i]:34s} {rv[i+1]-rv[i]:8d} clk")
print("ray total (per pair of rays)", rv[len(rn)] - rv[0])
print("ray blend build detail: math+st", rv[25]-rv[20], "fetch issue", rv[26]-rv[25], "st wait", rv[27]-rv[26], "sync", rv[21]-rv[27])
print("ray blend detail (ray 0): sBl", rv[20]-rv[2], "batch0 build A", rv[21]-rv[20], "batch0 MMA wait", rv[22]-rv[21], "rest of batches", rv[23]-rv[22], "softmax", rv[24]-rv[23])
