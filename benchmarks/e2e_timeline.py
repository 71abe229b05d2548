#!/usr/bin/env python
"""torch.profiler timeline of the e2e_fused_loss step (rasterization() + depth_supervised_loss() + backward + .item()): kernel /
memcpy list with durations and the GPU-idle gaps between them, for steady-state steps -> gpurun_out/e2e_trace.{json,txt}.
Where the 0.3 ms between `e2e_fused_loss` and `value` of bench.py goes (profiles/r02_kernel_notes.md)."""
import sys, json, torch
sys.path.insert(0, '.')
from qed_splatter_b200 import rasterization, depth_supervised_loss
from qed_splatter_b200.scenes import scene_s1
s = scene_s1(N=1_000_000).to('cuda')
bg = torch.tensor([0.1, 0.2, 0.3], device='cuda')
params = [t.clone().requires_grad_(True) for t in (s.means, s.quats, s.scales, s.opacities, s.sh)]
gt_rgb_h = (s.gt_rgb.cpu() * 255).round().clamp(0, 255).to(torch.uint8).pin_memory()
gt_d_h = s.gt_depth.cpu().pin_memory()
rgb_d = torch.empty_like(gt_rgb_h, device='cuda'); d_d = torch.empty_like(gt_d_h, device='cuda')
cs = torch.cuda.Stream(); ev = torch.cuda.Event()
def step():
    cs.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cs):
        rgb_d.copy_(gt_rgb_h, non_blocking=True); d_d.copy_(gt_d_h, non_blocking=True); ev.record(cs)
    for p in params: p.grad = None
    render, alpha, info = rasterization(*params, s.viewmats, s.Ks, s.width, s.height, tile_size=16, packed=False, near_plane=0.01, far_plane=1e10,
                                        render_mode="RGB+ED", sh_degree=3, sparse_grad=False, absgrad=True, rasterize_mode="classic")
    info["means2d"].retain_grad()
    torch.cuda.current_stream().wait_event(ev)
    loss = depth_supervised_loss(render, alpha, rgb_d, d_d, bg, rgb_weight=0.8, depth_lambda=0.2)[0]
    loss.backward()
    return float(loss.item())
for _ in range(10): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
prof.export_chrome_trace("gpurun_out/e2e_trace.json")
ev_ = json.load(open("gpurun_out/e2e_trace.json"))["traceEvents"]
k = sorted([e for e in ev_ if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")], key=lambda e: e["ts"])
t0 = k[0]["ts"]
prev_end = None
out = []
for e in k:
    gap = (e["ts"] - prev_end) if prev_end is not None else 0.0
    out.append(f'{e["ts"]-t0:9.1f} +{e["dur"]:7.1f} gap {gap:7.1f}  {e["cat"][:10]:10s} {e["name"][:70]}')
    prev_end = max(prev_end or 0, e["ts"] + e["dur"])
open("gpurun_out/e2e_trace.txt", "w").write("\n".join(out))
print("\n".join(out[len(out)//3:2*len(out)//3]))
