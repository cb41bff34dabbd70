"""Latency of ONE launch of lexp_fused_kernel as a function of how many work items share the GPU (0.25 .. 4 per SM): separates the
intrinsic latency of a work item (its serial chain) from throughput effects.  100x100 filterRects (the layer-0 cell of the
2048x1536 configuration), unary mode, CUDA-event timing of 20 launches each.  B200: python scripts/latency_probe.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import localexpstereo_b200 as L
from localexpstereo_b200 import synth

W, H, D, windR = 2048, 1536, 64, 20
dev = torch.device("cuda", 0)
vol = torch.rand((D, H, W), device=dev)
img = synth.synthetic_image(H, W, 42)
E = L.CostVolumeEnergy(img, None, vol, None, L.Parameters(windR=windR, filterName="GF", filter_param1=1e-4, th_col=0.5), D - 1)
stream = torch.cuda.current_stream(dev)
E.set_stream(stream.cuda_stream)
lay = L.LayerManager(W, H, windR).addLayer(20)
interior = [r for r in range(len(lay.unitRegions)) if lay.filterRegions[r][2] == 100 and lay.filterRegions[r][3] == 100]
cost = torch.zeros((H, W), device=dev)
rng = np.random.default_rng(0)
for n in (37, 74, 148, 222, 296, 444, 592, 1184):
    cells = [interior[i] for i in rng.choice(len(interior), size=n, replace=False)]
    plan = E.make_plan([lay.filterRegions[r] for r in cells], [lay.sharedRegions[r] for r in cells])
    planes = synth.synthetic_planes([lay.unitRegions[r] for r in cells], 1, D, 3)[0]
    d_pl = torch.from_numpy(planes).to(dev)
    os.environ["X"] = "1"
    for _ in range(3):
        plan.eval_device(d_pl.data_ptr(), cost.data_ptr(), W * 4, True, 0, planes_on_device=True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        plan.eval_device(d_pl.data_ptr(), cost.data_ptr(), W * 4, True, 0, planes_on_device=True)
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"items {plan.num_items:5d} ({plan.num_items / 148:.2f} per SM): {np.median(ts):7.1f} us per launch, {np.median(ts) / 50:.2f} us per 2-row chunk if one wave")
    plan.close()
E.close()
