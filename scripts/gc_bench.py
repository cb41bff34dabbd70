"""Timing of the device-side graph-cut iterations next to a PatchMatch iteration: wall clock around
stream-synchronised iterations, per layer.  Not part of bench.py's contract line; writes gpurun_out/gc_bench.json."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--W", type=int, default=1024); ap.add_argument("--H", type=int, default=768); ap.add_argument("--D", type=int, default=64)
    ap.add_argument("--windR", type=int, default=20); ap.add_argument("--layers", type=str, default="0,1,2")
    ap.add_argument("--budget-s", type=float, default=60.0, help="stop starting new layers once this much wall time is spent")
    a = ap.parse_args()
    import torch
    import localexpstereo_b200 as L
    from localexpstereo_b200 import synth
    from localexpstereo_b200.sweep import GCSweep, V3_PROPOSERS_DEVICE
    from oracle import lexp_oracle as O
    t0 = time.time()
    g = torch.Generator(device="cuda").manual_seed(1234)
    vol = torch.rand((a.D, a.H, a.W), generator=g, device="cuda", dtype=torch.float32)
    img = synth.synthetic_image(a.H, a.W, 42)
    torch.cuda.synchronize()
    E = L.CostVolumeEnergy(img, None, vol, None, L.Parameters(windR=a.windR, filterName="GF", filter_param1=1e-4, th_col=0.5), a.D - 1)
    S = GCSweep(E, proposers=V3_PROPOSERS_DEVICE, lam=1.0)
    rng = O.CvRNG(7)
    S.begin()
    S.init(np.stack([O.create_random_label(rng, u[0], u[1], 0.0, a.D - 1.0) for u in S.init_units]))
    E.sync()
    res = {"config": {"W": a.W, "H": a.H, "D": a.D, "windR": a.windR, "units": S.unit_sizes}, "layers": {}}
    t = time.time(); S.iteration(0, 5); E.sync(); res["pm_iteration_ms"] = (time.time() - t) * 1e3
    t = time.time(); S.iteration(1, 5); E.sync(); res["pm_iteration_ms"] = min(res["pm_iteration_ms"], (time.time() - t) * 1e3)
    for li in [int(x) for x in a.layers.split(",")]:
        if time.time() - t0 > a.budget_s:
            res["layers"][str(li)] = "skipped (time budget)"
            continue
        lay = S.lm.layers[li]
        nodes = [lay.sharedRegions[r][2] * lay.sharedRegions[r][3] for grp in lay.disjointRegionSets for r in grp]
        t = time.time(); n = S.gc_iteration(0, 9, layers=[li]); E.sync(); ms = (time.time() - t) * 1e3
        spg = n / max(1, len(lay.disjointRegionSets))   # proposal steps per group visit (the same for every group of a layer)
        res["layers"][str(li)] = {"ms": ms, "steps": n, "cells": len(nodes), "max_nodes_per_cell": int(max(nodes)), "moves": int(len(nodes) * spg),
                                  "node_moves_per_s": float(sum(nodes) * spg / (ms * 1e-3))}
        print(li, res["layers"][str(li)], flush=True)
    cost, lab = S.get()
    res["finite"] = bool(np.isfinite(cost).all())
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gc_bench.json"), "w") as f:
        json.dump(res, f)
    print(json.dumps(res))
    S.close(); E.close()


if __name__ == "__main__":
    main()
