#!/usr/bin/env python
"""Mints tests/golden/*.npz from the numpy oracle (oracle/lexp_oracle.py).

Run in the authoring container (needs /root/reference/data for the natural guide image):
    python tests/golden/make_golden.py
The reference ships no golden vectors for this path, so these are minted from
the restatement; inputs are embedded so that the GPU box (no /root/reference) can replay them."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import lexp_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def cones_crop():
    import cv2
    im = cv2.imread("/root/reference/data/MiddV2/cones/imL.png")
    imR = cv2.imread("/root/reference/data/MiddV2/cones/imR.png")
    return np.ascontiguousarray(im[100:228, 150:310]), np.ascontiguousarray(imR[100:228, 150:310])  # 128 x 160


def main():
    imL, imR = cones_crop()
    H, W, D, windR, eps, th = imL.shape[0], imL.shape[1], 20, 20, 1e-4, 0.5
    volL = O.synthetic_volume(D, H, W, 1234)
    volL = O.fill_out_of_view(volL, 0)              # main.cpp:360
    volR = O.fill_out_of_view(O.convert_volume_l2r(volL), 1)  # main.cpp:363-368
    E = O.CostVolumeEnergyOracle(imL, imR, volL, volR, windR, eps, th, D - 1)
    layer = O.make_layer(W, H, windR, 16)
    rng = O.CvRNG(2024)
    cases = []
    cells = [0, 1, 9, 10, 37, 44, 70, 79]  # corners, edges, interior, merged edge cells
    for mode in (0, 1):
        for r in cells:
            u = layer["unit"][r]
            p = O.create_random_label(rng, u[0] + u[2] // 2, u[1] + u[3] // 2, 0.0, D - 1.0)
            cases.append((mode, layer["filter"][r], layer["shared"][r], p, True))
    # branch coverage: below MIN, above MAX, NaN, steep, no-check, 1x1 target, filterRect == targetRect
    f, t = (20, 10, 110, 100), (40, 30, 70, 60)
    for p, chk in [((0, 0, -3.0, 0), True), ((0, 0, D + 2.0, 0), False), ((0.3, -0.2, 4.0, 0), True),
                   ((float("nan"), 0, 1.0, 0), False), ((1.7, 1.7, -150.0, 0), True), ((0.05, 0.02, 7.3, 0), False)]:
        cases.append((0, f, t, np.array(p, np.float32), chk))
    cases.append((0, (40, 40, 41, 41), (60, 60, 1, 1), np.array((0.01, 0.02, 5.0, 0), np.float32), True))
    cases.append((1, (50, 40, 60, 50), (50, 40, 60, 50), np.array((-0.1, 0.05, 9.0, 0), np.float32), True))
    # the volumes are regenerated from the seed by the tests (numpy PCG64 is stable); only checksums are stored
    out = dict(imL=imL, imR=imR, vol_seed=np.array(1234), vol_sums=np.array([volL.astype(np.float64).sum(), volR.astype(np.float64).sum()]),
               params=np.array([windR, eps, th, D - 1], np.float64), n=np.array(len(cases)))
    for i, (mode, fr, tr, p, chk) in enumerate(cases):
        ref = (E.compute_unary_potential if chk else E.compute_unary_potential_without_check)(fr, tr, p, mode)
        out[f"mode{i}"] = np.array(mode); out[f"frect{i}"] = np.array(fr, np.int32); out[f"trect{i}"] = np.array(tr, np.int32)
        out[f"plane{i}"] = np.asarray(p, np.float32); out[f"check{i}"] = np.array(int(chk)); out[f"ref{i}"] = ref
    out["stats0"] = E.filter[0].stats_f32()[:, ::8, ::8].copy()  # subsampled statistics of the left view
    np.savez_compressed(os.path.join(HERE, "cones_crop_d20.npz"), **out)
    print("wrote", os.path.join(HERE, "cones_crop_d20.npz"), len(cases), "cases")


if __name__ == "__main__":
    main()
