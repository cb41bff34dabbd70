"""Loaders for tests/golden/*.npz, minted by running the compiled reference (see tests/golden/make_golden.py)."""
import os

import numpy as np

from oracle import lexp_oracle as O

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cones_crop_d20.npz")


def load():
    z = np.load(PATH)
    imL, imR = z["imL"], z["imR"]
    windR, eps, th, maxd = z["params"]
    H, W = imL.shape[:2]
    D = int(maxd) + 1
    volL = O.fill_out_of_view(O.synthetic_volume(D, H, W, int(z["vol_seed"])), 0)
    volR = O.fill_out_of_view(O.convert_volume_l2r(volL), 1)
    sums = np.array([volL.astype(np.float64).sum(), volR.astype(np.float64).sum()])
    assert np.allclose(sums, z["vol_sums"], rtol=0, atol=1e-6), "regenerated volumes differ from the minted ones"
    cases = []
    for i in range(int(z["n"])):
        cases.append(dict(mode=int(z[f"mode{i}"]), frect=tuple(int(t) for t in z[f"frect{i}"]), trect=tuple(int(t) for t in z[f"trect{i}"]),
                          plane=z[f"plane{i}"], check=bool(z[f"check{i}"]), ref=z[f"ref{i}"]))
    return dict(imL=imL, imR=imR, volL=volL, volR=volR, windR=int(windR), eps=float(eps), th=float(th), D=D, cases=cases,
                stats0=z["stats0"])


NAIVE_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cones_crop_naive.npz")


def load_naive():
    """NaiveStereoEnergy cases on the same image crop (images come from cones_crop_d20.npz)."""
    z0 = np.load(PATH)
    z = np.load(NAIVE_PATH)
    windR, eps, th_col, th_grad, alpha, maxd = z["params"]
    cases = []
    for i in range(int(z["n"])):
        cases.append(dict(mode=int(z[f"mode{i}"]), frect=tuple(int(t) for t in z[f"frect{i}"]), trect=tuple(int(t) for t in z[f"trect{i}"]),
                          plane=z[f"plane{i}"], check=bool(z[f"check{i}"]), ref=z[f"ref{i}"]))
    return dict(imL=z0["imL"], imR=z0["imR"], windR=int(windR), eps=float(eps), th_col=float(th_col), th_grad=float(th_grad),
                alpha=float(alpha), D=int(maxd) + 1, cases=cases, exi0=z["exi0"])
