"""GPU parity of the NaiveStereoEnergy path (BASELINE.json configs[0] `-mode MiddV2`)."""
import numpy as np
import pytest

from oracle import lexp_oracle as O
from lexp_testlib import assert_costs_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene():
    import lexp_golden
    import localexpstereo_b200 as L
    G = lexp_golden.load()  # natural image pair (crop of data/MiddV2/cones)
    imL, imR = G["imL"], G["imR"]
    H, W = imL.shape[:2]
    prm = L.Parameters(lambda_=20, windR=20, filterName="GF", filter_param1=1e-4)  # main.cpp:72 paramsGF; th_col 10, th_grad 2, alpha 0.9
    D = 64
    E = L.NaiveStereoEnergy(imL, imR, prm, D - 1)
    Or = O.NaiveStereoEnergyOracle(imL, imR, 20, 1e-4, prm.th_col, prm.th_grad, prm.alpha, D - 1)
    yield dict(E=E, O=Or, L=L, H=H, W=W, D=D)
    E.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_naive_cells_match_oracle(scene, mode):
    L, E, Or, H, W, D = (scene[k] for k in "L E O H W D".split())
    lay = L.LayerManager(W, H, 20).addLayer(15)  # main.cpp:305
    rng = O.CvRNG(17 + mode)
    worst = 0.0
    for g in lay.disjointRegionSets[:5]:
        planes = np.stack([O.create_random_label(rng, *lay.unitRegions[r][:2], 0.0, D - 1.0) for r in g])
        fr = [lay.filterRegions[r] for r in g]
        tr = [lay.sharedRegions[r] for r in g]
        img = np.full((H, W), -7.0, np.float32)
        E.ComputeUnaryPotentialBatch(fr, tr, img, planes, mode=mode)
        for f, t, p in zip(fr, tr, planes):
            worst = max(worst, assert_costs_close(img[t[1]:t[1] + t[3], t[0]:t[0] + t[2]], Or.compute_unary_potential(f, t, p, mode), f"cell {f}"))
    print("naive worst rel err", worst)


def test_naive_virtuals_and_edge_planes(scene):
    E, Or, H, W, D = (scene[k] for k in "E O H W D".split())
    f, t = (20, 10, 110, 100), (40, 30, 70, 60)
    for p in [(0.0, 0.0, 5.25, 0.0), (0.3, -0.2, 12.0, 0.0), (0.0, 0.0, -40.0, 0.0), (0.0, 0.0, 300.0, 0.0), (-1.2, 0.9, 30.0, 0.0),
              (0.02, 0.01, 9.5, 0.75)]:
        p = np.array(p, np.float32)
        for chk in (False, True):
            img = np.zeros((H, W), np.float32)
            view = img[f[1]:f[1] + f[3], f[0]:f[0] + f[2]]
            (E.ComputeUnaryPotential if chk else E.ComputeUnaryPotentialWithoutCheck)(f, t, view, p)
            ref = (Or.compute_unary_potential if chk else Or.compute_unary_potential_without_check)(f, t, p)
            assert_costs_close(img[t[1]:t[1] + t[3], t[0]:t[0] + t[2]], ref, f"plane {p} chk={chk}")


def test_naive_matches_reference_minted_vectors(scene):
    """tests/golden/cones_crop_naive.npz: outputs of the reference's own NaiveStereoEnergy (compiled from its headers by
    oracle/build_ref.py in the authoring container).  The CUDA path repeats the reference's getAffineTransform LU solve and
    warpAffine inversion operation by operation (naive_inverse_affine), so no pixel may be out of tolerance."""
    import lexp_golden
    G = lexp_golden.load_naive()
    E, H, W = scene["E"], scene["H"], scene["W"]
    assert G["D"] == scene["D"]
    nbad = ntot = 0
    for i, c in enumerate(G["cases"]):
        f, t = c["frect"], c["trect"]
        img = np.zeros((H, W), np.float32)
        view = img[f[1]:f[1] + f[3], f[0]:f[0] + f[2]]
        (E.ComputeUnaryPotential if c["check"] else E.ComputeUnaryPotentialWithoutCheck)(f, t, view, c["plane"], mode=c["mode"])
        got = img[t[1]:t[1] + t[3], t[0]:t[0] + t[2]]
        inv = c["ref"] == O.COST_FOR_INVALID
        assert np.array_equal(got == O.COST_FOR_INVALID, inv), f"case {i}: COST_FOR_INVALID mask"
        err = np.abs(got[~inv].astype(np.float64) - c["ref"][~inv]) / np.maximum(np.abs(c["ref"][~inv]), 1e-3)
        nbad += int((err > 1e-4).sum()); ntot += int((~inv).sum())
    assert ntot > 20000 and nbad == 0, (nbad, ntot)
