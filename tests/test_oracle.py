"""CPU tests of the oracle itself: against the OpenCV kernels the reference calls (cv2 wheel), against the
independent C restatement, and against the committed golden vectors."""
import numpy as np
import pytest

from oracle import lexp_oracle as O
from lexp_testlib import assert_costs_close

cv2 = pytest.importorskip("cv2")


def test_box_filter_matches_cv_boxfilter():
    """GuidedFilter.h:40-45: cv::boxFilter(ksize 2R+1, normalize=false, BORDER_CONSTANT)."""
    rng = np.random.default_rng(0)
    for (h, w, R) in [(57, 83, 5), (30, 41, 10), (9, 9, 10), (100, 100, 16)]:
        X = rng.random((h, w))
        ref = cv2.boxFilter(X, -1, (2 * R + 1, 2 * R + 1), None, (-1, -1), False, cv2.BORDER_CONSTANT)
        assert np.abs(O.box_sum(X, R) - ref).max() < 1e-10
        assert np.abs(O.box_sum_fast(X, R) - ref).max() < 1e-10
        Xf = X.astype(np.float32)  # 32F source: OpenCV accumulates in double and rounds once
        reff = cv2.boxFilter(Xf, -1, (2 * R + 1, 2 * R + 1), None, (-1, -1), False, cv2.BORDER_CONSTANT)
        assert np.array_equal(O.box_sum_fast(Xf, R), reff)


def test_subregion_window_counts():
    """N = boxfilter(ones(rect.size())) (GuidedFilter.h:324)."""
    for rect, R in [((3, 4, 37, 25), 10), ((0, 0, 12, 50), 10), ((0, 0, 100, 100), 16)]:
        ones = np.ones((rect[3], rect[2]))
        ref = cv2.boxFilter(ones, -1, (2 * R + 1, 2 * R + 1), None, (-1, -1), False, cv2.BORDER_CONSTANT)
        assert np.array_equal(O.subregion_N(rect, R), ref)


def test_channel_sum_order_matches_cv_reduce():
    """IsValiLabel's ds = channelSum(coordinates.mul(label)) (Utilities.hpp:224-229): the float sum is left to right."""
    rng = np.random.default_rng(1)
    M = (rng.standard_normal((50000, 4)) * np.array([1e3, 1e3, 1e2, 0])).astype(np.float32)
    r = cv2.reduce(M, 1, cv2.REDUCE_SUM).ravel()
    assert np.array_equal(r, ((M[:, 0] + M[:, 1]) + M[:, 2]) + M[:, 3])


def test_subregion_filter_equals_whole_image_filter_inside_margin():
    """Comment invariant GuidedFilter.h:298-300."""
    img = O.synthetic_image(90, 120, 5)
    vol = O.synthetic_volume(12, 90, 120, 6)
    E = O.CostVolumeEnergyOracle(img, None, vol, None, 20, 1e-4, 0.5, 11)
    p = O.create_plane(0.1, -0.05, 0.99, 5.0, 60, 45)
    f, t = (10, 5, 100, 80), (30, 25, 60, 40)
    a = E.compute_unary_potential(f, t, p)
    b = E.compute_unary_potential((0, 0, 120, 90), t, p)
    assert np.array_equal(a, b)


def test_c_oracle_matches_numpy_oracle():
    from oracle.c_oracle import COracle
    H, W, D = 96, 128, 16
    img = O.synthetic_image(H, W, 42)
    vol = O.synthetic_volume(D, H, W, 1234)
    E = O.CostVolumeEnergyOracle(img, None, vol, None, 20, 1e-4, 0.5, D - 1)
    Cc = COracle(H, W, D, 20, 1e-4, 0.5, D - 1)
    Cc.set_image(0, img)
    Cc.set_volume(0, vol)
    s_np, s_c = E.filter[0].stats_f32(), Cc.stats(0)
    assert (np.abs(s_np - s_c) / np.abs(s_np[3:]).max(axis=0, keepdims=True).clip(1)).max() < 1e-6
    L = O.make_layer(W, H, 20, 12)
    rng = O.CvRNG(5)
    for r in [0, 5, 10, 33, 44, 79, 87]:
        u = L["unit"][r]
        pl = O.create_random_label(rng, u[0], u[1], 0, D - 1)
        assert np.array_equal(E.raw(L["filter"][r], pl), Cc.sample(0, L["filter"][r], pl))
        a = E.compute_unary_potential(L["filter"][r], L["shared"][r], pl)
        b = Cc.unary(0, L["filter"][r], L["shared"][r], pl)
        assert_costs_close(b, a, f"cell {r}")
        assert np.abs(a - b)[a != O.COST_FOR_INVALID].max(initial=0) < 1e-6
    # batched (OpenMP) form writes only the target rects
    g = L["groups"][3]
    out = np.full((H, W), -1.0, np.float32)
    pls = np.stack([O.create_random_label(rng, *L["unit"][r][:2], 0, D - 1) for r in g])
    Cc.unary_batch(0, [L["filter"][r] for r in g], [L["shared"][r] for r in g], pls, out)
    for r, p in zip(g, pls):
        t = L["shared"][r]
        assert_costs_close(out[t[1]:t[1] + t[3], t[0]:t[0] + t[2]], E.compute_unary_potential(L["filter"][r], t, p), f"batch {r}")
    Cc.close()


def test_oracles_reproduce_golden_vectors():
    import lexp_golden
    from oracle.c_oracle import COracle
    G = lexp_golden.load()
    E = O.CostVolumeEnergyOracle(G["imL"], G["imR"], G["volL"], G["volR"], G["windR"], G["eps"], G["th"], G["D"] - 1)
    H, W = G["imL"].shape[:2]
    Cc = COracle(H, W, G["D"], G["windR"], G["eps"], G["th"], G["D"] - 1)
    for m, (im, vol) in enumerate(((G["imL"], G["volL"]), (G["imR"], G["volR"]))):
        Cc.set_image(m, im)
        Cc.set_volume(m, vol)
    # the golden file is minted by the compiled reference (tests/golden/make_golden.py): the restatements agree with it to
    # the last float bit or one off (order of the double box sums)
    st = E.filter[0].stats_f32()[:, ::8, ::8]
    assert np.abs(st - G["stats0"]).max() <= 1e-6 * np.abs(G["stats0"]).max()
    for i, c in enumerate(G["cases"]):
        fn = E.compute_unary_potential if c["check"] else E.compute_unary_potential_without_check
        got = fn(c["frect"], c["trect"], c["plane"], c["mode"])
        assert np.array_equal(np.isnan(got), np.isnan(c["ref"])) and np.array_equal(got == O.COST_FOR_INVALID, c["ref"] == O.COST_FOR_INVALID)
        ok = np.isfinite(c["ref"]) & (c["ref"] != O.COST_FOR_INVALID)
        assert not ok.any() or (np.abs(got[ok].astype(np.float64) - c["ref"][ok]) <= 2e-7 * np.maximum(np.abs(c["ref"][ok]), 1e-3)).all(), \
            f"numpy oracle drifted from golden case {i}"
        got_c = Cc.unary(c["mode"], c["frect"], c["trect"], c["plane"], c["check"])
        if np.isnan(c["ref"]).any():
            assert np.array_equal(np.isnan(got_c), np.isnan(c["ref"]))
        else:
            assert_costs_close(got_c, c["ref"], f"C oracle, golden case {i}")
    Cc.close()


def test_volume_preparation():
    """fillOutOfView / convertVolumeL2R (main.cpp:146-199) on a volume with recognisable entries."""
    D, H, W = 5, 3, 12
    vol = (np.arange(D)[:, None, None] * 1000 + np.arange(H)[None, :, None] * 100 + np.arange(W)[None, None, :]).astype(np.float32)
    L = O.fill_out_of_view(vol.copy(), 0)
    for d in range(D):
        assert (L[d, :, :d] == vol[d, :, d:d + 1]).all() and (L[d, :, d:] == vol[d, :, d:]).all()
    R = O.convert_volume_l2r(vol)
    for d in range(D):
        assert (R[d, :, :W - d] == vol[d, :, d:]).all()         # volR[d][y][x] = volL[d][y][x+d]
        assert (R[d, :, W - 1 - d:] == vol[d, :, W - 1:W]).all()  # right edge replicated (main.cpp:192-195)
    R2 = O.fill_out_of_view(R.copy(), 1)
    for d in range(1, D):
        assert (R2[d, :, W - d:] == R[d, :, W - d - 1:W - d]).all()


def test_validity_mask_and_branches():
    D = 16
    r = (10, 20, 30, 25)
    assert O.is_valid_label(np.array([0, 0, 5.0, 0], np.float32), r, 0, D - 1).all()
    assert not O.is_valid_label(np.array([0, 0, -1.0, 0], np.float32), r, 0, D - 1).any()
    m = O.is_valid_label(np.array([0.5, 0, 0.0, 0], np.float32), r, 0, D - 1)  # ds = 0.5 x; corners need ds - 2.5 >= 0
    assert m.any() and not m.all() and (m == (0.5 * np.arange(10, 40) + 2.5 <= D - 1)[None, :]).all()
    assert not O.is_valid_label(np.array([0, 0, 5.0, np.inf], np.float32), r, 0, D - 1).any()   # 0 * inf = NaN in channelSum
    assert O.is_valid_label(np.array([0, 0, 5.0, np.inf], np.float32), (3, 3, 1, 1), 0, D - 1).all()  # 1x1 path ignores v
    vol = O.synthetic_volume(D, 40, 50, 3)
    p = O.sample_plane_cost(vol, (0, 0, 50, 40), np.array([np.nan, 0, 0, 0], np.float32), 1e9)
    assert (p == O.COST_FOR_INVALID).all()
    p = O.sample_plane_cost(vol, (0, 0, 50, 40), np.array([0, 0, -2.0, 0], np.float32), 1e9)
    assert np.array_equal(p, vol[0])
    p = O.sample_plane_cost(vol, (0, 0, 50, 40), np.array([0, 0, D - 1.0, 0], np.float32), 1e9)
    assert np.array_equal(p, vol[D - 1])
    p = O.sample_plane_cost(vol, (0, 0, 50, 40), np.array([0, 0, 3.25, 0], np.float32), 0.4)
    assert np.array_equal(p, np.minimum(np.float32(0.75) * vol[3] + np.float32(0.25) * vol[4], np.float32(0.4)))


def test_plane_helpers_and_rng():
    p = O.create_plane(0.1, 0.2, 0.9, 7.0, 3, 4)
    assert abs(float(O.plane_get_z(p, 3, 4)) - 7.0) < 1e-5
    n = O.plane_normal(p)
    assert abs(np.linalg.norm(n) - 1) < 1e-6 and n[2] > 0
    r1, r2 = O.CvRNG(7), O.CvRNG(7)
    assert [r1.next() for _ in range(5)] == [r2.next() for _ in range(5)]
    u = [O.CvRNG(9).uniform_float(2.0, 3.0) for _ in range(3)]
    assert all(2.0 <= x < 3.0 for x in u)
    pl = O.create_random_label(O.CvRNG(3), 10, 20, 0.0, 63.0)
    assert np.isfinite(pl).all() and 0 <= float(O.plane_get_z(pl, 10, 20)) < 63.0 + 1e-3


def test_naive_energy_oracle_against_opencv_kernels():
    """NaiveStereoEnergy restatement vs the OpenCV kernels the reference calls (StereoEnergy.h:651-654,727-729)."""
    rng = np.random.default_rng(3)
    im = O.synthetic_image(70, 90, 8)
    ex = O.build_exI(im, 0.9)
    I = im.astype(np.float32)
    gray = cv2.cvtColor(I, cv2.COLOR_BGR2GRAY)
    gx = cv2.Sobel(gray, cv2.CV_32F, 1, 0, ksize=1, scale=0.5, borderType=cv2.BORDER_REPLICATE)
    ref = np.dstack([I * np.float32(1.0 - float(np.float32(0.9))), gx * np.float32(0.9)]).astype(np.float32)
    assert np.abs(ex - ref).max() < 5e-5  # cv2 4.x evaluates the gray dot product with FMAs: <= 1 ulp of 255
    # fixed-point warp: exact for a given inverse matrix
    src = (rng.random((40, 60, 4)) * 100).astype(np.float32)
    for a, tx in [(1.0, 0.3), (0.93, 2.11), (1.21, -5.6)]:
        ref = cv2.warpAffine(src, np.array([[a, 0, tx], [0, 1, 0]]), (50, 30), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP,
                             borderMode=cv2.BORDER_REPLICATE)
        assert np.array_equal(O.warp_affine_linear_replicate(src, np.array([a, 0, tx, 0, 1, 0.0]), 50, 30), ref)
    # end to end vs getAffineTransform + warpAffine: identical except at exact 1/32-pixel rounding ties
    r2 = O.CvRNG(3)
    nd = npx = 0
    for _ in range(30):
        fx, fy = r2.uniform_int(0, 40), r2.uniform_int(0, 20)
        fr = (fx, fy, 45, 40)
        pl = O.create_random_label(r2, fx + 20, fy + 20, 0, 31.0)
        for mode in (0, 1):
            sign = np.float32(-1.0 if mode else 1.0)
            x00, y00 = np.float32(fx), np.float32(fy)
            x11, y11 = np.float32(x00 + 45), np.float32(y00 + 40)
            gz = lambda xx, yy: O.plane_get_z(pl, xx, yy)
            s3 = np.array([[x00 - sign * gz(x00, y00), y00], [x00 - sign * gz(x00, y11), y11], [x11 - sign * gz(x11, y00), y00]], np.float32)
            d3 = np.array([[0, 0], [0, y11 - y00], [x11 - x00, 0]], np.float32)
            refw = cv2.warpAffine(ex, cv2.getAffineTransform(s3, d3), (45, 40), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)
            mine = O.warp_affine_linear_replicate(ex, O.affine_inverse_for_plane(fr, pl, mode), 45, 40)
            d = np.abs(mine - refw).max(axis=2)
            nd += int((d > 1e-4).sum()); npx += d.size
    assert nd / npx < 1e-3, (nd, npx)


def test_naive_oracle_reproduces_reference_minted_vectors():
    """tests/golden/cones_crop_naive.npz holds outputs of the reference's own NaiveStereoEnergy (compiled, oracle/build_ref.py)."""
    import lexp_golden
    G = lexp_golden.load_naive()
    E = O.NaiveStereoEnergyOracle(G["imL"], G["imR"], G["windR"], G["eps"], G["th_col"], G["th_grad"], G["alpha"], G["D"] - 1)
    assert np.array_equal(E.ExI[0][::4, ::4], G["exi0"])
    nbad = ntot = 0
    for i, c in enumerate(G["cases"]):
        fn = E.compute_unary_potential if c["check"] else E.compute_unary_potential_without_check
        got = fn(c["frect"], c["trect"], c["plane"], c["mode"])
        inv = c["ref"] == O.COST_FOR_INVALID
        assert np.array_equal(got == O.COST_FOR_INVALID, inv), i
        err = np.abs(got[~inv].astype(np.float64) - c["ref"][~inv]) / np.maximum(np.abs(c["ref"][~inv]), 1e-3)
        nbad += int((err > 1e-4).sum()); ntot += int((~inv).sum())
    # the oracle repeats getAffineTransform's LU solve and warpAffine's inversion operation by operation: no outliers allowed
    assert ntot > 20000 and nbad == 0, (nbad, ntot)
