"""Randomised differential test of the kernel source (on the CPU emulator) against the numpy oracle: image sizes, filter
radii, filterRect / targetRect pairs (including ones the reference never produces: filterRect smaller than the dependency
cone, targets touching the filterRect border, 1-pixel targets, one-row and one-column rectangles), both views, planes from
benign to degenerate.  Deterministic seeds; the whole file runs in well under a minute."""
import numpy as np
import pytest

from emu import emu_lib
from oracle import lexp_oracle as O
from lexp_testlib import assert_costs_close, make_scene


@pytest.fixture(scope="module", autouse=True)
def _use_emulator():
    with emu_lib.emulated():
        yield


def _random_plane(rs, D):
    kind = rs.integers(0, 8)
    if kind == 0:
        return np.array([0, 0, rs.uniform(-2, D + 1), 0], np.float32)                       # fronto-parallel, maybe out of range
    if kind == 1:
        return np.array([rs.uniform(-3, 3), rs.uniform(-3, 3), rs.uniform(-50, 50), 0], np.float32)  # steep
    if kind == 2:
        return np.array([np.nan, 0, 1, 0], np.float32) if rs.integers(2) else np.array([0, 0, np.inf, 0], np.float32)
    if kind == 3:
        return np.array([1e-4, -1e-4, D - 1 - 1e-3, 0], np.float32)                          # grazes MAX
    a, b = rs.uniform(-0.3, 0.3, 2)
    return np.array([a, b, rs.uniform(0, D - 1), rs.uniform(-1, 1) if kind == 4 else 0], np.float32)


@pytest.mark.parametrize("seed", range(6))
def test_random_rects_and_planes_match_the_oracle(seed):
    import localexpstereo_b200 as L
    rs = np.random.default_rng(1000 + seed)
    H, W = int(rs.integers(40, 140)), int(rs.integers(40, 180))
    D = int(rs.integers(4, 40))
    windR = int(rs.choice([4, 7, 12, 20, 26, 32]))
    eps = float(rs.choice([1e-4, 1e-3, 1e-2]))
    th = float(rs.choice([0.3, 0.5, 2.0]))
    mn = float(rs.choice([0.0, 0.0, -3.0]))
    mx = float(D - 1 + mn)
    imL, imR, volL, volR = make_scene(H, W, D, seed=seed)
    prm = L.Parameters(windR=windR, filterName="GF", filter_param1=eps, th_col=th)
    E = L.CostVolumeEnergy(imL, imR, volL, volR, prm, mx, mn)
    Or = O.CostVolumeEnergyOracle(imL, imR, volL, volR, windR, eps, th, mx, mn)
    worst = 0.0
    try:
        for case in range(14):
            fw, fh = int(rs.integers(1, W + 1)), int(rs.integers(1, H + 1))
            if case % 5 == 0:
                fw, fh = W, H
            fx, fy = int(rs.integers(0, W - fw + 1)), int(rs.integers(0, H - fh + 1))
            tw, th_ = int(rs.integers(1, fw + 1)), int(rs.integers(1, fh + 1))
            if case % 7 == 3:
                tw = th_ = 1
            tx, ty = fx + int(rs.integers(0, fw - tw + 1)), fy + int(rs.integers(0, fh - th_ + 1))
            f, t = (fx, fy, fw, fh), (tx, ty, tw, th_)
            p = _random_plane(rs, D)
            mode = int(rs.integers(0, 2))
            chk = bool(rs.integers(0, 2))
            img = np.full((H, W), -7.0, np.float32)
            view = img[fy:fy + fh, fx:fx + fw]
            (E.ComputeUnaryPotential if chk else E.ComputeUnaryPotentialWithoutCheck)(f, t, view, p, mode=mode)
            ref = (Or.compute_unary_potential if chk else Or.compute_unary_potential_without_check)(f, t, p, mode)
            got = img[ty:ty + th_, tx:tx + tw]
            untouched = img.copy()
            untouched[ty:ty + th_, tx:tx + tw] = -7.0
            assert (untouched == -7.0).all(), f"wrote outside targetRect: f={f} t={t}"
            if np.isnan(ref).any():
                assert np.array_equal(np.isnan(got), np.isnan(ref)), (f, t, p)
                continue
            worst = max(worst, assert_costs_close(got, ref, f"seed {seed} case {case}: f={f} t={t} plane={p} mode={mode} chk={chk} windR={windR}"))
    finally:
        E.close()
    print("worst rel err", worst)


@pytest.mark.parametrize("seed", range(4))
def test_random_batches_with_forced_tilings(seed, monkeypatch):
    """Batched evaluations whose cells are cut into many work items (LEXP_TILE_OH forces short row segments, wide targets
    force column splits): every tiling must give the oracle's result; NaiveStereoEnergy on odd seeds."""
    import localexpstereo_b200 as L
    rs = np.random.default_rng(2000 + seed)
    naive = bool(seed & 1)
    H, W = int(rs.integers(90, 150)), int(rs.integers(150, 260))
    D = 24
    windR = int(rs.choice([8, 20]))
    monkeypatch.setenv("LEXP_TILE_OH", str(int(rs.choice([8, 16, 24, 48]))))
    imL, imR, volL, volR = make_scene(H, W, D, seed=10 + seed)
    if naive:
        prm = L.Parameters(windR=windR, filterName="GF", filter_param1=1e-4)
        E = L.NaiveStereoEnergy(imL, imR, prm, D - 1)
        Or = O.NaiveStereoEnergyOracle(imL, imR, windR, 1e-4, prm.th_col, prm.th_grad, prm.alpha, D - 1)
    else:
        prm = L.Parameters(windR=windR, filterName="GF", filter_param1=1e-4, th_col=0.5)
        E = L.CostVolumeEnergy(imL, imR, volL, volR, prm, D - 1)
        Or = O.CostVolumeEnergyOracle(imL, imR, volL, volR, windR, 1e-4, 0.5, D - 1)
    try:
        # disjoint targets on a coarse grid, filterRect = target +- windR clipped (LayerManager's rule), some much wider than a tile
        rects = []
        for gy in range(0, H, 60):
            for gx in range(0, W, 130):
                tw, th_ = int(rs.integers(5, 125)), int(rs.integers(5, 55))
                tx, ty = gx + int(rs.integers(0, 5)), gy + int(rs.integers(0, 5))
                tw, th_ = min(tw, W - tx), min(th_, H - ty)
                if tw <= 0 or th_ <= 0:
                    continue
                fx0, fy0 = max(tx - windR, 0), max(ty - windR, 0)
                fx1, fy1 = min(tx + tw + windR, W), min(ty + th_ + windR, H)
                rects.append(((fx0, fy0, fx1 - fx0, fy1 - fy0), (tx, ty, tw, th_)))
        rngp = O.CvRNG(77 + seed)
        planes = np.stack([O.create_random_label(rngp, t[0], t[1], 0.0, D - 1.0) for _, t in rects])
        for mode in (0, 1):
            img = np.full((H, W), -7.0, np.float32)
            E.ComputeUnaryPotentialBatch([f for f, _ in rects], [t for _, t in rects], img, planes, mode=mode)
            nbad = ntot = 0
            for (f, t), p in zip(rects, planes):
                got = img[t[1]:t[1] + t[3], t[0]:t[0] + t[2]]
                ref = Or.compute_unary_potential(f, t, p, mode)
                if naive:  # closed-form vs LU inverse affine are the same here (oracle == kernel formula): plain tolerance
                    assert_costs_close(got, ref, f"naive batch f={f} t={t}")
                else:
                    assert_costs_close(got, ref, f"batch f={f} t={t}")
    finally:
        E.close()


def test_one_large_cell_is_cut_into_many_work_items():
    """A layer-2-like cell (target 300 x 200): several column tiles x row segments, each with its own 2R warm-up, must
    reassemble to the oracle's result without seams."""
    import localexpstereo_b200 as L
    H, W, D, windR = 250, 350, 20, 20
    imL, imR, volL, volR = make_scene(H, W, D, seed=77)
    prm = L.Parameters(windR=windR, filterName="GF", filter_param1=1e-4, th_col=0.5)
    E = L.CostVolumeEnergy(imL, imR, volL, volR, prm, D - 1)
    Or = O.CostVolumeEnergyOracle(imL, imR, volL, volR, windR, 1e-4, 0.5, D - 1)
    try:
        f, t = (5, 5, 340, 240), (25, 25, 300, 200)
        plan = E.make_plan([f], [t])
        assert plan.num_items >= 8
        p = np.array([0.031, -0.022, 8.5, 0], np.float32)
        img = np.full((H, W), -7.0, np.float32)
        plan.eval_host(p[None, :], img, True, 0)
        plan.close()
        ref = Or.compute_unary_potential(f, t, p, 0)
        got = img[t[1]:t[1] + t[3], t[0]:t[0] + t[2]]
        worst = assert_costs_close(got, ref, "large cell")
        # seams: the error must not concentrate at tile borders (it would if a warm-up row or halo column were missing)
        err = np.abs(got.astype(np.float64) - ref) / np.maximum(np.abs(ref), 1e-3)
        assert err.max() < 3e-5, err.max()
        print("items", "worst rel err", worst)
    finally:
        E.close()
