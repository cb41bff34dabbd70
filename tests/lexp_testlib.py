"""Shared helpers for the parity tests: seeded synthetic scenes and comparison rules."""
import numpy as np

from oracle import lexp_oracle as O

REL_TOL = 1e-4      # north_star: "within 1e-4 relative float tolerance on aggregated costs"
ABS_FLOOR = 1e-3    # |delta| <= 1e-4 * max(|ref|, 1e-3)


def make_scene(H, W, D, seed=0, natural=None):
    img = natural if natural is not None else O.synthetic_image(H, W, 42 + seed)
    imgR = O.synthetic_image(H, W, 43 + seed) if natural is None else natural[:, ::-1].copy()
    volL = O.synthetic_volume(D, H, W, 1234 + seed)
    volR = O.synthetic_volume(D, H, W, 1235 + seed)
    return img, imgR, volL, volR


def assert_costs_close(got, ref, what=""):
    got = np.asarray(got)
    ref = np.asarray(ref)
    inv_ref = ref == O.COST_FOR_INVALID
    inv_got = got == O.COST_FOR_INVALID
    assert np.array_equal(inv_ref, inv_got), f"{what}: COST_FOR_INVALID mask differs at {np.argwhere(inv_ref != inv_got)[:5]}"
    ok = ~inv_ref
    if ok.any():
        err = np.abs(got[ok].astype(np.float64) - ref[ok].astype(np.float64))
        tol = REL_TOL * np.maximum(np.abs(ref[ok]), ABS_FLOOR)
        worst = float((err / tol).max())
        assert worst <= 1.0, f"{what}: max err/tol = {worst:.3f} (max abs err {err.max():.3e})"
        return worst * REL_TOL
    return 0.0


def random_planes(rng: O.CvRNG, units, D):
    out = []
    for (ux, uy, uw, uh) in units:
        n = rng.uniform_int(0, uw * uh)
        out.append(O.create_random_label(rng, ux + n % uw, uy + n // uw, 0.0, float(D - 1)))
    return np.stack(out)
