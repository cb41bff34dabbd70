"""(Runs last in the -m gpu suite: the 1436 x 992 x 290 case ended in an illegal address on the device once this round --
profiles/r2_gpu_validation.md -- and a fault there must not keep the other tests from being recorded.)

GPU parity at the other BASELINE.json sizes: Adirondack-shaped 1436 x 992 x 290 (configs[1]), an odd
size that is not a multiple of the 4 x 4 blocked volume layout (1437 x 991), and the 4K stress case 3840 x 2160 x 512 with
filterRadius 32 (configs[4]: the R = 16 kernel instantiation, 1099 x 1099 filterRects in layer 2, 17 GB volume) -- sampled cells
of all three layers against the numpy oracle, through the C-ABI, tolerance 1e-4 relative and exact COST_FOR_INVALID mask."""
import numpy as np
import pytest

from oracle import lexp_oracle as O
from lexp_testlib import assert_costs_close
from test_gpu_golden import _OffsetVolume

pytestmark = pytest.mark.gpu


def _oracle_cell(img, stats, slab, f, t, p, windR, D, th=0.5):
    """The oracle's unary costs of one call, from the filterRect slab of the volume and the whole-image statistics."""
    Or = O.CostVolumeEnergyOracle(img[f[1]:f[1] + f[3], f[0]:f[0] + f[2]], None, slab, None, windR, 1e-4, th, D - 1)
    sl = (slice(f[1], f[1] + f[3]), slice(f[0], f[0] + f[2]))
    Or.filter[0].I = [c[sl] for c in stats.I]; Or.filter[0].mean = [c[sl] for c in stats.mean]; Or.filter[0].inv = [c[sl] for c in stats.inv]
    fl, tl = (0, 0, f[2], f[3]), (t[0] - f[0], t[1] - f[1], t[2], t[3])
    raw = O.sample_plane_cost(_OffsetVolume(slab, f[0], f[1]), f, p, np.float32(th), 0.0, D - 1)
    q = O.guided_filter_sub(Or.filter[0], fl, raw)[tl[1]:tl[1] + tl[3], tl[0]:tl[0] + tl[2]]
    valid = O.is_valid_label(p, t, 0.0, D - 1)
    q = q.copy(); q[~valid] = O.COST_FOR_INVALID
    return q


@pytest.mark.parametrize("W,H,D,windR,picks", [
    (1436, 992, 290, 20, (4, 2, 2)),      # BASELINE configs[1] shape
    (1437, 991, 64, 20, (4, 2, 2)),       # ragged edge of the 4 x 4 blocked volume layout
    (3840, 2160, 512, 32, (3, 2, 1)),     # BASELINE configs[4]: R = 16, 1099^2 filterRects, 2 x 17 GB of volume on the device
])
def test_sampled_cells_of_every_layer(W, H, D, windR, picks):
    import torch
    import localexpstereo_b200 as L
    from localexpstereo_b200 import synth
    from localexpstereo_b200.sweep import v3_layer_units
    g = torch.Generator(device="cuda").manual_seed(99)
    vol = torch.rand((D, H, W), generator=g, device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()   # torch fills the volume on its own stream; the library's stream is non-blocking
    img = synth.synthetic_image(H, W, 42)
    prm = L.Parameters(windR=windR, filterName="GF", filter_param1=1e-4, th_col=0.5)
    E = L.CostVolumeEnergy(img, None, vol, None, prm, D - 1)
    try:
        stats = O.GuidedFilterStats(img, windR // 2, 1e-4)
        lm = L.LayerManager(W, H, windR)
        rng = np.random.default_rng(17)
        worst, biggest = 0.0, 0
        for li, u in enumerate(v3_layer_units(W)):
            lay = lm.addLayer(u)
            ng = len(lay.disjointRegionSets)
            # the last group holds the right/bottom border cells (ragged 4 x 4 blocks), a random one the interior
            for gi in {ng - 1, int(rng.integers(ng))}:
                grp = lay.disjointRegionSets[gi]
                planes = synth.synthetic_planes(lay.unitRegions, 2, D, 11 + li)[1][grp]
                fr = [lay.filterRegions[r] for r in grp]
                tr = [lay.sharedRegions[r] for r in grp]
                out = np.zeros((H, W), np.float32)
                E.ComputeUnaryPotentialBatch(fr, tr, out, planes)
                # always include the cell with the largest filterRect and the one that reaches furthest right/down
                order = sorted(range(len(grp)), key=lambda i: (fr[i][2] * fr[i][3], fr[i][0] + fr[i][1]), reverse=True)
                pick = set(order[:1]) | set(int(i) for i in rng.choice(len(grp), size=min(len(grp), picks[li]), replace=False))
                for i in pick:
                    f, t, p = fr[i], tr[i], planes[i]
                    biggest = max(biggest, f[2] * f[3])
                    slab = vol[:, f[1]:f[1] + f[3], f[0]:f[0] + f[2]].cpu().numpy()
                    q = _oracle_cell(img, stats, slab, f, t, p, windR, D)
                    worst = max(worst, assert_costs_close(out[t[1]:t[1] + t[3], t[0]:t[0] + t[2]], q, f"{W}x{H} layer {li} cell {grp[i]}"))
        print(f"{W}x{H}x{D} r{windR}: worst rel err {worst:.2e}, largest filterRect {biggest} px")
        if windR == 32:
            assert biggest >= 1099 * 1099
    finally:
        E.close()
