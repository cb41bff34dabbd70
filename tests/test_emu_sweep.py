"""bench.py's sweep object (localexpstereo_b200/sweep.py) on the CPU emulator: the work accounting the benchmark reports
(evals, target pixels, algorithmic bytes B_alg = 20 F + 36 A + 4 S) against an independent count, the rank shards, and one
complete small sweep (3 layers x all groups x K steps) against the oracle."""
import numpy as np
import pytest

from emu import emu_lib
from oracle import lexp_oracle as O
from lexp_testlib import assert_costs_close, make_scene


@pytest.fixture(scope="module", autouse=True)
def _use_emulator():
    with emu_lib.emulated():
        yield


def test_sweep_accounting_shards_and_a_full_small_sweep():
    import localexpstereo_b200 as L
    from localexpstereo_b200.sweep import UnarySweep
    from localexpstereo_b200 import synth
    H, W, D, windR = 110, 150, 12, 20
    imL, imR, volL, volR = make_scene(H, W, D, seed=2)
    prm = L.Parameters(windR=windR, filterName="GF", filter_param1=1e-4, th_col=0.5)
    E = L.CostVolumeEnergy(imL, imR, volL, volR, prm, D - 1)
    Or = O.CostVolumeEnergyOracle(imL, imR, volL, volR, windR, 1e-4, 0.5, D - 1)
    units, steps = [7, 21, 50], [3, 2, 1]
    sw = UnarySweep(E, unit_sizes=units, steps=steps)
    R = windR // 2
    # independent accounting from the oracle's geometry
    F = S = B = 0
    for u, K in zip(units, steps):
        lay = O.make_layer(W, H, windR, u)
        for f, t in zip(lay["filter"], lay["shared"]):
            ax0, ax1 = max(t[0] - R, f[0]), min(t[0] + t[2] + R, f[0] + f[2])
            ay0, ay1 = max(t[1] - R, f[1]), min(t[1] + t[3] + R, f[1] + f[3])
            A = (ax1 - ax0) * (ay1 - ay0)
            F += K * f[2] * f[3]; S += K * t[2] * t[3]; B += K * (20 * f[2] * f[3] + 36 * A + 4 * t[2] * t[3])
    assert (sw.total_filter_px, sw.total_target_px, sw.local_alg_bytes) == (F, S, B)
    assert sw.local_filter_px == F and sw.launches_per_sweep == sum(g.n_steps for g in sw.groups)
    # two shards partition the work
    parts = [UnarySweep(E, unit_sizes=units, steps=steps, rank=r, world=2) for r in range(2)]
    assert sum(p.local_filter_px for p in parts) == F and sum(p.local_alg_bytes for p in parts) == B
    assert all(p.total_filter_px == F for p in parts)
    for p in parts:
        p.close()
    # one complete sweep through the device-image entry point (device memory = host memory on the emulator)
    cost = np.full((H, W), -7.0, np.float32)
    worst, n = 0.0, 0
    for g in sw.groups:
        lay = sw.layer(g.layer)
        planes = synth.synthetic_planes(lay.unitRegions, g.n_steps, D, 7 + g.layer)[:, g.cells, :]
        for k in range(g.n_steps):
            g.plan.eval_device(np.ascontiguousarray(planes[k]), cost.ctypes.data, W * 4, True, 0)
        E.sync()
        for j, r in enumerate(g.cells[:3]):  # a few cells of every group against the oracle (the last step's planes)
            f, t = lay.filterRegions[r], lay.sharedRegions[r]
            worst = max(worst, assert_costs_close(cost[t[1]:t[1] + t[3], t[0]:t[0] + t[2]],
                                                  Or.compute_unary_potential(f, t, planes[g.n_steps - 1][j], 0), f"layer {g.layer} group {g.group} cell {r}"))
            n += 1
    assert n > 50 and E.launch_count >= sw.launches_per_sweep
    print("cells checked", n, "worst rel err", worst)
    sw.close()
    E.close()
