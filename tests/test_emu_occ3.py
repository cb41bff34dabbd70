"""The build-time kernel variant LEXP_OCC3 (three CTAs per SM: 56 registers, gather batches of 4, rolling statistics
prefetch in team C, no register history in team H, 75 KB shared-memory cap in the planner; lexp_kernels.cuh) on the CPU
emulator: same parity tests as the shipped kernel.  The variant is a prepared experiment -- its speed is unknown until it is
measured on a B200 (scripts/gpu_variants.sh); this file only makes sure that whatever is measured is correct."""
import numpy as np
import pytest

from emu import emu_lib
import test_gpu_parity as _p
import test_gpu_golden as _g
import test_gpu_naive as _n


@pytest.fixture(scope="module", autouse=True)
def _use_emulator():
    with emu_lib.emulated(variant="occ3"):
        yield


scene = _p.scene

test_occ3_cells_of_a_layer = _p.test_cells_of_a_layer
test_occ3_single_cell_virtuals = _p.test_single_cell_virtuals
test_occ3_branches_of_the_sampler = _p.test_branches_of_the_sampler
test_occ3_filter_rect_smaller_than_dependency_cone = _p.test_filter_rect_smaller_than_dependency_cone
test_occ3_other_filter_radii = _p.test_other_filter_radii
test_occ3_nonzero_min_disparity_and_odd_max = _p.test_nonzero_min_disparity_and_odd_max
test_occ3_golden_vectors_through_the_c_abi = _g.test_golden_vectors_through_the_c_abi


def test_occ3_equals_the_shipped_kernel_bit_for_bit(scene, monkeypatch):
    """Same sums in the same order: with the planner forced to the shipped tiling (2 CTA slots per SM, no shared-memory cap)
    the variant must reproduce the shipped kernel's emulated output exactly, under any thread schedule; with its own
    (narrower) tiles the running sums start elsewhere, so only the last bits may move."""
    from oracle import lexp_oracle as O
    L, H, W, D = (scene[k] for k in "L H W D".split())
    imL, imR, volL, volR = _p.make_scene(H, W, D)
    prm = L.Parameters(windR=20, filterName="GF", filter_param1=1e-4, th_col=0.5)
    lay = L.LayerManager(W, H, 20).addLayer(31)
    g = lay.disjointRegionSets[0]
    rng = O.CvRNG(3)
    planes = np.stack([O.create_random_label(rng, *lay.unitRegions[r][:2], 0.0, D - 1.0) for r in g])

    def run(variant, order, same_tiling):
        if same_tiling:
            monkeypatch.setenv("LEXP_SMEM_CAP", "0")
            monkeypatch.setenv("LEXP_CTAS_PER_SM", "2")
        else:
            monkeypatch.delenv("LEXP_SMEM_CAP", raising=False)
            monkeypatch.delenv("LEXP_CTAS_PER_SM", raising=False)
        with emu_lib.emulated(order=order, variant=variant):
            E = L.CostVolumeEnergy(imL, imR, volL, volR, prm, D - 1)
            img = np.full((H, W), -7.0, np.float32)
            E.ComputeUnaryPotentialBatch([lay.filterRegions[r] for r in g], [lay.sharedRegions[r] for r in g], img, planes, mode=1)
            E.close()
        return img

    base = run("", 0, True)
    assert np.array_equal(base, run("occ3", 0, True))
    assert np.array_equal(base, run("occ3", 2, True))
    own = run("occ3", 1, False)
    assert np.array_equal(own == 1e6, base == 1e6)
    ok = base != 1e6
    assert np.abs(own[ok] - base[ok]).max() <= 3e-6 * np.abs(base[ok]).max()


@pytest.fixture(scope="module")
def naive_scene():
    yield from _n.scene.__wrapped__() if hasattr(_n.scene, "__wrapped__") else _n.scene.__pytest_wrapped__.obj()


def test_occ3_naive_matches_reference_minted_vectors(naive_scene):
    _n.test_naive_matches_reference_minted_vectors(naive_scene)


@pytest.mark.parametrize("seed", [3, 11])
def test_occ3_fuzz(seed):
    """A slice of tests/test_emu_fuzz.py on this variant (random rects, radii, planes; the per-link buffer strides and the
    75 KB tiling are what differ from the shipped kernel)."""
    import test_emu_fuzz as _f
    _f.test_random_rects_and_planes_match_the_oracle(seed)
