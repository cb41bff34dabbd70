"""tests/test_gpu_naive.py (NaiveStereoEnergy path) executed on the CPU emulator of the kernel source -- see tests/test_emu_parity.py."""
import pytest

from emu import emu_lib
import test_gpu_naive as _n


@pytest.fixture(scope="module", autouse=True)
def _use_emulator():
    with emu_lib.emulated():
        yield


scene = _n.scene

test_emu_naive_cells_match_oracle = _n.test_naive_cells_match_oracle
test_emu_naive_virtuals_and_edge_planes = _n.test_naive_virtuals_and_edge_planes
test_emu_naive_matches_reference_minted_vectors = _n.test_naive_matches_reference_minted_vectors
