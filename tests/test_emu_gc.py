"""tests/test_gpu_gc.py (pairwise terms + graph-cut move on the device) on the CPU emulator of the kernel source: the graph
construction, the deterministic push-relabel and the host-side schedule, without a GPU."""
import pytest

from emu import emu_lib
import test_gpu_gc as _gc
import test_emu_parity as _ep


@pytest.fixture(scope="module", autouse=True)
def _use_emulator(monkeypatch_module=None):
    import os
    prev = os.environ.get("LEXP_GC_THREADS")
    os.environ["LEXP_GC_THREADS"] = "128"   # fibers per emulated CTA: fewer context switches per barrier, same arithmetic
    with emu_lib.emulated():
        yield
    if prev is None:
        os.environ.pop("LEXP_GC_THREADS", None)
    else:
        os.environ["LEXP_GC_THREADS"] = prev


@pytest.fixture(scope="module")
def devmem():
    return _ep._HostAsDeviceMemory()


def test_emu_smoothness_coefficients():
    _gc.test_smoothness_coefficients_equal_the_oracle()


def test_emu_pairwise_terms():
    _gc.test_pairwise_terms_equal_the_oracle_bit_for_bit()


def test_emu_gc_replay_small(devmem):
    _gc.test_gc_replay_small(devmem)


def test_emu_gc_replay_right_view_strong_smoothness(devmem):
    _gc.test_gc_replay_right_view_strong_smoothness(devmem)


def test_emu_gc_moves_never_raise_the_energy(devmem):
    _gc.test_gc_moves_never_raise_the_energy(devmem)
