"""The reference's own FastGCStereo loop, compiled from its headers with include/CudaCostVolumeEnergy.h installed as the
StereoEnergy, running against the KERNEL SOURCE on the CPU emulator (oracle/_ref/dropin_check_emu = oracle/dropin_check.cpp
linked to tests/emu/liblexp_emu.so): every proposal's costs are compared with the reference's CPU energy.  The same program
linked to liblexp_cuda.so is the GPU test tests/test_gpu_zz_dropin.py."""
import json
import os
import subprocess

import pytest

from oracle import build_ref

if not (build_ref.available() or build_ref.reference_present()):
    pytest.skip("oracle/_ref is not built and the reference sources are not on this machine", allow_module_level=True)


@pytest.mark.parametrize("energy,threads", [("CostVolumeEnergy", 1), ("NaiveStereoEnergy", 1), ("CostVolumeEnergy", 6), ("CostVolumeEnergy", "batched")])
def test_reference_loop_through_the_adapter_on_the_emulator(energy, threads):
    from emu import emu_lib
    emu_lib.load()           # builds tests/emu/liblexp_emu.so if needed
    assert build_ref.build() is not None
    if not os.path.exists(build_ref.DROPIN_EMU):
        pytest.skip("dropin_check_emu was not built")
    # "batched": the loop restructured as in INTEGRATION.md section 3 (CudaCostVolumeEnergy::GroupPlan, one evaluation per step)
    batched = threads == "batched"
    threads = 1 if batched else threads
    cmd = [build_ref.DROPIN_EMU, "--W", "64", "--H", "56", "--K", "1", "--threads", str(threads)] + (["--naive"] if energy == "NaiveStereoEnergy" else []) + (["--batched"] if batched else [])
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    d = json.loads(res.stdout.strip().splitlines()[-1])
    print(d)
    assert "error" not in d, d
    assert d["under_test"].startswith("CudaCostVolumeEnergy") and d["move_calls"] > 1000 and ("GroupPlan" in d["under_test"]) == batched
    assert d["mask_mismatch"] == 0 and d["ok"] is True and res.returncode == 0, d
    if energy == "CostVolumeEnergy":
        assert d["out_of_tolerance"] == 0 and d["worst_err_over_tol"] < 0.5, d
    if threads > 1:  # the reference's OpenMP loop: concurrent calls of the virtual are served by combined launches
        assert d["combined_calls"] > d["move_calls"] // 10 and d["combined_launches"] < d["combined_calls"], d
