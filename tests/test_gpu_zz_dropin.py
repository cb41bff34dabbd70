"""The drop-in end to end, in the reference's own types (runs last: it is the slowest GPU test).

oracle/_ref/dropin_check is built in the authoring container by oracle/build_ref.py from the reference's headers
(StereoEnergy.h, CostVolumeEnergy.h, LayerManager.h, Proposer.h, where they lie under /root/reference) plus
include/CudaCostVolumeEnergy.h, and linked against liblexp_cuda.so.  It runs the loop of
FastGCStereo::localExpansionMovesForLayer_CPU and evaluates every proposal through the StereoEnergy virtual interface twice:
with the reference's CPU energy and with the CUDA energy behind the adapter (see oracle/dropin_check.cpp)."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "dropin_check")


@pytest.mark.parametrize("energy,threads", [("CostVolumeEnergy", 1), ("NaiveStereoEnergy", 1), ("CostVolumeEnergy", 8), ("CostVolumeEnergy", "batched"),
                                            ("NaiveStereoEnergy", "batched")])
def test_reference_loop_through_the_adapter(energy, threads):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/dropin_check was not built (needs the reference sources at build time)")
    if not os.access(EXE, os.X_OK):  # the snapshot that carried the file here may have dropped the mode bits
        try:
            os.chmod(EXE, 0o755)
        except OSError:
            pytest.skip("oracle/_ref/dropin_check is not executable on this machine")
    # threads > 1: the cells of a group in an OpenMP parallel for, as FastGCStereo.h:30 -- concurrent calls of the virtual,
    # which the library combines into batched launches (lexp_combine_stats)
    # "batched": the loop as INTEGRATION.md section 3 restructures it (CudaCostVolumeEnergy::GroupPlan: one evaluation of all cells
    # of a group per proposal step, image- and tile-shaped host outputs alternating)
    batched = threads == "batched"
    threads = 1 if batched else threads
    cmd = [EXE, "--threads", str(threads)] + (["--naive"] if energy == "NaiveStereoEnergy" else []) + (["--batched"] if batched else [])
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    line = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else ""
    assert line.startswith("{"), (res.returncode, res.stdout[-500:], res.stderr[-2000:])
    d = json.loads(line)
    print(d)
    assert "error" not in d, d
    assert d["energy"] == energy and d["under_test"].startswith("CudaCostVolumeEnergy") and ("GroupPlan" in d["under_test"]) == batched
    assert d["init_calls"] > 1000 and d["move_calls"] > 10000
    assert d["mask_mismatch"] == 0, d
    assert d["ok"] is True and res.returncode == 0, d
    if threads > 1:
        assert d["combined_calls"] > 0 and d["combined_launches"] <= d["combined_calls"], d
