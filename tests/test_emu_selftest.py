"""Does the emulator catch what it claims to catch?  Toy pipelines with named barriers (tests/emu/selftest.cpp): a correct
one, one with a missing EMPTY barrier (data race) and one with a wrong barrier thread count (deadlock)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")


@pytest.fixture(scope="module")
def st(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "libemu_selftest.so")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-DLEXP_EMU", "-fPIC", "-shared", "-fvisibility-inlines-hidden", "-fno-gnu-unique", "-I", HERE, os.path.join(HERE, "selftest.cpp"), "-o", so])
    L = C.CDLL(so)
    L.emu_selftest_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_char_p, C.c_int]

    def run(bug, order, nchunks=9):
        out = np.zeros(32, np.int32)
        msg = C.create_string_buffer(512)
        rc = L.emu_selftest_run(bug, order, nchunks, out.ctypes.data, msg, 512)
        return rc, out, msg.value.decode()
    return run


def expected(nchunks):
    lane = np.arange(32)
    return sum((1000 * c + (31 - lane)) * (c + 1) for c in range(nchunks)).astype(np.int32)


def test_correct_pipeline_is_schedule_independent(st):
    for order in (0, 1, 2):
        rc, out, msg = st(0, order)
        assert rc == 0, msg
        assert np.array_equal(out, expected(9))


def test_missing_barrier_is_detected(st):
    """Without the EMPTY barrier the producer overruns the consumer: depending on the schedule that is either a wrong result
    or an unbalanced FULL barrier (two producer arrivals complete it without the consumer), reported as a protocol error."""
    detected = 0
    for order in (0, 1, 2):
        rc, out, msg = st(1, order)
        detected += int(rc != 0 or not np.array_equal(out, expected(9)))
    assert detected >= 1


def test_wrong_barrier_count_is_reported_as_deadlock(st):
    rc, out, msg = st(2, 0)
    assert rc == 1 and "deadlock" in msg, msg
