"""TEST INFRASTRUCTURE ONLY.  Compile the reference's own hot-path classes into oracle/_ref/liblexp_ref.so.

The reference (t-taniai/LocalExpStereo) is header-only C++ written against OpenCV 3.1 and MSVC.  Its own build
(a Visual Studio solution + NuGet OpenCV) cannot run here, but the path we care about -- GuidedFilter.h, StereoEnergy.h,
CostVolumeEnergy.h, Plane.h, Utilities.hpp, LayerManager.h, Proposer.h -- compiles with g++ from the files *where they
lie* (read-only, under LEXP_REFERENCE_DIR, default /root/reference/LocalExpansionStereo) once two things are supplied:

  1. oracle/cvshim/: this repository's own small implementation of the cv:: calls those headers make;
  2. one dialect fix: FastGuidedImageFilter<Type>::createSubregionFilter (GuidedFilter.h:301-326) names members of its
     dependent base class without `this->`, which MSVC accepts and ISO C++ two-phase lookup rejects.  The fix is applied
     on the fly to a scratch copy that exists only during the compiler run (oracle/_ref/gen/, deleted afterwards), and
     touches nothing but the right-hand sides of that function's assignments (`= R;` -> `= this->R;` ...).

No reference source is copied into the repository or kept on disk; outputs go to oracle/_ref/ only (git-ignored, but
shipped to the GPU box with the other built .so files).  When the reference directory is absent (the GPU box) build()
returns the existing library, if any.
"""
import os
import re
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.environ.get("LEXP_REFERENCE_DIR", "/root/reference/LocalExpansionStereo")
OUT_DIR = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT_DIR, "liblexp_ref.so")
DROPIN = os.path.join(OUT_DIR, "dropin_check")  # the reference's loop + include/CudaCostVolumeEnergy.h, linked to liblexp_cuda.so
DROPIN_EMU = os.path.join(OUT_DIR, "dropin_check_emu")  # same program linked to tests/emu/liblexp_emu.so (kernel source on CPU fibers)
CXX = os.environ.get("LEXP_REF_CXX", "/usr/bin/g++")
ROOT = os.path.dirname(HERE)
SOURCES = [os.path.join(HERE, "ref_driver.cpp"), os.path.join(HERE, "maxflow", "graph.h"), os.path.join(HERE, "refstub", "Evaluator.h"), os.path.join(HERE, "dropin_check.cpp"), os.path.join(HERE, "cvshim", "opencv2", "opencv.hpp"),
           os.path.join(ROOT, "include", "CudaCostVolumeEnergy.h"), os.path.join(ROOT, "include", "lexp_cuda.h"), os.path.abspath(__file__)]


def _dialect_fixed_guided_filter(text):
    """Add `this->` to the dependent-base names on the right-hand side of the assignments inside
    FastGuidedImageFilter::createSubregionFilter.  Returns (new_text, number_of_lines_changed)."""
    lines = text.split("\n")
    start = next(i for i, l in enumerate(lines) if "class FastGuidedImageFilter" in l)
    stop = next(i for i, l in enumerate(lines) if "class BilateralFilter" in l)
    pat = re.compile(r"^(\s*filter->[\w\[\]]+\s*=\s*)(?=[A-Za-z_])")
    n = 0
    for i in range(start, stop):
        new = pat.sub(r"\1this->", lines[i])
        if new != lines[i]:
            lines[i] = new
            n += 1
    return "\n".join(lines), n


def _dialect_fixed_optimiser(text):
    """MSVC-only constructs of PMStereoBase.h / FastGCStereo.h: default arguments that bind a temporary to a non-const reference
    (`cv::Mat& x = cv::Mat()`, PMStereoBase.h:87, FastGCStereo.h:133) and the qualified spelling `NaiveStereoEnergy::Reusable()`
    (FastGCStereo.h:111,128), both mapped to the shim's `lexp_msvc_default_arg` (a fresh default instance per use)."""
    new = re.sub(r"(cv::Mat\s*&\s*\w+\s*=\s*)cv::Mat\(\)", r"\1lexp_msvc_default_arg()", text)
    new = new.replace("NaiveStereoEnergy::Reusable()", "Reusable()")
    return new, sum(1 for a, b in zip(text.split("\n"), new.split("\n")) if a != b)


def available():
    return os.path.exists(LIB)


def reference_present():
    return os.path.isfile(os.path.join(REF_DIR, "CostVolumeEnergy.h"))


def build(force=False, verbose=False):
    """Returns the path of liblexp_ref.so, or None when it neither exists nor can be built."""
    if not reference_present():
        return LIB if available() else None
    cuda_dir = os.path.join(ROOT, "localexpstereo_b200")
    have_cuda_lib = os.path.exists(os.path.join(cuda_dir, "liblexp_cuda.so"))
    if available() and not force:
        newest = max(os.path.getmtime(p) for p in SOURCES)
        emu_so = os.path.join(ROOT, "tests", "emu", "liblexp_emu.so")
        emu_ok = not os.path.exists(emu_so) or (os.path.exists(DROPIN_EMU) and os.path.getmtime(DROPIN_EMU) >= max(newest, os.path.getmtime(emu_so)))
        if os.path.getmtime(LIB) >= newest and (not have_cuda_lib or (os.path.exists(DROPIN) and os.path.getmtime(DROPIN) >= newest)) and emu_ok:
            return LIB
    gen = os.path.join(OUT_DIR, "gen")
    os.makedirs(gen, exist_ok=True)
    try:
        with open(os.path.join(REF_DIR, "GuidedFilter.h"), "r", encoding="latin-1") as f:
            fixed, n = _dialect_fixed_guided_filter(f.read())
        if n != 17:
            raise RuntimeError(f"dialect fix touched {n} lines of GuidedFilter.h, expected 17: the reference changed")
        with open(os.path.join(gen, "GuidedFilter.h"), "w", encoding="latin-1") as f:
            f.write(fixed)
        for name, expect in (("PMStereoBase.h", 1), ("FastGCStereo.h", 3)):
            with open(os.path.join(REF_DIR, name), "r", encoding="latin-1") as f:
                fixed, n = _dialect_fixed_optimiser(f.read())
            if n != expect:
                raise RuntimeError(f"dialect fix touched {n} lines of {name}, expected {expect}: the reference changed")
            with open(os.path.join(gen, name), "w", encoding="latin-1") as f:
                f.write(fixed)
        cmd = [CXX, "-std=c++14", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fpermissive", "-w",
               "-I", gen, "-I-", "-I", os.path.join(HERE, "cvshim"), "-I", os.path.join(HERE, "refstub"), "-I", REF_DIR,
               os.path.join(HERE, "ref_driver.cpp"), "-o", LIB + ".tmp"]
        if verbose:
            print(" ".join(cmd))
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building oracle/_ref failed:\n" + r.stderr[-6000:])
        os.replace(LIB + ".tmp", LIB)
        if have_cuda_lib:  # the drop-in harness needs the product library to link against (it is run on the GPU box only)
            cmd = [CXX, "-std=c++14", "-O2", "-fopenmp", "-ffp-contract=off", "-fpermissive", "-w",
                   "-I", gen, "-I-", "-I", os.path.join(HERE, "cvshim"), "-I", os.path.join(HERE, "refstub"), "-I", REF_DIR, "-I", os.path.join(ROOT, "include"),
                   os.path.join(HERE, "dropin_check.cpp"), "-o", DROPIN + ".tmp", "-L", cuda_dir, "-llexp_cuda",
                   "-Wl,-rpath,$ORIGIN/../../localexpstereo_b200"]
            if verbose:
                print(" ".join(cmd))
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("building oracle/_ref/dropin_check failed:\n" + r.stderr[-6000:])
            os.replace(DROPIN + ".tmp", DROPIN)
        emu_dir = os.path.join(ROOT, "tests", "emu")
        if os.path.exists(os.path.join(emu_dir, "liblexp_emu.so")):  # CPU twin: the adapter + the kernel source, no GPU
            cmd = [CXX, "-std=c++14", "-O2", "-fopenmp", "-ffp-contract=off", "-fpermissive", "-w",
                   "-I", gen, "-I-", "-I", os.path.join(HERE, "cvshim"), "-I", os.path.join(HERE, "refstub"), "-I", REF_DIR, "-I", os.path.join(ROOT, "include"),
                   os.path.join(HERE, "dropin_check.cpp"), "-o", DROPIN_EMU + ".tmp", "-L", emu_dir, "-llexp_emu",
                   "-Wl,-rpath,$ORIGIN/../../tests/emu"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("building oracle/_ref/dropin_check_emu failed:\n" + r.stderr[-6000:])
            os.replace(DROPIN_EMU + ".tmp", DROPIN_EMU)
    finally:
        shutil.rmtree(gen, ignore_errors=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
