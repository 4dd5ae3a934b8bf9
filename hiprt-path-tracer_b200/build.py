#!/usr/bin/env python3
"""Builds libb2pt.so (sm_100a) in-tree with nvcc. No GPU is needed to build (nvcc cross-compiles).

    python hiprt-path-tracer_b200/build.py [--force] [--jobs N]

The kernels translation unit is compiled per (flavour, variant group, counters) so the objects build in parallel:
  * precise  : -fmad=false, IEEE division / sqrt  -> the flavour parity with the reference CPU path is asserted on
  * fast     : FMA contraction + --use_fast_math   -> the A/B flavour
Feature macros (B2_HAVE_*) mirror which strategy headers exist, so a partially built tree still links.
"""
import argparse
import concurrent.futures
import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
# Experiment builds: B2PT_BUILD_TAG=mb8 B2PT_EXTRA_DEFS="-DB2_MK_MIN_BLOCKS=8" -> libb2pt_mb8.so (load it with B2PT_LIB=...)
TAG = os.environ.get("B2PT_BUILD_TAG", "")
EXTRA_DEFS = os.environ.get("B2PT_EXTRA_DEFS", "").split()
OUT = PKG / (f"libb2pt_{TAG}.so" if TAG else "libb2pt.so")
OBJ = PKG / (f"build_{TAG}" if TAG else "build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC,-fopenmp,-fno-fast-math", f"-I{ROOT / 'include'}", f"-I{CSRC}",
          "-diag-suppress", "821", "-Xptxas", "-v"]
FEATURES = ["-DB2_HAVE_PRINCIPLED", "-DB2_HAVE_RIS", "-DB2_HAVE_RESTIR_DI"]

PRECISE_FLAGS = os.environ.get("B2PT_PRECISE_FLAGS", "-fmad=false -prec-div=true -prec-sqrt=true -ftz=false").split()
FAST_FLAGS = os.environ.get("B2PT_FAST_FLAGS", "--use_fast_math").split()
# b2_kernels.cu is compiled once per (flavour, unit, counters): unit 1 = c1, 2 = c2, 3 = c3, 4 = c4 CameraRays + MegaKernel,
# 5 = c4 ReSTIR DI passes, 6 / 7 / 8 = the c4 MegaKernel with ReSTIR_DI_LaterBouncesSamplingStrategy 0 / 1 / 2. B2PT_ONLY_UNITS="2,3" / B2PT_NO_COUNTERS=1 / B2PT_ONLY_FLAVOUR=precise shorten experiment builds.
ONLY_UNITS = [int(u) for u in os.environ.get("B2PT_ONLY_UNITS", "1,2,3,4,5,6,7,8").split(",")]
ONLY_FLAVOUR = os.environ.get("B2PT_ONLY_FLAVOUR", "")
COUNTER_BUILDS = [0] if os.environ.get("B2PT_NO_COUNTERS") else [0, 1]
UNIT_FEATURE = {1: None, 2: "-DB2_HAVE_PRINCIPLED", 3: "-DB2_HAVE_RIS", 4: "-DB2_HAVE_RESTIR_DI", 5: "-DB2_HAVE_RESTIR_DI",
                6: "-DB2_HAVE_RESTIR_DI", 7: "-DB2_HAVE_RESTIR_DI", 8: "-DB2_HAVE_RESTIR_DI"}


def kernel_units(feats):
    units = []
    for flavour, flags in (("precise", ["-DB2_FLAVOUR_PRECISE", *PRECISE_FLAGS]), ("fast", ["-DB2_FLAVOUR_FAST", *FAST_FLAGS])):
        if ONLY_FLAVOUR and flavour != ONLY_FLAVOUR:
            continue
        for unit in ONLY_UNITS:
            if UNIT_FEATURE[unit] and UNIT_FEATURE[unit] not in feats:
                continue
            for counters in COUNTER_BUILDS:
                units.append(("b2_kernels.cu", f"b2_kernels_{flavour}_u{unit}{'c' if counters else ''}.o", [*flags, f"-DB2_UNIT={unit}", f"-DB2_UNIT_COUNTERS={counters}"]))
    # the longest compiles first
    units.sort(key=lambda u: (-int(u[1].split("_u")[1][0]), u[1]))
    return units


HOST_UNITS = [
    # (source, object name, extra flags)
    ("b2_capi.cu", "b2_capi.o", []),
    ("b2_comm.cu", "b2_comm.o", []),
    ("b2_renderer.cpp", "b2_renderer.o", ["-x", "cu"]),
    ("b2_host_scene.cpp", "b2_host_scene.o", ["-x", "cu"]),
    ("b2_bvh_build.cpp", "b2_bvh_build.o", ["-x", "cu"]),
]


def enabled_features():
    feats = []
    text = {p.name: p.read_text() for p in CSRC.glob("*.h")}
    if "principled_bsdf_eval(const b2pt_render_data& rd, BsdfCtx& ctx, float& pdf)\n{" in text.get("b2_principled.h", ""):
        feats.append("-DB2_HAVE_PRINCIPLED")
    if "placeholder" not in text.get("b2_ris.h", "placeholder"):
        feats.append("-DB2_HAVE_RIS")
    if "placeholder" not in text.get("b2_restir_di.h", "placeholder"):
        feats.append("-DB2_HAVE_RESTIR_DI")
    return feats


def source_digest():
    h = hashlib.sha256()
    h.update(" ".join([*EXTRA_DEFS, *PRECISE_FLAGS, *FAST_FLAGS, *map(str, ONLY_UNITS), ONLY_FLAVOUR, *map(str, COUNTER_BUILDS)]).encode())
    for p in sorted(list(CSRC.glob("*")) + list((ROOT / "include").glob("*.h")) + [Path(__file__), ROOT / "tools" / "traverse_selftest.cu", ROOT / "tools" / "div_selftest.cu"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def compile_unit(unit, feats, log_dir):
    src, obj, extra = unit
    cmd = ["nvcc", *ARCH, *COMMON, *feats, *EXTRA_DEFS, *extra, "-c", str(CSRC / src), "-o", str(OBJ / obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    (log_dir / (obj + ".log")).write_text(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src} ({obj}):\n{res.stdout}\n{res.stderr}")
    return obj


def build(force=False, jobs=None, verbose=True):
    OBJ.mkdir(exist_ok=True)
    digest = source_digest()
    stamp = OUT.with_suffix(".so.digest")  # next to the library: build/ (objects, logs) does not travel to the GPU box (.gpurunignore)
    if not force and OUT.exists() and stamp.exists() and stamp.read_text() == digest:
        if verbose:
            print(f"{OUT} is up to date")
        return OUT
    feats = enabled_features()
    UNITS = kernel_units(feats) + HOST_UNITS
    with concurrent.futures.ThreadPoolExecutor(max_workers=jobs or min(8, os.cpu_count() or 1)) as pool:
        list(pool.map(lambda u: compile_unit(u, feats, OBJ), UNITS))
    link = ["nvcc", *ARCH, "-shared", "-Xcompiler", "-fopenmp", "-o", str(OUT), *[str(OBJ / u[1]) for u in UNITS], "-lgomp", "-ldl"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    selftest = ["nvcc", *ARCH, "-O3", "-std=c++17", "-fmad=false", "-lineinfo", f"-I{ROOT / 'include'}", f"-I{CSRC}", "-diag-suppress", "821",
                str(ROOT / "tools" / "traverse_selftest.cu"), str(CSRC / "b2_bvh_build.cpp"), "-Xcompiler", "-fopenmp", "-o", str(ROOT / "tools" / "traverse_selftest")]
    res = subprocess.run(selftest, capture_output=True, text=True) if not TAG else subprocess.CompletedProcess(selftest, 0)
    if res.returncode != 0:
        raise RuntimeError(f"traverse_selftest build failed:\n{res.stdout}\n{res.stderr}")
    divtest = ["nvcc", *ARCH, "-O3", "-std=c++17", *PRECISE_FLAGS, "-lineinfo", f"-I{ROOT / 'include'}", f"-I{CSRC}", "-diag-suppress", "821",
               str(ROOT / "tools" / "div_selftest.cu"), "-o", str(ROOT / "tools" / "div_selftest")]
    res = subprocess.run(divtest, capture_output=True, text=True) if not TAG else subprocess.CompletedProcess(divtest, 0)
    if res.returncode != 0:
        raise RuntimeError(f"div_selftest build failed:\n{res.stdout}\n{res.stderr}")
    stamp.write_text(digest)
    if verbose:
        print(f"built {OUT} with features {feats or ['(c1 only)']}")
    return OUT


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--jobs", type=int, default=None)
    a = ap.parse_args()
    try:
        build(a.force, a.jobs)
    except RuntimeError as e:
        print(e, file=sys.stderr)
        sys.exit(1)
