"""The reference's own GPUKernel class bound to libb2pt.so: hiprt-path-tracer_b200/binding/GPUKernel_b2pt.cpp
implements every member of src/Compiler/GPUKernel.h (the reference's header, unmodified) on include/b2pt.h and is linked with the
reference's real GPUKernelCompilerOptions.cpp (oracle/Makefile).
 * CPU (oracle/_ref/binding_check): every macro name the library understands is one the reference registers, every macro served at one
   value only has that value as the reference's default, and ~190 combinations of GPUKernelCompilerOptions (the BASELINE configs, every
   bias-correction weighting, later-bounces strategy, visibility macro, presampling) resolve to a precompiled variant -- or are refused
   where nothing is built -- exactly as GPUKernel::compile would pass them (b2pt_kernel_query: the resolution of b2pt_kernel_get
   without a device).
 * GPU (oracle/_ref/boundary_gpukernel_<variant>): the boundary proof of tests/test_gpu_boundary.py with every kernel held, compiled and
   launched as a reference GPUKernel object. Added after the round's GPU budget was spent: first run by the driver."""
import ctypes as C
import json
import subprocess

import pytest

from conftest import DATA, REF_DIR, ROOT


def test_binding_check_on_the_host(built):
    exe = REF_DIR / "binding_check"
    if not exe.exists():
        pytest.skip(f"{exe} not built (needs /root/reference at build time)")
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120, cwd=ROOT)
    lines = [l for l in out.stdout.splitlines() if l.startswith('{"binding_check"')]
    assert lines, out.stdout[-1500:] + out.stderr[-1500:]
    report = json.loads(lines[-1])
    assert out.returncode == 0 and report["failures"] == 0, out.stderr[-3000:]
    assert report["fixed"] + report["switched"] == report["library_macros"]
    assert report["combinations_walked"] > 150


def test_kernel_query_needs_no_device(b2pt):
    """b2pt_kernel_query: same answers as b2pt_kernel_get, no context (the library's own view of the option surface, through ctypes)."""
    lib = b2pt.load_library()

    def query(entry, **options):
        arr = (b2pt.Option * len(options))(*[b2pt.Option(k.encode(), v) for k, v in options.items()])
        return lib.b2pt_kernel_query(entry.encode(), arr, len(options))

    assert query("MegaKernel") == 0
    assert query("MegaKernel", BSDFOverride=1, DirectLightSamplingStrategy=1, DirectLightSamplingBaseStrategy=0) == 0
    assert query("ReSTIR_DI_SpatiotemporalReuse", DirectLightSamplingStrategy=5, ReSTIR_DI_BiasCorrectionWeights=3) == 0
    assert query("ReSTIR_DI_SpatiotemporalReuse", DirectLightSamplingStrategy=5, ReSTIR_DI_BiasCorrectionWeights=0) == 0
    assert query("ReSTIR_DI_SpatialReuse", DirectLightSamplingStrategy=5, ReSTIR_DI_BiasCorrectionWeights=0) != 0  # 1/M: the reference's does not compile
    assert query("MegaKernel", DirectLightSamplingStrategy=5, ReSTIR_DI_LaterBouncesSamplingStrategy=1) == 0
    assert query("MegaKernel", NoSuchMacro=1) != 0
    assert query("NoSuchKernel") != 0
    names = C.POINTER(C.c_char_p)()
    count = C.c_int(0)
    assert lib.b2pt_kernel_option_names(C.byref(names), C.byref(count)) == 0
    listed = {names[i].decode() for i in range(count.value)}
    assert {"DirectLightSamplingStrategy", "ReSTIR_DI_BiasCorrectionWeights", "KernelWorkgroupThreadCount"} <= listed
    value = C.c_int(-1)
    assert lib.b2pt_kernel_option_fixed_value(b"NestedDielectricsStackSize", C.byref(value)) == 0 and value.value == 4
    assert lib.b2pt_kernel_option_fixed_value(b"DirectLightSamplingStrategy", C.byref(value)) != 0


def test_every_named_variant_resolves_without_a_device(b2pt):
    """The renderer's variant table (c1 ... c4x2): every pass of every variant must name a precompiled kernel, in both flavours, with and
    without traversal counters -- walked on the host (b2pt_renderer_variant_query), so a slip in the table shows here and not on a GPU."""
    lib = b2pt.load_library()
    lib.b2pt_renderer_variant_name.restype = C.c_char_p
    names = [lib.b2pt_renderer_variant_name(i).decode() for i in range(lib.b2pt_renderer_variant_count())]
    assert {"c1", "c2", "c3", "c4", "c4p", "c4pu", "c4b1", "c4b7", "c4v5", "c4l0", "c4x2"} <= set(names)
    for name in names:
        for precise in (1, 0):
            assert lib.b2pt_renderer_variant_query(name.encode(), precise) == 0, (name, precise, lib.b2pt_last_error_string())
    assert lib.b2pt_renderer_variant_query(b"c9", 1) != 0


@pytest.mark.gpu
@pytest.mark.parametrize("variant,scene,extra", [("c1", "cornell_pbr_16x9", []), ("c2", "cornell_pbr_16x9", []), ("c3", "synthetic_20k", []),
                                                  ("c4", "cornell_pbr_16x9", []), ("c4", "cornell_pbr_16x9", ["--no-fused"])])
def test_reference_gpukernel_objects_drive_the_device_layer(built, variant, scene, extra):
    exe = REF_DIR / f"boundary_gpukernel_{variant}"
    if not exe.exists():
        pytest.skip(f"{exe} not built (needs /root/reference at build time)")
    out = subprocess.run([str(exe), "--scene", str(DATA / f"{scene}.b2scene"), "--luts", str(DATA / "brdf_luts.b2luts"), "--width", "256", "--height", "144",
                          "--spp", "3", "--bounces", "4", *extra], capture_output=True, text=True, timeout=300, cwd=ROOT)
    lines = [l for l in out.stdout.splitlines() if l.startswith('{"boundary_check"')]
    assert lines, out.stdout[-1500:] + out.stderr[-1500:]
    assert any(l.startswith('{"through_gpukernel_class"') for l in out.stdout.splitlines())
    report = json.loads(lines[-1])
    assert report["pixels_not_bit_identical"] <= (2 if variant == "c4" else 0), report
    assert out.returncode == 0 or variant == "c4"
    assert report["mean"] > 0
