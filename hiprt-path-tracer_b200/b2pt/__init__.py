"""ctypes binding of libb2pt.so (C ABI: include/b2pt.h) -- the host-side mirror used by tests/ and bench.py.

`Renderer` mirrors the reference's renderer front end (set_scene / set_envmap / render / get_framebuffer,
/root/reference/src/Renderer/GPURenderer.h:61-346, CPURenderer.h:33-109). There is no CPU fallback: if the CUDA library
is missing or no device is present, construction raises.
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

from . import partition

PKG_DIR = Path(__file__).resolve().parent.parent
LIB_PATH = PKG_DIR / "libb2pt.so"
REPO_ROOT = PKG_DIR.parent

# index = the B2_PASS_* enumerator of csrc/b2_frame_loop.h (tests/test_abi.py keeps the two in step)
PASS_NAMES = ["CameraRays", "ReSTIR_DI_LightsPresampling", "ReSTIR_Directional_Reuse_Compute", "ReSTIR_DI_InitialCandidates",
              "ReSTIR_DI_TemporalReuse", "ReSTIR_DI_SpatialReuse", "ReSTIR_DI_SpatiotemporalReuse", "MegaKernel"]


class B2ptError(RuntimeError):
    pass


class DeviceProps(C.Structure):
    _fields_ = [("name", C.c_char * 256), ("sm_count", C.c_int), ("compute_major", C.c_int), ("compute_minor", C.c_int),
                ("total_memory_bytes", C.c_size_t), ("l2_bytes", C.c_size_t), ("shared_memory_per_block_optin", C.c_size_t),
                ("clock_khz", C.c_int)]


class BvhStats(C.Structure):
    _fields_ = [("node_count", C.c_uint64), ("triangle_count", C.c_uint64), ("bytes", C.c_uint64), ("max_depth", C.c_int),
                ("sah_cost", C.c_double), ("build_seconds", C.c_double)]


class Counters(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("nodes", C.c_uint64), ("triangles", C.c_uint64)]


class Faults(C.Structure):
    _fields_ = [("traversal_guard_trips", C.c_uint32), ("boundary_guard_trips", C.c_uint32), ("stack_overflows", C.c_uint32)]


class Option(C.Structure):
    _fields_ = [("name", C.c_char_p), ("value", C.c_int)]


# Byte offsets inside b2pt_render_data of the settings tests and benches edit (checked against the reference's
# HIPRTRenderData by oracle/layout_check.cpp through include/b2pt_render_data.h; see tests/test_layout.py).
_RS = 728
SETTINGS_OFFSETS = {
    "nb_bounces": (_RS + 68, "i"),
    "samples_per_frame": (_RS + 64, "i"),
    "sample_number": (_RS + 56, "I"),
    "need_to_reset": (_RS + 8, "B"),
    "do_russian_roulette": (_RS + 72, "B"),
    "russian_roulette_min_depth": (_RS + 76, "i"),
    "freeze_random": (_RS + 88, "i"),
    "accumulate": (_RS + 49, "B"),
    "enable_direct": (_RS + 28, "B"),
    "ris_number_of_light_candidates": (_RS + 152, "i"),
    "ris_number_of_bsdf_candidates": (_RS + 156, "i"),
    "restir_di_do_temporal_reuse_pass": (_RS + 160 + 0, "B"),
    "restir_di_do_spatial_reuse_pass": (_RS + 160 + 24 + 0, "B"),
    "restir_di_number_of_passes": (_RS + 160 + 24 + 8, "i"),
    "restir_di_reuse_radius": (_RS + 160 + 24 + 12, "i"),
    "restir_di_reuse_neighbor_count": (_RS + 160 + 24 + 20, "i"),
    "restir_di_disocclusion_reuse_count": (_RS + 160 + 24 + 28, "i"),
    "restir_di_use_adaptive_directional_spatial_reuse": (_RS + 160 + 24 + 49, "B"),
    "restir_di_m_cap": (_RS + 160 + 180, "i"),
    "restir_di_use_confidence_weights": (_RS + 160 + 184, "B"),
    "restir_di_do_fused_spatiotemporal": (_RS + 160 + 272, "B"),
    "current_camera_do_jittering": (2328 + 216, "B"),
    "ambient_light_type": (2088 + 0, "i"),
    "envmap_intensity": (2088 + 24, "f"),
    "white_furnace_mode": (464 + 0, "B"),
}
FRAMEBUFFER_POINTER_OFFSET = None  # filled below from the C side (offsetof(b2pt_render_data, buffers.accumulated_ray_colors))
_PTR = C.c_void_p
_CT = {"i": C.c_int, "I": C.c_uint, "B": C.c_uint8, "f": C.c_float}


def load_library():
    global LIB_PATH
    if os.environ.get("B2PT_LIB"):  # experiment builds (build.py: B2PT_BUILD_TAG)
        LIB_PATH = Path(os.environ["B2PT_LIB"]).resolve()
    if not LIB_PATH.exists():
        raise B2ptError(f"{LIB_PATH} is missing: build it with `python hiprt-path-tracer_b200/build.py` "
                        "(there is no CPU fallback for the B200 path)")
    lib = C.CDLL(str(LIB_PATH))
    lib.b2pt_last_error_string.restype = C.c_char_p
    lib.b2pt_version_string.restype = C.c_char_p
    lib.b2pt_renderer_render_data.restype = C.c_void_p
    lib.b2pt_bvh_device_handle.restype = C.c_void_p
    global FRAMEBUFFER_POINTER_OFFSET
    FRAMEBUFFER_POINTER_OFFSET = lib.b2pt_render_data_framebuffer_offset()
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = load_library()
    return _lib


def check(rc):
    if rc != 0:
        raise B2ptError(f"b2pt error {rc}: {lib().b2pt_last_error_string().decode()}")


def declared_symbols():
    """Names of every function include/b2pt.h declares (used by the CPU-side ABI test)."""
    import re
    text = (REPO_ROOT / "include" / "b2pt.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2pt_[a-z0-9_]+)\s*\(", text)))


class Context:
    def __init__(self, device=0):
        self.handle = C.c_void_p()
        check(lib().b2pt_ctx_create(C.c_int(device), C.byref(self.handle)))

    def props(self):
        p = DeviceProps()
        check(lib().b2pt_ctx_device_props(self.handle, C.byref(p)))
        return p

    def counters_reset(self):
        check(lib().b2pt_counters_reset(self.handle))

    def counters_read(self):
        c = Counters()
        check(lib().b2pt_counters_read(self.handle, C.byref(c)))
        return dict(rays=c.rays, nodes=c.nodes, triangles=c.triangles)

    def cuda_launches(self):
        n = C.c_uint64()
        check(lib().b2pt_cuda_launch_count(self.handle, C.byref(n)))
        return n.value

    WF_KINDS = ("trace", "stage1", "stage2", "sort", "begin", "advance", "pass_a", "pass_b", "pass_c", "trace_pass")  # B2PT_WF_KIND_NAMES

    def wavefront_profile(self, enabled):
        """Per-kernel CUDA-event timing of the wavefront launch sequences on / off (resets the totals)."""
        check(lib().b2pt_wavefront_profile(self.handle, int(enabled)))

    def wavefront_profile_read(self):
        """{kind: dict(ms=..., groups=...)} accumulated since wavefront_profile(True)."""
        n = len(self.WF_KINDS)
        ms = (C.c_double * n)()
        groups = (C.c_uint64 * n)()
        check(lib().b2pt_wavefront_profile_read(self.handle, ms, groups, n))
        return {k: dict(ms=ms[i], groups=int(groups[i])) for i, k in enumerate(self.WF_KINDS)}

    def faults(self, reset=True):
        f = Faults()
        check(lib().b2pt_faults_read(self.handle, C.byref(f), int(reset)))
        return dict(traversal_guard_trips=f.traversal_guard_trips, boundary_guard_trips=f.boundary_guard_trips, stack_overflows=f.stack_overflows)

    def close(self):
        if self.handle:
            lib().b2pt_ctx_destroy(self.handle)
            self.handle = C.c_void_p()


class Communicator:
    """NCCL communicator of the C layer (b2pt_comm): rank 0 draws the unique id, torch.distributed broadcasts it."""

    def __init__(self, ctx, rank, world, torch, dist):
        ident = (C.c_ubyte * 128)()
        if rank == 0:
            check(lib().b2pt_comm_unique_id(ident))
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor(list(ident), dtype=torch.uint8, device=device)
        dist.broadcast(t, src=0)
        ident = (C.c_ubyte * 128)(*t.cpu().tolist())
        self.handle = C.c_void_p()
        check(lib().b2pt_comm_create(ctx.handle, ident, rank, world, C.byref(self.handle)))
        self.rank, self.world = rank, world

    def stats(self):
        e, g, b = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(lib().b2pt_comm_stats(self.handle, C.byref(e), C.byref(g), C.byref(b)))
        return dict(exchanges=e.value, gathers=g.value, bytes_sent=b.value)

    def close(self):
        if self.handle:
            lib().b2pt_comm_destroy(self.handle)
            self.handle = C.c_void_p()


def read_hdr(path, flip_y=True):
    """b2pt_image_read_hdr: the library's Radiance .hdr decoder (host only) -> (height, width, 4) float32."""
    w, h = C.c_int(), C.c_int()
    check(lib().b2pt_image_read_hdr(str(path).encode(), int(flip_y), None, C.byref(w), C.byref(h)))
    out = np.empty((h.value, w.value, 4), np.float32)
    check(lib().b2pt_image_read_hdr(str(path).encode(), int(flip_y), out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(w), C.byref(h)))
    return out


def partition_rows(height, rank, world, band_rows, halo=0):
    """b2pt_partition_rows (host logic of the C layer): rows within `halo` of rank's bands."""
    rows = (C.c_int * height)()
    n = C.c_int()
    check(lib().b2pt_partition_rows(height, rank, world, band_rows, halo, rows, C.byref(n)))
    return np.array(rows[:n.value], dtype=np.int64)


class Renderer:
    """Headless renderer: one instance = one (resolution, variant, flavour) on one GPU."""

    def __init__(self, ctx, width, height, variant="c1", precise=True):
        self.ctx, self.width, self.height, self.variant = ctx, width, height, variant
        self.handle = C.c_void_p()
        check(lib().b2pt_renderer_create(ctx.handle, width, height, variant.encode(), int(bool(precise)), C.byref(self.handle)))
        self._rd = lib().b2pt_renderer_render_data(self.handle)

    # -- scene ------------------------------------------------------------------------------------------------
    def set_scene_file(self, path):
        check(lib().b2pt_renderer_set_scene_file(self.handle, str(path).encode()))

    def set_scene(self, scene):
        """scene: dict as returned by tools/b2scene.read_scene (host numpy arrays)."""
        f32 = lambda a: np.ascontiguousarray(a, np.float32)
        pos, nrm, uv = f32(scene["positions"]), f32(scene["normals"]), f32(scene["texcoords"])
        has = np.ascontiguousarray(scene["has_normals"], np.uint8)
        idx = np.ascontiguousarray(scene["indices"], np.int32)
        mid = np.ascontiguousarray(scene["material_indices"], np.int32)
        mats = f32(scene["materials"])
        cam = scene["camera"]
        cam52 = np.concatenate([f32(cam["inverse_view"]).ravel(), f32(cam["inverse_projection"]).ravel(), f32(cam["view_projection"]).ravel(),
                                f32(cam["position"]).ravel(), np.array([cam["vertical_fov"]], np.float32)])
        p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        check(lib().b2pt_renderer_set_scene_arrays(self.handle, p(pos, C.c_float), p(nrm, C.c_float), p(uv, C.c_float), p(has, C.c_uint8),
                                                   len(has), p(idx, C.c_int), p(mid, C.c_int), len(mid), p(mats, C.c_float), len(mats),
                                                   p(cam52, C.c_float)))

    def reupload_scene(self):
        """Async host->device re-send of all scene arrays + BVH blob on the renderer's stream; returns bytes copied."""
        n = C.c_uint64()
        check(lib().b2pt_renderer_reupload_scene(self.handle, C.byref(n)))
        self._scene_upload_bytes = n.value
        return n.value

    def scene_upload_bytes(self):
        return getattr(self, "_scene_upload_bytes", 0)

    def set_luts_file(self, path):
        check(lib().b2pt_renderer_set_luts_file(self.handle, str(path).encode()))

    def set_envmap_file(self, path):
        check(lib().b2pt_renderer_set_envmap_file(self.handle, str(path).encode()))

    def set_envmap_hdr_file(self, path):
        check(lib().b2pt_renderer_set_envmap_hdr_file(self.handle, str(path).encode()))

    def set_envmap_matrices(self, envmap_to_world, world_to_envmap):
        a = np.ascontiguousarray(envmap_to_world, np.float32).reshape(9)
        b = np.ascontiguousarray(world_to_envmap, np.float32).reshape(9)
        check(lib().b2pt_renderer_set_envmap_matrices(self.handle, a.ctypes.data_as(C.POINTER(C.c_float)), b.ctypes.data_as(C.POINTER(C.c_float))))

    def update_materials(self, materials):
        m = np.ascontiguousarray(materials, np.float32)
        check(lib().b2pt_renderer_update_materials(self.handle, m.ctypes.data_as(C.POINTER(C.c_float)), len(m)))

    def set_envmap(self, rgba):
        if rgba is None:
            check(lib().b2pt_renderer_set_envmap(self.handle, None, 0, 0))
            return
        rgba = np.ascontiguousarray(rgba, np.float32)
        h, w, c = rgba.shape
        assert c == 4
        check(lib().b2pt_renderer_set_envmap(self.handle, rgba.ctypes.data_as(C.POINTER(C.c_float)), w, h))

    # -- settings ---------------------------------------------------------------------------------------------
    def set(self, **kw):
        for k, v in kw.items():
            off, t = SETTINGS_OFFSETS[k]
            _CT[t].from_address(self._rd + off).value = v

    def get(self, name):
        off, t = SETTINGS_OFFSETS[name]
        return _CT[t].from_address(self._rd + off).value

    def set_host_seed(self, seed):
        check(lib().b2pt_renderer_set_host_seed(self.handle, C.c_uint(seed)))

    def set_counters(self, enabled):
        check(lib().b2pt_renderer_set_counters(self.handle, int(enabled)))

    def set_macro(self, name, value):
        """GPUKernelCompilerOptions::set_macro_value + recompile: a kernel option on top of the variant's (b2pt_renderer_set_macro)."""
        check(lib().b2pt_renderer_set_macro(self.handle, name.encode(), int(value)))

    def set_profile_passes(self, enabled):
        check(lib().b2pt_renderer_set_profile_passes(self.handle, int(enabled)))

    def reset(self):
        check(lib().b2pt_renderer_reset(self.handle))

    # -- multi-GPU --------------------------------------------------------------------------------------------
    def set_partition(self, rank, world, torch=None, dist=None, dynamic_camera=False):
        """This renderer renders only rank's interleaved row bands. The exchanges (ReSTIR DI reservoir halos, framebuffer
        all-gather) are NCCL calls issued from C on the renderer's stream (csrc/b2_comm.cu); torch.distributed only carries the
        128-byte ncclUniqueId from rank 0 to the other ranks once."""
        self.rank, self.world = rank, world
        if world > 1 and getattr(self, "_comm", None) is None:
            if torch is None or dist is None:
                raise B2ptError("a partition over several GPUs needs torch + torch.distributed to bootstrap the NCCL communicator")
            self._comm = Communicator(self.ctx, rank, world, torch, dist)
        comm = self._comm.handle if world > 1 else None
        check(lib().b2pt_renderer_set_partition(self.handle, rank, world, comm, int(bool(dynamic_camera))))

    @property
    def exchange_count(self):
        return self._comm.stats()["exchanges"] if getattr(self, "_comm", None) else 0

    def partition_info(self):
        cam, own, band, halo = C.c_uint64(), C.c_uint64(), C.c_int(), C.c_int()
        check(lib().b2pt_renderer_partition_info(self.handle, C.byref(cam), C.byref(own), C.byref(band), C.byref(halo)))
        return dict(camera_pixels=cam.value, owned_pixels=own.value, band_rows=band.value, halo_rows=halo.value)

    def owned_pixels(self):
        return self.partition_info()["owned_pixels"]

    def camera_rays_pixels(self):
        """Pixels CameraRays covers on this rank (ReSTIR DI partitions also fill the halo rows its reuse passes read)."""
        return self.partition_info()["camera_pixels"]

    def partition_description(self):
        if getattr(self, "world", 1) <= 1:
            return "single GPU"
        info = self.partition_info()
        halo = f", ReSTIR DI halo {info['halo_rows']} rows (CameraRays recomputed, reservoirs by NCCL send/recv)" if info["halo_rows"] else ""
        return f"interleaved {info['band_rows']}-row bands, scene replicated{halo}, framebuffer ncclAllGather per step, all on the renderer's stream"

    def framebuffer_device_pointer(self):
        return _PTR.from_address(self._rd + FRAMEBUFFER_POINTER_OFFSET).value

    def gather_framebuffer(self):
        check(lib().b2pt_renderer_gather_framebuffer(self.handle))

    def make_gather(self, torch=None, dist=None):
        """Returns gather(): all-gathers the framebuffer bands so every rank holds the whole frame (in place). The collective is
        issued from C on the renderer's stream and counted by last_render_ms(), hence the 0.0 extra milliseconds."""
        def gather():
            self.gather_framebuffer()
            return 0.0

        return gather

    def set_camera(self, camera52):
        cam = np.ascontiguousarray(camera52, np.float32)
        assert cam.size == 52
        check(lib().b2pt_renderer_set_camera(self.handle, cam.ctypes.data_as(C.POINTER(C.c_float))))

    # -- rendering --------------------------------------------------------------------------------------------
    def render(self, samples=1, sync=True):
        check(lib().b2pt_renderer_render(self.handle, int(samples)))
        if sync:
            self.synchronize()

    def synchronize(self):
        check(lib().b2pt_renderer_synchronize(self.handle))

    def last_render_ms(self):
        ms = C.c_float()
        check(lib().b2pt_renderer_last_render_ms(self.handle, C.byref(ms)))
        return ms.value

    def pass_times(self):
        out = {}
        for i, name in enumerate(PASS_NAMES):
            ms, n = C.c_float(), C.c_uint64()
            check(lib().b2pt_renderer_pass_ms(self.handle, i, C.byref(ms), C.byref(n)))
            if n.value:
                out[name] = dict(ms=ms.value, launches=n.value)
        return out

    def framebuffer(self, out=None):
        if out is None:
            out = np.empty((self.height, self.width, 3), np.float32)
        check(lib().b2pt_renderer_download_framebuffer(self.handle, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def gbuffer(self):
        n = self.width * self.height
        prim = np.empty(n, np.int32)
        pos = np.empty((n, 3), np.float32)
        geo, shade = np.empty(n, np.uint32), np.empty(n, np.uint32)
        mat = np.empty((n, 27), np.uint32)
        check(lib().b2pt_renderer_download_gbuffer(self.handle, prim.ctypes.data_as(C.POINTER(C.c_int)), pos.ctypes.data_as(C.POINTER(C.c_float)),
                                                   geo.ctypes.data_as(C.POINTER(C.c_uint32)), shade.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                   mat.ctypes.data_as(C.POINTER(C.c_uint32))))
        return dict(prim=prim, position=pos, geometric_normal=geo, shading_normal=shade, material=mat)

    def reservoirs(self, which=3):
        """which: 0 initial, 1 spatial_1, 2 spatial_2, 3 current ReSTIR output. Returns (9, H*W) uint32 planes."""
        out = np.empty((9, self.width * self.height), np.uint32)
        check(lib().b2pt_renderer_download_reservoirs(self.handle, which, out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out

    def scene_info(self):
        st, nt, ne = BvhStats(), C.c_int(), C.c_int()
        check(lib().b2pt_renderer_scene_info(self.handle, C.byref(st), C.byref(nt), C.byref(ne)))
        return dict(triangles=nt.value, emissive_triangles=ne.value, bvh_nodes=st.node_count, bvh_bytes=st.bytes,
                    bvh_depth=st.max_depth, bvh_sah=st.sah_cost, bvh_build_seconds=st.build_seconds)

    def derived(self, n_materials):
        info = self.scene_info()
        packed = np.empty((n_materials, 33), np.uint32)
        areas = np.empty(info["triangles"], np.float32)
        emissive = np.empty(info["emissive_triangles"], np.int32)
        probas = np.empty(info["emissive_triangles"], np.float32)
        alias = np.empty(info["emissive_triangles"], np.int32)
        check(lib().b2pt_renderer_download_derived(self.handle, packed.ctypes.data_as(C.POINTER(C.c_uint32)), areas.ctypes.data_as(C.POINTER(C.c_float)),
                                                   emissive.ctypes.data_as(C.POINTER(C.c_int)), probas.ctypes.data_as(C.POINTER(C.c_float)),
                                                   alias.ctypes.data_as(C.POINTER(C.c_int))))
        return dict(packed_materials=packed, triangle_areas=areas, emissive_triangles=emissive, alias_probas=probas, alias_alias=alias)

    def close(self):
        if self.handle:
            lib().b2pt_renderer_destroy(self.handle)
            self.handle = C.c_void_p()
        if getattr(self, "_comm", None) is not None:
            self._comm.close()
            self._comm = None
