#!/usr/bin/env python3
"""bench.py -- Mpaths/s of the per-pixel path-tracing hot path on N B200s (driver contract: one JSON line on rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the whole hot path (CameraRays -> [ReSTIR DI passes] -> MegaKernel, `spp` samples per pixel)
over one full frame of the workload. The default workload is BASELINE.json configs[1]: Cornell box at 1920x1080,
principled BSDF, 16 spp, NEE + BSDF MIS with power-proportional light sampling (variant "c2"), 8 bounces.

 value      : W*H*spp*steps / device time (CUDA events on the renderer's stream, max over ranks), scene resident in HBM
 e2e        : same metric through the C ABI with HOST buffers: per step the scene arrays are re-uploaded from pinned
              host memory (geometry, materials, BVH blob: what GPURenderer::set_scene hands to the device layer), the
              frame is rendered, and the fp32 framebuffer is read back to the host -- all inside the timed region
 roofline   : MegaKernel's algorithmic bytes (B_path from measured traversal steps) / its CUDA-event time
 cpu_baseline: the reference's own CPU path (oracle/_ref, compiled from /root/reference) on this box's host cores, on
              a bounded pixel window of the same workload

Multi-GPU (N > 1): the frame is partitioned into interleaved 16-row pixel bands (rank = band % N), scene replicated,
no data-path collective inside a sample (ReSTIR DI variants all-gather the reservoir planes a reuse pass reads, once
per frame); the framebuffer bands are all-gathered with NCCL at the end of every step (inside the timed region: the
step's device time = the renderer's CUDA events + the all-gather's). Total work is fixed => "scaling": "strong".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402

METRIC = "Mpaths/s (pixels*spp/s) at 1080p"
WORKLOADS = {
    # name: variant, scene blob, width, height, spp per step, bounces, description
    "c1": dict(variant="c1", scene="cornell_pbr_1x1", width=512, height=512, spp=1, bounces=3,
               desc="Cornell 512x512 1spp 3 bounces Lambertian + uniform one-light NEE (BASELINE configs[0])"),
    "c2": dict(variant="c2", scene="cornell_pbr_16x9", width=1920, height=1080, spp=16, bounces=8,
               desc="Cornell 1080p principled BSDF 16spp NEE+BSDF MIS power-proportional light sampling (BASELINE configs[1])"),
    "c3": dict(variant="c3", scene="cornell_pbr_16x9", width=1920, height=1080, spp=16, bounces=8,
               desc="Cornell 1080p principled BSDF 16spp RIS 4+1 power (shipped defaults)"),
    "c4": dict(variant="c4", scene="cornell_pbr_16x9", width=1920, height=1080, spp=16, bounces=8, fused=1,
               desc="Cornell 1080p principled BSDF ReSTIR DI fused spatiotemporal, 16 frames x 1spp"),
    # the synthetic many-light scene (tools/make_synthetic_scene.py, seed 0xB200: 10,000,000 triangles, 100,000 emissive) is
    # generated on first use into gpurun_out/scenes/ (325 MB; deterministic), like the procedural HDR sky (tools/make_envmap.py)
    "c3s": dict(variant="c3", scene="synthetic:10000000:100000", width=1920, height=1080, spp=4, bounces=8,
                desc="synthetic 10M-triangle scene 1080p 8 bounces RIS 4+1 over 100k emissive triangles (BASELINE configs[2]), 4 spp per step"),
    "c4s": dict(variant="c4", scene="synthetic:10000000:100000", envmap="sky:2048:1024", width=1920, height=1080, spp=4, bounces=8, fused=1,
                desc="synthetic 10M-triangle scene 1080p ReSTIR DI fused spatiotemporal + HDR envmap alias table (BASELINE configs[3]), 4 frames x 1spp per step"),
    "c5": dict(variant="c4", scene="synthetic:10000000:100000", envmap="sky:2048:1024", width=3840, height=2160, spp=4, bounces=8, fused=1,
               desc="synthetic 10M-triangle scene 4K ReSTIR DI fused + envmap, tile-partitioned across GPUs (BASELINE configs[4]), 4 frames x 1spp per step"),
}
SCENE_CACHE = ROOT / "gpurun_out" / "scenes"


def scene_file(w):
    """Path of the workload's B2SCENE1 blob; synthetic scenes are generated (seeded, deterministic) on first use."""
    name = w["scene"]
    if not name.startswith("synthetic:"):
        return ROOT / "data" / f"{name}.b2scene"
    _, tris, emissive = name.split(":")
    path = SCENE_CACHE / f"synthetic_{tris}_{emissive}_{w['width']}x{w['height']}.b2scene"
    if not path.exists():
        SCENE_CACHE.mkdir(parents=True, exist_ok=True)
        tmp = path.with_suffix(f".tmp{os.getpid()}")
        subprocess.run([sys.executable, str(ROOT / "tools" / "make_synthetic_scene.py"), str(tmp), "--triangles", tris, "--emissive", emissive,
                        "--aspect", repr(w["width"] / w["height"])], check=True, stdout=subprocess.DEVNULL)
        os.replace(tmp, path)
    return path


def envmap_file(w):
    name = w.get("envmap")
    if not name:
        return None
    _, ew, eh = name.split(":")
    path = SCENE_CACHE / f"sky_{ew}x{eh}.b2env"
    if not path.exists():
        SCENE_CACHE.mkdir(parents=True, exist_ok=True)
        tmp = path.with_suffix(f".tmp{os.getpid()}")
        subprocess.run([sys.executable, str(ROOT / "tools" / "make_envmap.py"), str(tmp), "--width", ew, "--height", eh], check=True)
        os.replace(tmp, path)
    return path


def oracle_args(w):
    extra = ["--envmap", str(envmap_file(w))] if w.get("envmap") else []
    if w["variant"] == "c4" and not w.get("fused", 0):
        extra.append("--no-fused")
    return extra


# bytes per traversal step and per shaded vertex / pixel
S_NODE, S_TRI, S_HIT, S_LIGHT, S_PX = 80, 48, 164, 76, 288


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.device_index = device_index
        self.proc = None
        self.lines = []
        self.nvml_samples = []  # (sm MHz, max sm MHz, W, reasons bitmask) every 5 ms through NVML: short timed regions (N = 8:
        self.nvml_stop = None   # ~60 ms) end before nvidia-smi's first line arrives
        self.nvml_thread = None

    def _nvml_pump(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.device_index)
            sm_max = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.nvml_stop.is_set():
                self.nvml_samples.append((float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), float(sm_max),
                                          pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))))
                self.nvml_stop.wait(0.005)
        except Exception:  # no NVML / no permission: the nvidia-smi lines remain
            pass

    def start(self):
        try:
            self.nvml_stop = threading.Event()
            self.nvml_thread = threading.Thread(target=self._nvml_pump, daemon=True)
            self.nvml_thread.start()
        except Exception:
            self.nvml_thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device_index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def _nvml_summary(self):
        if not self.nvml_samples:
            return None
        bits = dict(hw_slowdown=0x8, sw_thermal_slowdown=0x20, hw_thermal_slowdown=0x40, sw_power_cap=0x4)  # nvmlClocksEventReason*
        seen = 0
        for s_ in self.nvml_samples:
            seen |= s_[3]
        return dict(sm_mhz=float(np.median([s_[0] for s_ in self.nvml_samples])), sm_max_mhz=float(max(s_[1] for s_ in self.nvml_samples)),
                    power_w_max=float(max(s_[2] for s_ in self.nvml_samples)), samples=len(self.nvml_samples), source="nvml, 5 ms period",
                    reasons=sorted(n for n, b in bits.items() if seen & b))

    def stop(self):
        if self.nvml_thread is not None:
            self.nvml_stop.set()
            self.nvml_thread.join(timeout=2)
        if not self.proc:
            return self._nvml_summary() or dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, sm_max, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                sm_max.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for n, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return self._nvml_summary() or dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(sm_max)), power_w_max=float(max(power)), samples=len(sm), reasons=sorted(reasons))


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except (KeyError, ValueError):
            pass
    return 6650.0, "fallback"


def pick_workload(name, b2pt, ctx):
    """The workload BASELINE.json's metric is quoted on is c2; fall back (loudly) only if that variant is not built."""
    if name:
        return name, WORKLOADS[name]
    order = ["c2", "c1"]
    for n in order:
        w = WORKLOADS[n]
        try:
            r = b2pt.Renderer(ctx, 16, 16, w["variant"], precise=False)
            r.set_scene_file(scene_file(w))
            r.set_luts_file(ROOT / "data" / "brdf_luts.b2luts")
            r.render(1)
            r.close()
            return n, w
        except b2pt.B2ptError as e:
            print(f"bench: workload {n} unavailable ({e})", file=sys.stderr)
    raise SystemExit("bench: no workload can run")


def reference_arm(a):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = a.workload or "c2"
    w = dict(WORKLOADS[name])
    exe = ROOT / "oracle" / "_ref" / f"ref_render_{w['variant']}"
    if not exe.exists():
        entry.build_oracles()
    entry.ensure_data()
    cores = host_cores()
    # bounded sample: a centred window of the frame, 1 spp per step, sized for ~1-3 s per step
    big = w["scene"].startswith("synthetic:")  # the reference's octree build over 10M triangles takes ~12 s per run
    win_w, win_h = (min(w["width"], 160), min(w["height"], 90)) if big else (min(w["width"], 480), min(w["height"], 270))
    x0, y0 = (w["width"] - win_w) // 2, (w["height"] - win_h) // 2
    spp = 1
    times = []
    for i in range(a.warmup + a.steps):
        out = subprocess.run([str(exe), "--scene", str(scene_file(w)), "--luts", str(ROOT / "data" / "brdf_luts.b2luts"),
                              "--width", str(w["width"]), "--height", str(w["height"]), "--spp", str(spp), "--bounces", str(w["bounces"]),
                              "--threads", str(cores), "--window", str(x0), str(y0), str(x0 + win_w), str(y0 + win_h), *oracle_args(w)],
                             check=True, capture_output=True, text=True)
        meta = json.loads(out.stdout.strip().splitlines()[-1])
        if i >= a.warmup:
            times.append(meta["render_seconds"])
    total = sum(times)
    value = win_w * win_h * spp * len(times) / total / 1e6
    sample = f"{win_w}x{win_h} centre window of the {w['width']}x{w['height']} frame, {spp} spp per step, {w['bounces']} bounces"
    line = dict(metric=METRIC, value=value, unit="Mpaths/s", n_gpus=a.gpus, steps=a.steps, warmup=a.warmup, ms_per_step=1e3 * total / len(times),
                higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=w["desc"], variant=w["variant"], width=w["width"], height=w["height"], spp=w["spp"], bounces=w["bounces"]),
                cpu_baseline=dict(value=value, unit="Mpaths/s", cores=cores, kind="reference", sample=sample),
                e2e=dict(value=value, unit="Mpaths/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


def cpu_baseline(w, seconds_budget=20.0):
    exe = ROOT / "oracle" / "_ref" / f"ref_render_{w['variant']}"
    kind = "reference"
    extra = []
    if not exe.exists():
        exe = ROOT / "oracle" / "_port" / "port_render"
        kind = "port"
        extra = ["--variant", w["variant"]]
    if not exe.exists():
        return dict(value=None, unit="Mpaths/s", cores=0, kind="unavailable", sample="oracle binaries not built")
    cores = host_cores()
    big = w["scene"].startswith("synthetic:")
    win_w, win_h = (min(w["width"], 160), min(w["height"], 90)) if big else (min(w["width"], 480), min(w["height"], 270))
    x0, y0 = (w["width"] - win_w) // 2, (w["height"] - win_h) // 2
    spp, total_s, total_paths = (2 if big else 1), 0.0, 0
    t_start = time.time()
    while True:
        out = subprocess.run([str(exe), *extra, "--scene", str(scene_file(w)), "--luts", str(ROOT / "data" / "brdf_luts.b2luts"),
                              "--width", str(w["width"]), "--height", str(w["height"]), "--spp", str(spp), "--bounces", str(w["bounces"]),
                              "--threads", str(cores), "--window", str(x0), str(y0), str(x0 + win_w), str(y0 + win_h), *oracle_args(w)],
                             check=True, capture_output=True, text=True)
        meta = json.loads(out.stdout.strip().splitlines()[-1])
        total_s += meta["render_seconds"]
        total_paths += win_w * win_h * spp
        if big or time.time() - t_start > seconds_budget * 0.5 or total_s > seconds_budget:
            break
        spp = min(spp * 2, 16)
    return dict(value=total_paths / total_s / 1e6, unit="Mpaths/s", cores=cores, kind=kind,
                sample=f"{win_w}x{win_h} centre window of the {w['width']}x{w['height']} frame, {total_paths // (win_w * win_h)} spp total, {w['bounces']} bounces, {total_s:.1f} s of CPU render time")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default=None, choices=[None, *WORKLOADS.keys()])
    ap.add_argument("--precise", type=int, default=0, help="1 = time the parity flavour instead of the fast flavour")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--spp", type=int, default=None)
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl != "reference" else a.warmup

    if a.impl == "reference":
        reference_arm(a)
        return

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the CPU reference arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    entry.ensure_data()
    b2pt = entry.load_b2pt()
    ctx = b2pt.Context(local_rank)
    name, w = pick_workload(a.workload, b2pt, ctx)
    w = dict(w)
    if a.spp:
        w["spp"] = a.spp
    W, H, spp = w["width"], w["height"], w["spp"]

    r = b2pt.Renderer(ctx, W, H, w["variant"], precise=bool(a.precise))
    if rank == 0:
        scene_file(w), envmap_file(w)   # generate once, before the other ranks look for the files
    if world > 1:
        dist.barrier()
    r.set_scene_file(scene_file(w))
    r.set_luts_file(ROOT / "data" / "brdf_luts.b2luts")
    r.set(nb_bounces=w["bounces"])
    if w["variant"] == "c4":
        r.set(restir_di_do_fused_spatiotemporal=int(w.get("fused", 0)))
    if envmap_file(w):
        r.set_envmap_file(envmap_file(w))
    if world > 1:
        r.set_partition(rank, world, torch, dist)
    info = r.scene_info()

    fb_host = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
    fb_np = fb_host.numpy()
    gather = None
    if world > 1:
        gather = r.make_gather(torch, dist)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(e2e):
        """Returns the device milliseconds of the step: the renderer's CUDA events on its stream (+ the all-gather's
        events on torch's stream when N > 1; the two regions do not overlap)."""
        r.reset()
        r.set_host_seed(42)
        if e2e:
            r.reupload_scene()
        r.render(spp, sync=False)
        gather_ms = gather() if gather is not None else 0.0
        if e2e:
            r.framebuffer(out=fb_np)
        r.synchronize()
        return r.last_render_ms() + gather_ms

    # ---- device-timed region (inputs resident in HBM) -------------------------------------------------------
    for _ in range(a.warmup):
        step(False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms = 0.0
    launches_before = sum(v["launches"] for v in r.pass_times().values())
    for _ in range(a.steps):
        dev_ms += step(False)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = sum(v["launches"] for v in r.pass_times().values()) - launches_before
    clocks = sampler.stop() if rank == 0 else None
    # the renderer's own CUDA events bracket every step on its stream; the wall clock around the barrier pair is an upper bound
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())

    # ---- end to end through the C ABI with host buffers -----------------------------------------------------
    for _ in range(2):
        step(True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step(True)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    h2d = r.scene_upload_bytes()
    d2h = W * H * 12

    paths_per_step = W * H * spp
    value = paths_per_step * a.steps / (dev_ms * 1e-3) / 1e6
    e2e_value = paths_per_step * a.steps / (e2e_ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (MegaKernel), outside the timed region. Every rank renders (the ReSTIR DI variants
    # run collectives inside render()); rank 0 reports. ---------------------------------------------------------------
    roofline = None
    extra = {}
    r.reset()
    r.set_host_seed(42)
    r.set_profile_passes(True)
    before = r.pass_times()
    r.render(spp)
    after = r.pass_times()
    r.set_profile_passes(False)
    pass_ms = {k: after[k]["ms"] - before.get(k, dict(ms=0.0))["ms"] for k in after}
    n_launch = {k: after[k]["launches"] - before.get(k, dict(launches=0))["launches"] for k in after}
    r.reset()
    r.set_host_seed(42)
    r.set_counters(True)
    ctx.counters_reset()
    r.render(spp)
    c = ctx.counters_read()
    r.set_counters(False)
    if rank == 0:
        if n_launch.get("MegaKernel", 0) <= 0 or pass_ms.get("MegaKernel", 0.0) <= 0.0:
            raise RuntimeError(f"no MegaKernel launch was timed by the profiled render (passes seen: {sorted(pass_ms)}): the roofline would be empty")
        mk_ms = pass_ms["MegaKernel"] / n_launch["MegaKernel"]
        own = r.owned_pixels() if world > 1 else W * H
        rays = c["rays"] / spp   # per launch (one launch = one sample of every pixel)
        # B_path: traversal bytes + per-vertex hit reconstruction + light candidates + per-pixel G-buffer / framebuffer traffic.
        # CameraRays' primary rays are part of `rays` but belong to its own launch: subtract one ray per pixel.
        mk_rays = max(rays - own, 0.0)
        frac_mk = mk_rays / max(rays, 1.0)
        trav_bytes = frac_mk * (c["nodes"] * S_NODE + c["triangles"] * S_TRI) / spp
        vertices = c.get("vertices", 0) / spp if c.get("vertices") else mk_rays / 3.0
        lights = c.get("light_samples", 0) / spp if c.get("light_samples") else vertices
        algo_bytes = trav_bytes + vertices * S_HIT + lights * S_LIGHT + own * S_PX
        peak, which = measured_peak_gbs()
        achieved = algo_bytes / (mk_ms * 1e-3) / 1e9 if mk_ms > 0 else 0.0
        roofline = dict(bound="hbm", kernel="MegaKernel", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=None,
                        peak_source=which, algorithmic_bytes_per_launch=algo_bytes, launch_ms=mk_ms,
                        note=("10M-triangle scene: 0.6 GB of nodes + triangles, random-sector HBM/L2 gathers bind" if w["scene"].startswith("synthetic:")
                              else "Cornell (BVH+triangles < 1 MB) is L2-resident: issue rate, not HBM, binds; see DESIGN.md"))
        # dram__bytes_read.sum + dram__bytes_write.sum of one MegaKernel launch from the committed `ncu --set full` capture
        traffic_file = ROOT / "profiles" / "ncu_traffic.json"
        if traffic_file.exists():
            roofline["traffic"] = json.loads(traffic_file.read_text()).get(name, {}).get("dram_bytes_per_launch")
        extra = dict(pass_ms_per_step=pass_ms, mrays_per_s=c["rays"] / (sum(pass_ms.values()) * 1e-3) / 1e6 if pass_ms else None,
                     traversal=dict(rays_per_path=c["rays"] / (own * spp), nodes_per_ray=c["nodes"] / max(c["rays"], 1), triangles_per_ray=c["triangles"] / max(c["rays"], 1)),
                     scene=info, wall_ms_timed_region=wall_ms)

    # ---- the same frame with the OTHER codegen flavour, one step, N=1 only: `value` is the fast flavour (FMA contraction + fast
    # intrinsics; statistically equal to the reference), the parity flavour (no contraction, IEEE div/sqrt, C-library calls evaluated
    # in double like the reference's host build) is the one that is bit-identical to it -- both numbers belong next to each other.
    other = None
    if world == 1 and not a.no_cpu_baseline:
        r2 = b2pt.Renderer(ctx, W, H, w["variant"], precise=not bool(a.precise))
        r2.set_scene_file(scene_file(w))
        r2.set_luts_file(ROOT / "data" / "brdf_luts.b2luts")
        r2.set(nb_bounces=w["bounces"])
        if w["variant"] == "c4":
            r2.set(restir_di_do_fused_spatiotemporal=int(w.get("fused", 0)))
        if envmap_file(w):
            r2.set_envmap_file(envmap_file(w))
        ms2 = []
        for i in range(3):
            r2.reset()
            r2.set_host_seed(42)
            r2.render(spp, sync=False)
            r2.synchronize()
            ms2.append(r2.last_render_ms())
        r2.close()
        other = dict(flavour="fast" if a.precise else "precise", value=paths_per_step / (min(ms2[1:]) * 1e-3) / 1e6, unit="Mpaths/s", ms_per_step=min(ms2[1:]),
                     note="one warm-up + best of two steps; the precise flavour is the one asserted bit-identical to the reference (DESIGN.md section 3)")

    if rank == 0:
        cpu = dict(value=None, unit="Mpaths/s", cores=0, kind="skipped", sample="--no-cpu-baseline") if a.no_cpu_baseline or world > 1 else cpu_baseline(w)
        line = dict(metric=METRIC, value=value, unit="Mpaths/s", n_gpus=world, steps=a.steps, warmup=a.warmup, ms_per_step=dev_ms / a.steps,
                    higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32", data="synthetic",
                    config=dict(workload=w["desc"], variant=w["variant"], width=W, height=H, spp=spp, bounces=w["bounces"],
                                flavour="precise" if a.precise else "fast", partition="interleaved 16-row bands, scene replicated" if world > 1 else "single GPU",
                                l2_policy="every step re-renders from reset: G-buffer + framebuffer + aux (> 126 MB at 1080p) are rewritten between timed iterations"),
                    clocks=clocks,
                    e2e=dict(value=e2e_value, unit="Mpaths/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=e2e_ms / a.steps),
                    gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu, gpaths_bounce_per_s=value * (w["bounces"] + 1) / 1000.0, other_flavour=other, **extra)
        print(json.dumps(line))

    r.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
