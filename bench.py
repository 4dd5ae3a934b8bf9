#!/usr/bin/env python3
"""bench.py -- Mpaths/s of the per-pixel path-tracing hot path on N B200s (driver contract: one JSON line on rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c3s|c4s|c5|...] [--precise 0|1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the whole hot path (CameraRays -> [ReSTIR DI passes] -> MegaKernel, `spp` samples per pixel) over one
full frame of the workload. Default workload = the north-star target: N = 1 -> "c4s" (BASELINE.json configs[3]: the synthetic
10M-triangle scene at 1080p, ReSTIR DI fused spatiotemporal reuse + HDR envmap alias table, RIS at later bounces, 8 bounces,
4 frames x 1 spp per step); N > 1 -> "c5" (configs[4]: the same scene at 4K, partitioned across the GPUs). The timed kernels
are the PRECISE flavour -- the one tests/ assert bit-identical to the reference CPURenderer; `other_flavour` carries the
fast-math build of the same frame, `workloads` the c2 / c3s lines (configs[1], configs[2]).

 value      : W*H*spp*steps / device time (CUDA events on the renderer's stream, max over ranks), scene resident in HBM
 e2e        : same metric through the C ABI with HOST buffers: per step the scene arrays + BVH blob are re-uploaded from pinned
              host memory (what GPURenderer::set_scene hands to the device layer), the frame is rendered, and the fp32
              framebuffer is read back to the host -- all inside the timed region (wall clock, max over ranks)
 roofline   : MegaKernel's algorithmic bytes (B_path from measured traversal steps) / its CUDA-event time
 cpu_baseline: the reference's own CPU path (oracle/_ref, compiled from /root/reference) on this box's host cores, on a bounded
              pixel window of the same workload, several renders in ONE process (one octree build)
 --impl reference: the same thing as the driver's reference arm: K timed + W warm-up renders of the window in one process

Multi-GPU (N > 1): the frame is partitioned into interleaved pixel-row bands (rank = band % N), scene replicated. No
data-path collective inside a sample except for ReSTIR DI, whose reuse passes read neighbouring pixels: the reservoir rows next
to a rank's bands (halo) are exchanged with NCCL send/recv issued from C on the renderer's stream, and CameraRays also covers
the halo rows. The framebuffer bands are all-gathered (ncclAllGather from C, same stream) at the end of every step, inside the
timed region. Total work is fixed => "scaling": "strong". Rank 0 also renders the same frame alone once, outside the timed
region: `single_gpu_same_workload` (the efficiency denominator on the SAME workload) and `multi_gpu_bit_identical`.
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


def default_workload(n_gpus):
    """The north-star target workloads (BASELINE.json configs[3] on one GPU, configs[4] across GPUs)."""
    return "c4s" if n_gpus <= 1 else "c5"


def workload_config(w):
    """`config` of the JSON line: names the workload only, identical for both arms."""
    return dict(workload=w["desc"], variant=w["variant"], width=w["width"], height=w["height"], spp=w["spp"], bounces=w["bounces"],
                scene=w["scene"], envmap=w.get("envmap"))


def cpu_window(w):
    """Bounded sample of the workload the CPU arm renders per step: the whole frame at 1 spp on the Cornell scenes (~0.6 s on 16
    cores), a centred 256x144 window at 1 spp on the 10M-triangle scene (~2 s)."""
    big = w["scene"].startswith("synthetic:")
    win_w, win_h = (min(w["width"], 256), min(w["height"], 144)) if big else (w["width"], w["height"])
    x0, y0 = (w["width"] - win_w) // 2, (w["height"] - win_h) // 2
    return x0, y0, win_w, win_h


def run_reference(w, repeats, cores):
    """Renders the workload's CPU window `repeats` times in ONE process of oracle/_ref/ref_extras_<variant> (the reference's own
    kernels + octree, compiled from /root/reference; one scene load + BVH build). Returns (seconds per repeat, kind, sample)."""
    exe = ROOT / "oracle" / "_ref" / f"ref_extras_{w['variant']}"
    if not exe.exists() and (Path("/root/reference") / "src").exists():
        entry.build_oracles()
    if not exe.exists():
        return None, "unavailable", "oracle/_ref binaries not built"
    x0, y0, win_w, win_h = cpu_window(w)
    out = subprocess.run([str(exe), "--scene", str(scene_file(w)), "--luts", str(ROOT / "data" / "brdf_luts.b2luts"),
                          "--width", str(w["width"]), "--height", str(w["height"]), "--spp", "1", "--bounces", str(w["bounces"]),
                          "--threads", str(cores), "--window", str(x0), str(y0), str(x0 + win_w), str(y0 + win_h), "--repeat", str(repeats), *oracle_args(w)],
                         check=True, capture_output=True, text=True)
    seconds = None
    for line in out.stdout.splitlines():
        if line.startswith('{"render_seconds_per_repeat"'):
            seconds = json.loads(line)["render_seconds_per_repeat"]
    if not seconds:
        raise RuntimeError("reference oracle printed no timings: " + out.stdout[-500:] + out.stderr[-500:])
    sample = (f"{win_w}x{win_h} {'centre window' if (win_w, win_h) != (w['width'], w['height']) else 'full frame'} of the {w['width']}x{w['height']} frame, "
              f"1 spp per step, {w['bounces']} bounces, {repeats} renders in one process")
    return seconds, "reference", sample


def reference_arm(a):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = a.workload or default_workload(a.gpus)
    w = dict(WORKLOADS[name])
    entry.ensure_data()
    cores = host_cores()
    seconds, kind, sample = run_reference(w, a.warmup + a.steps, cores)
    if seconds is None:
        print(json.dumps(dict(impl="reference", unavailable=sample)))
        return
    timed = seconds[a.warmup:]
    _, _, win_w, win_h = cpu_window(w)
    total = sum(timed)
    value = win_w * win_h * len(timed) / total / 1e6
    line = dict(metric=METRIC, value=value, unit="Mpaths/s", n_gpus=a.gpus, steps=a.steps, warmup=a.warmup, ms_per_step=1e3 * total / len(timed),
                higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=workload_config(w),
                cpu_baseline=dict(value=value, unit="Mpaths/s", cores=cores, kind=kind, sample=sample),
                e2e=dict(value=value, unit="Mpaths/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


def cpu_baseline(w, repeats=6):
    cores = host_cores()
    seconds, kind, sample = run_reference(w, repeats + 1, cores)
    if seconds is None:
        return dict(value=None, unit="Mpaths/s", cores=0, kind=kind, sample=sample)
    _, _, win_w, win_h = cpu_window(w)
    timed = seconds[1:]
    return dict(value=win_w * win_h * len(timed) / sum(timed) / 1e6, unit="Mpaths/s", cores=cores, kind=kind,
                sample=sample + f" (first one untimed), {sum(timed):.1f} s of CPU render time")


def make_renderer(b2pt, ctx, w, precise):
    r = b2pt.Renderer(ctx, w["width"], w["height"], w["variant"], precise=precise)
    r.set_scene_file(scene_file(w))
    r.set_luts_file(ROOT / "data" / "brdf_luts.b2luts")
    r.set(nb_bounces=w["bounces"])
    if w["variant"].startswith("c4"):
        r.set(restir_di_do_fused_spatiotemporal=int(w.get("fused", 0)))
    if envmap_file(w):
        r.set_envmap_file(envmap_file(w))
    return r


def quick_line(b2pt, ctx, w, precise, steps=2):
    """One warm-up + `steps` timed steps of a workload on this GPU, scene resident: a sub-line of the main JSON line."""
    r = make_renderer(b2pt, ctx, w, precise)
    ms = []
    for i in range(steps + 1):
        r.reset()
        r.set_host_seed(42)
        r.render(w["spp"], sync=False)
        r.synchronize()
        ms.append(r.last_render_ms())
    r.close()
    best = min(ms[1:])
    return dict(value=w["width"] * w["height"] * w["spp"] / (best * 1e-3) / 1e6, unit="Mpaths/s", ms_per_step=best, flavour="precise" if precise else "fast",
                config=workload_config(w), note=f"one warm-up + best of {steps} steps, scene resident")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default=None, choices=[None, *WORKLOADS.keys()])
    ap.add_argument("--precise", type=int, default=1, help="1 (default) = the parity flavour, bit-identical to the reference; 0 = the fast-math A/B flavour")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU baseline, the other-flavour line and the c2 / c3s sub-lines")
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
    name = a.workload or default_workload(world)
    w = dict(WORKLOADS[name])
    if a.spp:
        w["spp"] = a.spp
    W, H, spp = w["width"], w["height"], w["spp"]
    precise = bool(a.precise)

    if rank == 0:
        scene_file(w), envmap_file(w)   # generate once, before the other ranks look for the files
    if world > 1:
        dist.barrier()
    r = make_renderer(b2pt, ctx, w, precise)
    if world > 1:
        r.set_partition(rank, world, torch, dist)
    info = r.scene_info()
    partition_text = r.partition_description() if world > 1 else "single GPU"

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
        """Returns the device milliseconds of the step: the renderer's CUDA events on its stream (they bracket the framebuffer
        all-gather too when it is issued from C on that stream; a torch-side gather adds its own events)."""
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
    launches_before = ctx.cuda_launches()
    for _ in range(a.steps):
        dev_ms += step(False)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = ctx.cuda_launches() - launches_before   # CUDA kernels of the passes (the NCCL pack / unpack kernels of N > 1 are extra)
    clocks = sampler.stop() if rank == 0 else None
    # the renderer's own CUDA events bracket every step on its stream; the wall clock around the barrier pair is an upper bound
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())

    # ---- end to end through the C ABI with host buffers -----------------------------------------------------
    e2e_steps = a.steps if a.steps <= 5 else max(5, a.steps // 2)   # a bounded share of the run on the heavy workloads
    for _ in range(2):
        step(True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
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
    e2e_value = paths_per_step * e2e_steps / (e2e_ms * 1e-3) / 1e6

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
    # per-kernel times of the wavefront sequences (scenes beyond L2): CUDA events between the kernels of one more render
    r.reset()
    r.set_host_seed(42)
    ctx.wavefront_profile(True)
    r.render(spp)
    wf = ctx.wavefront_profile_read()
    ctx.wavefront_profile(False)
    # whole-job ray / node / triangle counts: every rank counted its own pixels
    c_all = torch.tensor([c["rays"], c["nodes"], c["triangles"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(c_all, op=dist.ReduceOp.SUM)
    rays_all, nodes_all, tris_all = (float(v) for v in c_all.tolist())
    if rank == 0:
        if n_launch.get("MegaKernel", 0) <= 0 or pass_ms.get("MegaKernel", 0.0) <= 0.0:
            raise RuntimeError(f"no MegaKernel launch was timed by the profiled render (passes seen: {sorted(pass_ms)}): the roofline would be empty")
        mk_ms = pass_ms["MegaKernel"] / n_launch["MegaKernel"]
        own = r.owned_pixels() if world > 1 else W * H
        rays = c["rays"] / spp   # rank 0's, per launch (one launch = one sample of every owned pixel)
        # B_path: traversal bytes + per-vertex hit reconstruction + light candidates + per-pixel G-buffer / framebuffer traffic.
        # Rays of the other passes (CameraRays' primary rays, the ReSTIR passes' candidate / shadow rays) are part of `rays` but belong
        # to their own launches: one primary ray per rendered CameraRays pixel is subtracted; the rest is attributed pro rata.
        camera_pixels = own if not (world > 1 and w["variant"].startswith("c4")) else r.camera_rays_pixels()
        mk_rays = max(rays - camera_pixels, 0.0)
        frac_mk = mk_rays / max(rays, 1.0)
        trav_bytes = frac_mk * (c["nodes"] * S_NODE + c["triangles"] * S_TRI) / spp
        vertices = mk_rays / 3.0
        lights = vertices * (4.0 if w["variant"] in ("c3", "c4", "c4p") else 1.0)   # RIS: 4 light candidates per vertex
        algo_bytes = trav_bytes + vertices * S_HIT + lights * S_LIGHT + own * S_PX
        peak, which = measured_peak_gbs()
        achieved = algo_bytes / (mk_ms * 1e-3) / 1e9 if mk_ms > 0 else 0.0
        roofline = dict(bound="hbm", kernel="MegaKernel", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=None,
                        peak_source=which, algorithmic_bytes_per_launch=algo_bytes, launch_ms=mk_ms,
                        note=("10M-triangle scene: 0.72 GB of nodes + triangles, random-sector HBM/L2 gathers (dependent node fetches) bind" if w["scene"].startswith("synthetic:")
                              else "Cornell (BVH+triangles < 1 MB) is L2-resident: issue rate, not HBM, binds; see DESIGN.md"))
        if wf["trace"]["groups"] > 0:
            # Wavefront form: the MegaKernel pass is ~80 kernels; the dominant one is named from the live per-kernel times. wf_trace
            # (all rays of a bounce from one HBM queue) is the HBM-bound one: algorithmic bytes = node + triangle fetches of the
            # MegaKernel's rays + 64 B ray in / 64 B hit out + 4 B queue entry per ray.
            kernels_ms = {k: v["ms"] / spp for k, v in wf.items() if v["groups"]}
            trace_bytes = trav_bytes + mk_rays * (64 + 64 + 4)
            trace_ms = wf["trace"]["ms"] / spp            # all wf_trace launches of one MegaKernel pass (one sample)
            trace_launches = wf["trace"]["groups"] / spp
            achieved = trace_bytes / (trace_ms * 1e-3) / 1e9
            dominant = max(("trace", "stage1", "stage2"), key=lambda k: wf[k]["ms"])
            roofline = dict(bound="hbm", kernel="wf_trace", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=None, peak_source=which,
                            algorithmic_bytes_per_launch=trace_bytes / trace_launches, launch_ms=trace_ms / trace_launches, launches_per_sample=trace_launches,
                            dominant_kernel_by_time={"trace": "wf_trace", "stage1": "wf_stage1", "stage2": "wf_stage2"}[dominant],
                            kernels_ms_per_sample=kernels_ms,
                            megakernel_pass=dict(achieved=algo_bytes / (mk_ms * 1e-3) / 1e9, algorithmic_bytes=algo_bytes, ms=mk_ms),
                            note="wavefront form (scene beyond L2): per-kernel CUDA-event times measured live; wf_trace = dependent 80-B node fetches over a 0.72 GB BVH "
                                 "(random 32-B sectors); the shading stages park ~1 KB of path state per path in HBM between rays (not algorithmic bytes)")
        # dram__bytes_read.sum + dram__bytes_write.sum of one MegaKernel launch from the committed `ncu --set full` capture
        traffic_file = ROOT / "profiles" / "ncu_traffic.json"
        if traffic_file.exists():
            key = {"c5": "c4s"}.get(name, name) + ("" if precise else "_fast")
            if roofline["kernel"] == "wf_trace":
                key += "_wf_trace"
            roofline["traffic"] = json.loads(traffic_file.read_text()).get(key, {}).get("dram_bytes_per_launch")
        total_pass_ms = sum(pass_ms.values())
        extra = dict(pass_ms_per_step=pass_ms, mrays_per_s=rays_all / (total_pass_ms * 1e-3) / 1e6 if total_pass_ms else None,
                     traversal=dict(rays_per_path=rays_all / (W * H * spp), nodes_per_ray=nodes_all / max(rays_all, 1), triangles_per_ray=tris_all / max(rays_all, 1)),
                     scene=info, wall_ms_timed_region=wall_ms)

    # ---- N > 1: rank 0 renders the SAME frame alone (unpartitioned), outside the timed region: the single-GPU rate on this very
    # workload and the bit-identity of the partitioned render (every rank's gathered framebuffer vs rank 0's own). ---------------
    single = None
    identical = None
    if world > 1:
        r.reset()
        r.set_host_seed(42)
        r.render(spp, sync=False)
        gather()
        r.synchronize()
        multi_fb = r.framebuffer().copy()
        if rank == 0:
            r.set_partition(0, 1, torch, dist)
            ms1 = []
            for i in range(2):
                r.reset()
                r.set_host_seed(42)
                r.render(spp, sync=False)
                r.synchronize()
                ms1.append(r.last_render_ms())
            single_fb = r.framebuffer()
            single = dict(value=paths_per_step / (ms1[-1] * 1e-3) / 1e6, unit="Mpaths/s", ms_per_step=ms1[-1], note="rank 0 alone, whole frame, second of two steps")
            ref_fb = torch.from_numpy(single_fb.view(np.int32).copy()).cuda()
        else:
            ref_fb = torch.empty((H, W, 3), dtype=torch.int32, device="cuda")
        dist.broadcast(ref_fb, src=0)
        same = torch.tensor([int(bool((torch.from_numpy(multi_fb.view(np.int32).copy()).cuda() == ref_fb).all().item()))], device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        identical = bool(same.item())

    # ---- sub-lines, N = 1 only: the other codegen flavour on this workload, and the lighter BASELINE configs ----------------------
    other = None
    sublines = {}
    if world == 1 and not a.no_cpu_baseline:
        r.close()
        r = None
        other = quick_line(b2pt, ctx, w, not precise)
        for sub in ("c2", "c3s"):
            if sub != name:
                sublines[sub] = quick_line(b2pt, ctx, dict(WORKLOADS[sub]), precise, steps=2)

    if rank == 0:
        cpu = dict(value=None, unit="Mpaths/s", cores=0, kind="skipped", sample="--no-cpu-baseline / N > 1") if a.no_cpu_baseline or world > 1 else cpu_baseline(w)
        line = dict(metric=METRIC, value=value, unit="Mpaths/s", n_gpus=world, steps=a.steps, warmup=a.warmup, ms_per_step=dev_ms / a.steps,
                    higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32", data="synthetic",
                    config=workload_config(w),
                    impl_detail=dict(flavour="precise (bit-identical to the reference CPURenderer, tests/)" if precise else "fast (--use_fast_math A/B build)",
                                     partition=partition_text,
                                     l2_policy="every step re-renders from reset: G-buffer + framebuffer + reservoirs (> 126 MB) are rewritten between timed iterations; the scene (0.72 GB BVH) exceeds L2"),
                    clocks=clocks,
                    e2e=dict(value=e2e_value, unit="Mpaths/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=e2e_ms / e2e_steps, steps=e2e_steps),
                    gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu, gpaths_bounce_per_s=value * (w["bounces"] + 1) / 1000.0,
                    other_flavour=other, workloads=sublines, single_gpu_same_workload=single, multi_gpu_bit_identical=identical, **extra)
        print(json.dumps(line))

    if r is not None:
        r.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
