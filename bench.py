#!/usr/bin/env python3
"""bench.py — headline benchmark of the BT.709 4:2:0 -> sRGB BGRA8 decode path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload c2_1080p|c3_uhd|c2_rescale|c2_twopass|c2_fused|c4_alpha|c5_encode]
                    [--gamma bt709|apple|srgb|linear] [--no-configs] [--no-e2e] [--no-live-traffic]

Metric (BASELINE.json): Gpix/s of BT.709 4:2:0 -> sRGB BGRA8, plus % of the HBM roofline.
A "step" is one pass of the hot path over one resident batch of synthetic frames (configs[1]:
600 x 1920x1080 per GPU, 1.87 GB in / 4.98 GB out, far larger than the 126 MB L2, so no flush is
needed between steps).  `value` counts pixels of all ranks / max-over-ranks device time.
Multi-GPU: one process per GPU (torchrun), frames sharded by index, no data-path collective;
torch.distributed is used only for the barrier and for reducing the timing scalars.

The JSON line also carries `configs`: every other BASELINE configuration (UHD batch, RGB + alpha, pass 2 alone,
two-pass and fused decode -> rescale, the 4:2:0 encode) timed in the same run, each with its own value,
roofline fraction and clock record (single-GPU runs only; --no-configs skips them).
`roofline.traffic` of the headline is read from the DRAM counters in the same run (a child process under
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`, outside the timed regions; --no-live-traffic or a missing
ncu falls back to the committed constants of profiles/traffic.json, and `traffic_source` says which).

`e2e` is the same metric through the reference-facing host API (MetalBT709Decoder_decodeBT709
with pinned HOST buffers): upload, kernel and read-back are all inside the timed region; its output is
compared with the device-resident kernel's output for the same frames.

--impl reference times the reference's own CPU implementation of the path (oracle/_ref: the
unmodified BT709.h compiled where it lies; else the oracle port) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OUT_W, OUT_H = 2560, 1440  # pass-2 target of configs[1]
WORKLOADS = {
    # name: (w, h, frames per GPU, kind)
    "c2_1080p": (1920, 1080, 600, "decode"),
    "c3_uhd": (3840, 2160, 300, "decode"),       # 2400 frames / 8 GPUs per BASELINE configs[2]
    "c2_rescale": (1920, 1080, 200, "rescale"),  # pass 2 alone, 1080p -> 2560x1440
    "c2_twopass": (1920, 1080, 200, "twopass"),  # decode -> same-size intermediate in HBM -> rescale (two launches)
    "c2_fused": (1920, 1080, 200, "fused"),      # bt709_decode_rescale_nv12_fused: one kernel, no intermediate image
    "c4_alpha": (2048, 1536, 200, "alpha"),
    "c5_encode": (2048, 1536, 200, "encode"),
}
KIND_TEXT = {
    "decode": "NV12 video-range frames -> sRGB BGRA8 (decode, pass 1)",
    "alpha": "RGB NV12 + alpha-stream Y plane -> BGRA8 with A (decode, sRGB gamma forced)",
    "rescale": f"sRGB BGRA8 frames -> {OUT_W}x{OUT_H} (pass 2 alone, bilinear in linear light)",
    "twopass": f"NV12 frames -> sRGB BGRA8 intermediate -> {OUT_W}x{OUT_H} (decode + rescale, two launches)",
    "fused": f"NV12 frames -> {OUT_W}x{OUT_H} sRGB BGRA8 (fused decode -> rescale, one launch)",
    "encode": "sRGB BGRA8 frames -> NV12 video-range (4:2:0 encode, sRGB in / Apple out)",
}
GAMMAS = {"apple": 0, "srgb": 1, "linear": 2, "bt709": 3}


def _dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampler running during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,clocks.mem")

    def __init__(self, index, lms=100):
        self.index = index
        self.lms = lms
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", str(self.lms)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, mem, pw = [], [], set(), [], []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
            try:
                pw.append(float(f[3]))
                if len(f) > 9:
                    mem.append(float(f[9]))
            except ValueError:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        if mem:
            out["mem_mhz"] = sorted(mem)[len(mem) // 2]
        if pw:
            out["power_w_max"] = max(pw)
        return out


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kind, pixels_per_launch):
    """DRAM bytes per launch of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum of the
    committed `ncu --set full` capture (profiles/traffic.json, bytes per pixel) scaled to this launch."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            t = json.load(open(p))
            if kind == "twopass":  # two kernels per step: the sum of both, per step
                a, b = t.get("decode_bytes_per_pixel"), t.get("rescale_bytes_per_pixel")
                return None if a is None or b is None else (a + b) * pixels_per_launch * 2
            bpp = t.get(kind + "_bytes_per_pixel")
            return None if bpp is None else bpp * pixels_per_launch
        except Exception:
            return None
    return None


_KERNELS_OF = {"decode": "k_decode", "alpha": "k_decode", "rescale": "k_rescale", "fused": "k_fused", "encode": "k_encode",
               "twopass": "k_decode|k_rescale"}


def parse_ncu_dram_bytes(csv_path):
    """(sum of dram__bytes_read.sum + dram__bytes_write.sum in bytes, number of counter rows) of an `ncu --csv` log"""
    import csv
    rows = [x for x in csv.reader(open(csv_path)) if len(x) > 10]
    if not rows:
        return 0.0, 0
    ix = {n: i for i, n in enumerate(rows[0])}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    total, seen = 0.0, 0
    for x in rows[1:]:
        if x[ix["Metric Name"]] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            total += float(x[ix["Metric Value"]].replace(",", "")) * scale[x[ix["Metric Unit"]]]
            seen += 1
    return total, seen


def live_traffic(workload, gamma_name, kind, w, h):
    """DRAM bytes per pixel of this workload's kernel(s), read from the hardware counters IN THIS RUN: a child process
    (`bench.py --traffic-probe`, three untimed steps of the same workload on a batch still far larger than L2) runs under
    `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`, one launch (two for the two-pass step) is captured after
    two warm-up steps.  Outside every timed region.  Returns (bytes per pixel, note) or (None, why not)."""
    import shutil
    import tempfile
    if any(k.startswith(("NV_NSIGHT", "NVTX_INJECTION", "CUDA_INJECTION")) for k in os.environ):
        return None, "running under a profiler"
    ncu = shutil.which("ncu") or ("/usr/local/cuda/bin/ncu" if os.path.exists("/usr/local/cuda/bin/ncu") else None)
    if not ncu:
        return None, "ncu not found"
    frames = max(8, int(300 * (1920 * 1080) / float(w * h)))  # >= 3 GB per launch: the L2 carry-over is ~2 %, as in the bench
    per_step = 2 if kind == "twopass" else 1
    log = tempfile.NamedTemporaryFile(suffix=".csv", delete=False).name
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k",
           "regex:" + _KERNELS_OF[kind], "-s", str(2 * per_step), "-c", str(per_step), "--csv", "--log-file", log,
           sys.executable, os.path.abspath(__file__), "--traffic-probe", "--workload", workload, "--gamma", gamma_name,
           "--frames", str(frames)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=180)
        if r.returncode != 0:
            return None, "ncu probe failed (rc %d)" % r.returncode
        total, seen = parse_ncu_dram_bytes(log)
        if seen != 2 * per_step:
            return None, "ncu probe captured %d of %d counters" % (seen, 2 * per_step)
        return total / (float(frames) * w * h), ("dram__bytes_read.sum + dram__bytes_write.sum read by ncu in this run (child process, "
                                                 "%d frames, %d launch%s captured after two warm-up steps), scaled to this launch"
                                                 % (frames, per_step, "" if per_step == 1 else "es"))
    except Exception as e:  # the constants of profiles/traffic.json stay
        return None, "ncu probe failed (%s)" % type(e).__name__
    finally:
        try:
            os.unlink(log)
        except OSError:
            pass


def run_traffic_probe(args):
    """child of live_traffic: three steps of the workload, nothing printed"""
    import torch
    import metalbt709decoder_b200 as pkg
    w, h, frames, kind = WORKLOADS[args.workload]
    frames = args.frames or frames
    gamma = 1 if kind == "alpha" else GAMMAS[args.gamma]
    ctx = pkg.BT709Context(0)
    d = make_batch(torch, w, h, frames, 709, kind)
    step = make_step(ctx, d, kind, gamma, torch.cuda.current_stream())
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    return 0


# ------------------------------------------------------------------------- CPU reference
def _cpu_decode_frames(gamma, w, h, nframes, threads, keep=None):
    """Decode `nframes` synthetic 1080p-class frames with the reference's CPU code on `threads`
    host threads (one frame per task).  Returns (seconds, kind).  keep: a dict that receives the first frame's
    input and CPU output (the parity check of the cpu_baseline leg compares the GPU's output with it)."""
    import ctypes
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _oracle import Oracle, frame_random, ref_oracle
    ref = ref_oracle()
    kind = "reference" if ref is not None else "port"
    impl = ref if ref is not None else Oracle()
    frames = [frame_random(w, h, seed=709 + i) for i in range(min(nframes, 4))]
    idx = iter(range(nframes))
    lock = threading.Lock()

    def work():
        while True:
            with lock:
                i = next(idx, None)
            if i is None:
                return
            y, c = frames[i % len(frames)]
            if kind == "reference":
                out = impl.decode_nv12(gamma, y, c)
            else:
                out = impl.decode_nv12(gamma, y, c, nthreads=1)
            if keep is not None and i == 0:
                keep["y"], keep["c"], keep["out"] = y, c, out

    ths = [threading.Thread(target=work) for _ in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return time.perf_counter() - t0, kind


def cpu_baseline(gamma, w, h, budget_core_s=16.0, gpu_decode=None):
    """The reference's CPU path timed on the host cores.  gpu_decode(y, c) -> (h, w, 4) uint8: when given, the GPU's output
    for the first frame of the sample is compared with the CPU's (north_star: within 1 LSB, fraction of bit-exact
    pixels reported) - the checker at work inside its own leg, not on the timed path."""
    cores = os.cpu_count() or 1
    # ~0.4 core-seconds per 1080p frame with a transfer curve, ~0.09 without
    per_frame = (0.42 if gamma != 1 else 0.09) * (w * h) / (1920 * 1080)
    nframes = max(cores, int(budget_core_s * max(1, cores // 8) / per_frame))
    nframes = min(nframes, cores * 8)
    keep = {} if gpu_decode is not None else None
    secs, kind = _cpu_decode_frames(gamma, w, h, nframes, cores, keep)
    out = {"value": nframes * w * h / secs / 1e9, "unit": "Gpix/s", "cores": cores, "kind": kind,
           "sample": f"{nframes} synthetic {w}x{h} frames, one frame per task on {cores} host threads"}
    if keep:
        import numpy as np
        got = gpu_decode(keep["y"], keep["c"])
        d = np.abs(got.astype(np.int16) - keep["out"].astype(np.int16))
        out["parity"] = {"exact_fraction": float((d == 0).mean()), "max_abs_diff": int(d.max()), "pixels": int(w * h),
                         "against": "the reference's BT709.h / sRGB.h compiled where they lie (oracle/_ref)" if kind == "reference"
                                    else "oracle port", "what": "bt709_decode_nv12 output of the sample's first frame, all four channels"}
    return out, secs


def workload_text(name, frames, gamma_name):
    w, h, _, kind = WORKLOADS[name]
    g = "srgb" if kind == "alpha" else gamma_name
    extra = "" if kind in ("rescale", "encode") else f", gamma={g}"
    return f"{name}: {frames} x {w}x{h} {KIND_TEXT[kind]}{extra}"


def config_of(args, frames, world, fps=None):
    """The `config` object of the JSON line; the reference arm prints the SAME object for the same arguments."""
    c = {"workload": workload_text(args.workload, frames, args.gamma), "frames_per_gpu": frames,
         "l2": "inputs+outputs per step >> 126 MB L2; no flush needed",
         "parallelism": f"frame-sharded x{world}, no collective"}
    if fps is not None:
        c["fps"] = fps
    return c


def run_reference(args):
    rank, _, world = _dist_env()
    if rank != 0:
        return 0
    w, h, frames, kind = WORKLOADS[args.workload]
    if args.frames:
        frames = args.frames
    gamma = GAMMAS[args.gamma]
    cores = os.cpu_count() or 1
    per_step = max(cores, min(cores * 2, 256))
    for _ in range(args.warmup if args.warmup < 2 else 1):
        _cpu_decode_frames(gamma, w, h, max(1, cores // 2), cores)
    t = 0.0
    kind_impl = "port"
    for _ in range(args.steps):
        s, kind_impl = _cpu_decode_frames(gamma, w, h, per_step, cores)
        t += s
    value = per_step * args.steps * w * h / t / 1e9
    sample = (f"{per_step} synthetic {w}x{h} frames per step (a bounded sample of the {frames}-frame batch), "
              f"one frame per task on {cores} host threads")
    line = {
        "impl": "reference", "metric": "Gpix/s BT.709 4:2:0->sRGB BGRA8", "value": value, "unit": "Gpix/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (double pow), u8 I/O",
        "data": "synthetic",
        "config": config_of(args, frames, max(1, world)),
        "note": "reference CPU path (BT709.h / sRGB.h) on the host cores; each step is a bounded sample of the workload",
        "cpu_baseline": {"value": value, "unit": "Gpix/s", "cores": cores, "kind": kind_impl, "sample": sample},
        "e2e": {"value": value, "unit": "Gpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------ GPU arm
def make_batch(torch, w, h, frames, seed, kind):
    """F-random synthetic frames generated on the device: Y~U{16..235}, Cb,Cr~U{16..240}."""
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    d = {}
    if kind in ("decode", "alpha", "twopass", "fused"):
        d["y"] = torch.randint(16, 236, (frames, h, w), dtype=torch.uint8, device="cuda", generator=g)
        d["c"] = torch.randint(16, 241, (frames, h // 2, w), dtype=torch.uint8, device="cuda", generator=g)
    if kind in ("decode", "alpha", "twopass"):
        d["out"] = torch.empty((frames, h, w, 4), dtype=torch.uint8, device="cuda")
    if kind == "alpha":
        d["a"] = torch.randint(16, 236, (frames, h, w), dtype=torch.uint8, device="cuda", generator=g)
    if kind == "fused":
        # work area of the fused kernel (the reference's _resizeTexture argument): ~60 MB of it are used, as per-SM
        # slices that stay in L2; the intermediate image is never written out
        d["scratch"] = torch.empty((frames, h, w, 4), dtype=torch.uint8, device="cuda")
    if kind in ("rescale", "twopass", "fused"):
        if kind == "rescale":
            d["src"] = torch.randint(0, 256, (frames, h, w, 4), dtype=torch.uint8, device="cuda", generator=g)
            d["src"][..., 3] = 255  # what pass 1 hands over for opaque video (AAPLShaders.metal:243)
        d["big"] = torch.empty((frames, OUT_H, OUT_W, 4), dtype=torch.uint8, device="cuda")
    if kind == "encode":
        d["src"] = torch.randint(0, 256, (frames, h, w, 4), dtype=torch.uint8, device="cuda", generator=g)
        d["y"] = torch.empty((frames, h, w), dtype=torch.uint8, device="cuda")
        d["c"] = torch.empty((frames, h // 2, w), dtype=torch.uint8, device="cuda")
    return d


def algorithmic_bytes(kind, w, h, frames):
    """bytes the workload has to move per step, independent of how a kernel moves them."""
    px = w * h * frames
    opx = OUT_W * OUT_H * frames
    if kind == "decode":
        return 5.5 * px
    if kind == "alpha":
        return 6.5 * px
    if kind == "encode":
        return 5.5 * px
    if kind == "rescale":
        return 4.0 * px + 4.0 * opx
    if kind == "twopass":
        return 5.5 * px + 4.0 * px + 4.0 * opx
    if kind == "fused":
        return 1.5 * px + 4.0 * opx
    raise ValueError(kind)


def make_step(ctx, d, kind, gamma, stream):
    if kind == "decode":
        return lambda: ctx.decode_nv12(gamma, d["y"], d["c"], d["out"], stream=stream)
    if kind == "alpha":
        return lambda: ctx.decode_nv12(1, d["y"], d["c"], d["out"], alpha_y=d["a"], stream=stream)
    if kind == "rescale":
        return lambda: ctx.rescale_bgra(d["src"], d["big"], stream=stream)
    if kind == "twopass":
        # the policy entry (AAPLRenderer.m:940-977): pass 1 into the same-size intermediate, pass 2 out of it
        return lambda: ctx.decode_rescale_nv12(gamma, d["y"], d["c"], d["out"], d["big"], stream=stream)
    if kind == "fused":
        return lambda: ctx.decode_rescale_nv12_fused(gamma, d["y"], d["c"], d["scratch"], d["big"], stream=stream)
    if kind == "encode":
        return lambda: ctx.encode_bgra_to_nv12(1, 0, d["src"], d["y"], d["c"], stream=stream)
    raise ValueError(kind)


def roofline_of(kind, w, h, frames, avg_step_s, launches_per_step, live=None):
    """live: (bytes per pixel, note) from live_traffic, else the committed constants of profiles/traffic.json.
    `traffic` is per launch like `achieved` (two-pass: the mean of its two launches)."""
    peak, peak_src = measured_peak_gbs()
    ach = algorithmic_bytes(kind, w, h, frames) / avg_step_s / 1e9
    if live and live[0] is not None:
        traffic, traffic_src = live[0] * float(frames) * w * h / launches_per_step, live[1]
    else:
        traffic = ncu_traffic(kind, float(frames) * w * h / launches_per_step)
        traffic_src = ("profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per pixel of the "
                       "committed `ncu --set full` capture of this kernel, scaled to this launch")
        if live and live[1]:
            traffic_src += " (live probe: %s)" % live[1]
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": traffic, "peak_source": peak_src,
            "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": algorithmic_bytes(kind, w, h, frames) / launches_per_step,
            "avg_launch_ms": avg_step_s * 1e3 / launches_per_step, "kernels_per_step": launches_per_step}


def timed_steps(torch, ctx, step, stream, steps, device_index, lms=25):
    """`steps` steps bracketed by CUDA events and a synchronize on both sides, nvidia-smi clocks sampled meanwhile."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(device_index, lms=lms)
    sampler.start()
    n0 = ctx.launch_count
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    launches = ctx.launch_count - n0
    clocks = sampler.stop()
    return e0.elapsed_time(e1) / 1e3, launches, clocks


def encode_roundtrip_histogram(torch, ctx, w, h):
    """north_star: "encode round-trip error reported as a histogram".  One frame of random colours, constant over each 2x2
    block (so that the 4:2:0 subsampling loses nothing and what remains is the quantisation round trip, like the reference's
    per-colour histograms, CoreImageMetalFilterTests.m:420-813): sRGB BGRA -> bt709_encode_bgra_to_nv12 (sRGB in, Apple
    out) -> bt709_decode_nv12 (Apple curve) -> |difference| per B, G, R byte.  Device only, untimed."""
    g = torch.Generator(device="cuda")
    g.manual_seed(1886)
    blocks = torch.randint(0, 256, (1, h // 2, w // 2, 4), dtype=torch.uint8, device="cuda", generator=g)
    src = blocks.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2).contiguous()
    src[..., 3] = 255
    y = torch.empty((1, h, w), dtype=torch.uint8, device="cuda")
    c = torch.empty((1, h // 2, w), dtype=torch.uint8, device="cuda")
    back = torch.empty((1, h, w, 4), dtype=torch.uint8, device="cuda")
    ctx.encode_bgra_to_nv12(1, 0, src, y, c, wait=True)
    ctx.decode_nv12(0, y, c, back, wait=True)
    d = (back[..., :3].to(torch.int16) - src[..., :3].to(torch.int16)).abs().flatten()
    hist = torch.bincount(d.to(torch.int64)).cpu().tolist()
    return {"what": "abs difference per B,G,R byte after encode (sRGB->Apple 4:2:0) + decode (Apple->sRGB), 2x2-constant random colours",
            "samples": int(d.numel()), "histogram": {str(i): int(n) for i, n in enumerate(hist) if n},
            "max_abs_diff": len(hist) - 1}


def run_extra_config(torch, ctx, name, gamma, gamma_name, device_index):
    """One of the other BASELINE configurations, timed in the same run, two ways: `value` is a short run after an idle
    gap (like the headline under the driver's --steps 20: the SM clock is still at its maximum), `sustained` is the same
    step repeated for >= 0.4 s, by when the 1000 W power cap has pulled the SM clock down (sw_power_cap in `clocks`)."""
    w, h, frames, kind = WORKLOADS[name]
    stream = torch.cuda.current_stream()
    d = make_batch(torch, w, h, frames, 709, kind)
    step = make_step(ctx, d, kind, gamma, stream)
    for _ in range(3):
        step()
    est, _, _ = timed_steps(torch, ctx, step, stream, 3, device_index)
    est /= 3
    px_step = float(frames) * w * h

    def entry(secs, steps, launches, clocks):
        lps = max(1, launches // steps)
        return {"value": px_step * steps / secs / 1e9, "unit": "Gpix/s", "fps": frames * steps / secs, "steps": steps,
                "ms_per_step": secs / steps * 1e3, "roofline": roofline_of(kind, w, h, frames, secs / steps, lps),
                "gpu_launches": launches, "clocks": clocks}

    time.sleep(0.3)  # idle gap: the boost clock comes back
    b_steps = int(min(100, max(20, 0.06 / max(est, 1e-6))))
    b = entry(*((lambda r: (r[0], b_steps, r[1], r[2]))(timed_steps(torch, ctx, step, stream, b_steps, device_index))))
    s_steps = int(min(1000, max(b_steps, 0.4 / max(est, 1e-6))))
    sres = timed_steps(torch, ctx, step, stream, s_steps, device_index)
    out = {"name": name, "workload": workload_text(name, frames, gamma_name), "value_basis": "input pixels", "warmup": 6}
    out.update(b)
    out["sustained"] = entry(sres[0], s_steps, sres[1], sres[2])
    if kind == "encode":
        out["roundtrip"] = encode_roundtrip_histogram(torch, ctx, w, h)
    del d
    torch.cuda.empty_cache()
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    import metalbt709decoder_b200 as pkg
    from metalbt709decoder_b200.sharding import frame_slice, gather_job_stats

    rank, local_rank, world = _dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    w, h, frames_per_gpu, kind = WORKLOADS[args.workload]
    if args.frames:
        frames_per_gpu = args.frames
    gamma = GAMMAS[args.gamma]
    if kind == "alpha":
        gamma = 1
    lo, hi = frame_slice(frames_per_gpu * world, rank, world)  # weak scaling: fixed frames per GPU
    frames = hi - lo
    ctx = pkg.BT709Context(local_rank)
    d = make_batch(torch, w, h, frames, 709 + lo, kind)
    stream = torch.cuda.current_stream()
    step = make_step(ctx, d, kind, gamma, stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = ctx.launch_count
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    evs[0].record(stream)
    for i in range(args.steps):
        step()
        evs[i + 1].record(stream)
    barrier()
    launches = ctx.launch_count - n0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = evs[0].elapsed_time(evs[-1])
    per_step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    units = float(frames) * w * h * args.steps
    secs, all_units = gather_job_stats(total_ms / 1e3, units)

    # ---- end-to-end through the reference-facing host API (host buffers, copies timed)
    e2e = None
    if kind == "decode" and not args.no_e2e:
        e2e = run_e2e(args, ctx, local_rank, w, h, frames, gamma)
        if world > 1:
            e_secs, e_units = gather_job_stats(e2e["_secs"], e2e["_units"])
            c_secs, _ = gather_job_stats(e2e["_copy_secs"] or 0.0, 0.0)
        else:
            e_secs, e_units = e2e["_secs"], e2e["_units"]
            c_secs = e2e["_copy_secs"] or 0.0
        ceiling = e_units / c_secs / 1e9 if c_secs > 0 else None
        e2e = {"value": e_units / e_secs / 1e9, "unit": "Gpix/s", "h2d_bytes_per_step": e2e["h2d"],
               "copy_only_value": ceiling,
               "frac_of_copy_ceiling": (e_units / e_secs / 1e9) / ceiling if ceiling else None,
               "copy_only_note": "the same bytes per frame over PCIe (pinned host memory, same stream count) with no kernel: "
                                 "what this host and its PCIe links allow for one-frame-per-call traffic at this GPU count",
               "d2h_bytes_per_step": e2e["d2h"], "frames_per_step": e2e["frames"], "steps": e2e["steps"],
               "gamma": e2e["gamma"], "verified": e2e["verified"],
               "api": "MetalBT709Decoder_decodeBT709 (pinned host NV12 in, pinned host BGRA out, one frame per call, "
                      "4-deep stream ring)"}

    # ---- the headline step repeated for >= 0.5 s: what the 1000 W power cap leaves of it (single GPU only)
    sustained = None
    if world == 1 and not args.no_configs:
        n_s = int(min(2000, max(args.steps, 0.5 / max(total_ms / 1e3 / args.steps, 1e-6))))
        s_secs, s_launches, s_clocks = timed_steps(torch, ctx, step, stream, n_s, local_rank)
        sustained = {"value": float(frames) * w * h * n_s / s_secs / 1e9, "unit": "Gpix/s", "steps": n_s,
                     "ms_per_step": s_secs / n_s * 1e3,
                     "roofline": roofline_of(kind, w, h, frames, s_secs / n_s, max(1, s_launches // n_s)), "clocks": s_clocks}

    # ---- the other BASELINE configurations, same run (single GPU only)
    configs = None
    if world == 1 and not args.no_configs:
        del d
        torch.cuda.empty_cache()
        configs = []
        for name in ("c3_uhd", "c4_alpha", "c2_rescale", "c2_twopass", "c2_fused", "c5_encode", "c2_1080p"):
            if name == args.workload:
                continue
            configs.append(run_extra_config(torch, ctx, name, gamma, args.gamma, local_rank))

    if rank == 0:
        avg_step_s = (sum(per_step_ms) / len(per_step_ms)) / 1e3
        launches_per_step = max(1, launches // max(1, args.steps))
        # DRAM traffic of the headline kernel from the counters, in this run (single GPU; outside the timed regions)
        live = live_traffic(args.workload, args.gamma, kind, w, h) if world == 1 and not args.no_live_traffic else None
        def gpu_decode(y_np, c_np):
            ty, tc = torch.from_numpy(y_np).cuda()[None], torch.from_numpy(c_np).cuda()[None]
            o = torch.zeros((1, h, w, 4), dtype=torch.uint8, device="cuda")
            ctx.decode_nv12(gamma, ty, tc, o, wait=True)
            return o[0].cpu().numpy()

        cpu = cpu_baseline(gamma, w, h, gpu_decode=gpu_decode if kind == "decode" else None)[0] if world == 1 else None
        line = {
            "metric": "Gpix/s BT.709 4:2:0->sRGB BGRA8", "value": all_units / secs / 1e9, "unit": "Gpix/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": secs / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 math, u8 I/O",
            "data": "synthetic",
            "config": config_of(args, frames, world),  # the same object the reference arm prints
            "fps": all_units / (w * h) / secs,
            "roofline": roofline_of(kind, w, h, frames, avg_step_s, launches_per_step, live),
            "cpu_baseline": cpu,
            "e2e": e2e if e2e else {"value": None, "unit": "Gpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                                    "note": "e2e is measured on the decode workloads"},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if sustained is not None:
            line["sustained"] = sustained
        if configs is not None:
            line["configs"] = configs
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_e2e(args, ctx, device, w, h, frames, gamma):
    """Per-frame calls of the reference-facing host API with pinned host buffers, on the headline's gamma; the
    read-back of the last pass is compared with the device-resident kernel's output for the same frames."""
    import numpy as np
    from metalbt709decoder_b200 import renderer as R
    import torch
    tf = {0: R.kMBDTransferFunction_ITU_R_709_2, 1: R.kMBDTransferFunction_sRGB, 2: R.kMBDTransferFunction_Linear,
          3: R.kMBDTransferFunction_ITU_R_709_2}
    mg = {0: R.MetalBT709GammaApple, 1: R.MetalBT709GammaSRGB, 2: R.MetalBT709GammaLinear, 3: R.MetalBT709GammaBT709Spec}[gamma]
    mrc = R.MetalRenderContext(device)
    assert mrc.setupMetal()
    dec = R.MetalBT709Decoder()
    dec.metalRenderContext = mrc
    dec.gamma = mg
    ring = 16
    rng = np.random.default_rng(709)
    bufs, texs = [], []
    for i in range(ring):
        pb = R.BGRAToBT709Converter.createCoreVideoYCbCrBuffer(w, h)
        pb.y[:] = rng.integers(16, 236, (h, w), dtype=np.uint8)
        pb.cbcr[:] = rng.integers(16, 241, (h // 2, w), dtype=np.uint8)
        pb.setAttachments(R.kMBDYCbCrMatrix_ITU_R_709_2, tf[gamma])
        bufs.append(pb)
        texs.append(mrc.makeBGRATexture(w, h))
    n = min(frames, 200)
    steps = 3

    def one_pass():
        for i in range(n):
            ok = dec.decodeBT709(bufs[i % ring], None, texs[i % ring], None, w, h, False)
            assert ok
        assert mrc.waitUntilCompleted()

    import torch.distributed as dist
    multi = dist.is_available() and dist.is_initialized()

    def rank_barrier():
        # all ranks start a timed leg together: they share one host and its PCIe root, which is what limits this leg
        torch.cuda.synchronize()
        if multi:
            dist.barrier()

    one_pass()
    rank_barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_pass()
    torch.cuda.synchronize()
    secs = time.perf_counter() - t0
    # ---- the ceiling of this call pattern: the same bytes over PCIe with no kernel in between (pinned host buffers,
    # one H2D of a frame's NV12 planes and one D2H of its BGRA image per frame, same 4 streams)
    copy_secs, copy_pass = None, None
    try:
        hin = [torch.empty(w * h * 3 // 2, dtype=torch.uint8).pin_memory() for _ in range(4)]
        hout = [torch.empty(w * h * 4, dtype=torch.uint8).pin_memory() for _ in range(4)]
        din = [torch.empty(w * h * 3 // 2, dtype=torch.uint8, device="cuda") for _ in range(4)]
        dout = [torch.empty(w * h * 4, dtype=torch.uint8, device="cuda") for _ in range(4)]
        streams = [torch.cuda.Stream() for _ in range(4)]

        def copy_pass():
            for i in range(n):
                k = i % 4
                with torch.cuda.stream(streams[k]):
                    din[k].copy_(hin[k], non_blocking=True)
                    hout[k].copy_(dout[k], non_blocking=True)
            for s_ in streams:
                s_.synchronize()

        copy_pass()
    except Exception:
        copy_pass = None
    rank_barrier()  # every rank reaches this, whether or not its buffers could be allocated
    if copy_pass is not None:
        t1 = time.perf_counter()
        for _ in range(steps):
            copy_pass()
        copy_secs = time.perf_counter() - t1
    # the host path must have produced what the (parity-tested) device-resident kernel produces for the same frames
    verified = True
    for i in (0, ring - 1):
        ty = torch.from_numpy(np.ascontiguousarray(bufs[i].y)).cuda()[None]
        tc = torch.from_numpy(np.ascontiguousarray(bufs[i].cbcr)).cuda()[None]
        out = torch.empty((1, h, w, 4), dtype=torch.uint8, device="cuda")
        ctx.decode_nv12(gamma, ty, tc, out, wait=True)
        verified = verified and bool(np.array_equal(out[0].cpu().numpy(), texs[i].pixels))
    if not verified:
        raise SystemExit("bench.py: e2e output differs from the device-resident kernel's output")
    out = {"_secs": secs, "_units": float(n) * w * h * steps, "h2d": int(n * w * h * 1.5), "d2h": n * w * h * 4,
           "frames": n, "steps": steps, "gamma": args.gamma, "verified": verified, "_copy_secs": copy_secs}
    for b in bufs:
        b.release()
    for t in texs:
        t.release()
    dec.release()
    mrc.release()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2_1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--gamma", default="bt709", choices=sorted(GAMMAS))
    ap.add_argument("--frames", type=int, default=0, help="override frames per GPU")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` array (the other BASELINE configurations)")
    ap.add_argument("--no-live-traffic", action="store_true",
                    help="roofline.traffic from the committed constants instead of an ncu child process in this run")
    ap.add_argument("--traffic-probe", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.traffic_probe:
        return run_traffic_probe(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
