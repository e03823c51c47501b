#!/usr/bin/env python
"""Benchmark of the B200 JPEG encode path (driver contract: one JSON line on stdout).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on host cores

Workload (BASELINE.json configs[2], the configuration the Mpixel/s metric is quoted on): a batch of 256
synthetic 1920x1080 RGB frames, 4:2:0, quality 75 -> 256 JFIF files, sharded across the N GPUs
(rank g encodes frames [g*256/N, (g+1)*256/N): strong scaling, the named batch; `--scaling weak` gives
every rank its own 256 frames instead).  One "step" is one pass of the whole hot path (K1 fused block
kernel; K2-K4 fused entropy kernel: block bits, scan with look-back, bit packing; 0xFF stuffing + JFIF
framing) over that batch.  Frames are independent: no data-path collective.

value : whole-job Mpixel/s (source pixels) with the batch already resident in HBM, CUDA-event timed,
        max over ranks.  The 1.59 GB input is far larger than the 126 MB L2, so every step streams
        from HBM (config.l2 = "input > L2").
e2e   : the same metric through the host-buffer entry point of the C ABI: ONE call of
        b2j_multi_encoder_encode_host on rank 0 over the devices of all ranks (pinned host RGB of the
        whole batch in, JFIF bytes in pinned host memory out, H2D and D2H copies inside the timed
        region), with the box's concurrent host-to-device copy time for the same batch measured
        beside it (e2e.frac_of_h2d_ceiling).
roofline : K1 (the dominant kernel), algorithmic bytes 6 B per padded pixel (3 B RGB read + 1.5
        int16 coefficients written) / its mean duration measured with CUDA events inside the timed
        steps, against MEASURED_PEAKS.json hbm_gbs.
cpu_baseline : the plain-C oracle port (oracle/c/b2j_oracle.c, OpenMP) on this box's host cores on
        a bounded sample of the same batch.  The reference's own Python model cannot travel to the
        GPU box (and runs at 0.069 Mpixel/s); see DESIGN.md.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W, SUB, QUALITY, ZRL = 1080, 1920, "420", 75, "eager"
FRAMES_PER_GPU = 256
UNIQUE_FRAMES = 32
CONFIG_IDX = 2
KIND = "photo"
WORKLOAD = "batch of %d synthetic 1920x1080 RGB frames, 4:2:0, quality 75 (BASELINE configs[2])"
# The other BASELINE.json configs, for the per-config table in profiles/ (the driver's line is the default, configs[2]).
# name: (height, width, subsampling, quality, frames of the batch, seed index, label)
PRESETS = {
    "1": (1080, 1920, "420", 75, 1, 1, "single synthetic 1920x1080 RGB frame, 4:2:0, quality 75 (BASELINE configs[1]); frames=%d"),
    "2": (1080, 1920, "420", 75, 256, 2, WORKLOAD),
    "3a": (2160, 3840, "444", 90, 64, 3, "batch of %d synthetic 3840x2160 frames, 4:4:4, quality 90 (BASELINE configs[3])"),
    "3b": (2160, 3840, "420", 90, 64, 3, "batch of %d synthetic 3840x2160 frames, 4:2:0, quality 90 (BASELINE configs[3])"),
    "4": (4320, 7680, "420", 75, 16, 4, "batch of %d synthetic 7680x4320 frames, 4:2:0, quality 75 (BASELINE configs[4])"),
}
METRIC = "Mpixel/s full JPEG encode (RGB -> JFIF bytes)"
UNIT = "Mpixel/s"
CPU_SAMPLE_FRAMES = 256      # the whole 256-frame step (0.5 s on 16 threads = ~8 CPU-seconds per pass)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=None, help="frames of the batch (default: the named workload)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default): the named batch is sharded across the GPUs; weak: every GPU gets the whole batch size")
    ap.add_argument("--config", default="2", choices=sorted(PRESETS), help="BASELINE.json config (default 2, the headline)")
    ap.add_argument("--dct", default="model", choices=["model", "fixed"],
                    help="model = float64 model arithmetic (parity path, the headline); fixed = [ext] gateware fixed point, non-parity")
    ap.add_argument("--graph", action="store_true",
                    help="replay the step from a CUDA graph even when the input fits the L2 (the single-frame config: no L2 "
                         "flush between the steps then); stage times come from an untimed ungraphed pass")
    ap.add_argument("--no-graph", action="store_true",
                    help="launch every step eagerly (default: the timed steps of a batch larger than the L2 are replayed from a "
                         "CUDA graph of the same four launches, which removes ~20 us of launch gaps per step)")
    ap.add_argument("--restart", type=int, default=0, help="[ext] restart interval in MCUs (0 = none, the headline)")
    ap.add_argument("--kind", default="photo", choices=["photo", "noise", "grey"], help="synthetic content (SURVEY 8d stress)")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    args = ap.parse_args()
    global H, W, SUB, QUALITY, CONFIG_IDX, KIND, WORKLOAD, UNIQUE_FRAMES
    H, W, SUB, QUALITY, frames, CONFIG_IDX, WORKLOAD = PRESETS[args.config]
    KIND = args.kind
    if args.frames is None:
        args.frames = frames
    UNIQUE_FRAMES = max(1, min(UNIQUE_FRAMES, (32 * 1080 * 1920) // (H * W)))     # bound the host-side generation time
    if KIND != "photo":
        WORKLOAD += " [content: %s]" % KIND
    if args.graph:
        WORKLOAD += " [step replayed from a CUDA graph, no L2 flush]"
    if args.dct != "model":
        WORKLOAD += " [dct: %s, NON-PARITY arithmetic mode]" % args.dct
    if args.restart:
        WORKLOAD += " [restart interval: %d MCUs]" % args.restart
    return args


def shard(n_frames, world, rank):
    """Contiguous frame range of `rank` (the rule of litejpeg_b200.shard_range)."""
    return n_frames * rank // world, n_frames * (rank + 1) // world


def k1_source_sha():
    """Short hash of the sources K1 is built from (ties profiles/k1_traffic.json to the code it was captured on)."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "jpeg_encoder_gsoc2017_ishan-bansal_b200", "csrc")
    for name in ("b2j_fused.cu", "b2j_fused_pair.cuh", "b2j_fused_stream.cuh", "b2j_block.cuh", "b2j_math.cuh", "b2j_common.cuh"):
        try:
            h.update(open(os.path.join(d, name), "rb").read())
        except OSError:
            h.update(name.encode())
    return h.hexdigest()[:12]


def workload_config(n_gpus, frames, scaling="strong"):
    total = frames if scaling == "strong" else frames * n_gpus
    per = [shard(total, n_gpus, r)[1] - shard(total, n_gpus, r)[0] for r in range(n_gpus)]
    return {"workload": (WORKLOAD % total) + (", sharded across %d GPUs" % n_gpus if n_gpus > 1 else ""),
            "frames_per_gpu": max(per), "global_frames": total, "height": H, "width": W, "subsampling": SUB,
            "quality": QUALITY, "zrl_policy": ZRL, "unique_frames": min(UNIQUE_FRAMES, total),
            "parallelism": "frames [g*B/G, (g+1)*B/G) on GPU g of %d, no data-path collective" % n_gpus,
            "l2": "input (%.2f GB per GPU) %s L2 (126 MB)%s" % (max(per) * H * W * 3 / 1e9, ">" if max(per) * H * W * 3 > 126e6 else "<=",
                                                            ": every step streams from HBM" if max(per) * H * W * 3 > 126e6
                                                            else "; an L2 flush (256 MB write) runs between the timed steps")}


# ---------------------------------------------------------------------------------- CPU arm
class OraclePort:
    def __init__(self):
        path = os.path.join(ROOT, "oracle", "_build", "libb2j_oracle.so")
        if not os.path.exists(path):
            res = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], capture_output=True, text=True)
            if res.returncode:
                raise RuntimeError("cannot build the C oracle: " + res.stderr)
        self.lib = ctypes.CDLL(path)
        # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1; that is not "all")
        try:
            self.threads = len(os.sched_getaffinity(0))
        except AttributeError:
            self.threads = os.cpu_count() or 1

    def encode_batch(self, frames_np, out_np, sizes_np, cap):
        n = frames_np.shape[0]
        rc = self.lib.b2jo_encode_batch(frames_np.ctypes.data_as(ctypes.c_void_p), n, H, W, int(SUB), QUALITY,
                                        0 if ZRL == "eager" else 1, out_np.ctypes.data_as(ctypes.c_void_p),
                                        ctypes.c_int64(cap), sizes_np.ctypes.data_as(ctypes.c_void_p), self.threads)
        assert rc == 0


def cpu_baseline(frames_torch, steps=1):
    """Times the oracle port on a bounded sample (first CPU_SAMPLE_FRAMES frames), all host threads."""
    import numpy as np
    port = OraclePort()
    n = max(1, min(CPU_SAMPLE_FRAMES * 1080 * 1920 // (H * W), frames_torch.shape[0]))
    sample = frames_torch[:n].numpy()
    cap = H * W * 3 + 4096
    out = np.zeros((n, cap), np.uint8)
    sizes = np.zeros(n, np.int64)
    port.encode_batch(sample[:min(8, n)], out, sizes, cap)          # warm-up (page faults, table init)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        port.encode_batch(sample, out, sizes, cap)
        times.append(time.perf_counter() - t0)
    assert (sizes > 0).all()
    dt = sum(times) / len(times)
    return {"value": n * H * W / dt / 1e6, "unit": UNIT, "cores": port.threads, "kind": "port",
            "sample": "%d of the %d frames (%.1f Mpixel) per step, OpenMP over frames, %d threads; %.2f s per step"
                      % (n, frames_torch.shape[0], n * H * W / 1e6, port.threads, dt)}, dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import importlib
    import litejpeg_b200  # noqa: F401  (only for the synthetic frame generator; no GPU work on this arm)
    synth = importlib.import_module("litejpeg_b200.synth")
    total = args.frames if args.scaling == "strong" else args.frames * max(1, args.gpus)
    n = max(1, min(CPU_SAMPLE_FRAMES * 1080 * 1920 // (H * W), total))
    frames = synth.batch(n, H, W, CONFIG_IDX, KIND, unique=min(UNIQUE_FRAMES, n))
    for _ in range(max(0, args.warmup - 1)):
        cpu_baseline(frames, 1)
    base, dt = cpu_baseline(frames, max(1, args.steps))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus, args.frames, args.scaling), "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU arm = plain-C port of the reference's Python model (oracle/c/b2j_oracle.c); the model itself "
                    "is Python-only, cannot travel to the GPU box and measures 0.069 Mpixel/s (BASELINE.md)"}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock, power and throttle reasons of one GPU through NVML from a background thread
    (every ~2 ms, so that even a 25 ms timed region holds samples).  `window()` brackets the timed
    regions; the median is taken over the samples inside them."""
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu_index):
        import threading
        self.samples, self.windows, self.err = [], [], None
        self._stop = threading.Event()
        self.thread = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(int(gpu_index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception as e:          # no NVML: report that, never fake a clock
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                self.samples.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                                     int(reasons(self.h)), nv.nvmlDeviceGetPowerUsage(self.h) / 1e3))
            except Exception as e:
                self.err = repr(e)
                return
            time.sleep(0.002)

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is None:
            out["error"] = self.err
            return out
        self._stop.set()
        self.thread.join(timeout=2)
        inside = [x for x in self.samples if any(a <= x[0] <= b for a, b in self.windows)]
        if not inside:
            out["error"] = self.err or "no sample fell inside the timed regions"
            return out
        mask = 0
        for x in inside:
            mask |= x[2]
        return {"sm_mhz": statistics.median(x[1] for x in inside), "sm_min_mhz": min(x[1] for x in inside),
                "sm_max_mhz": self.max_mhz, "reasons": [n for bit, n in self.REASONS if mask & bit],
                "samples": len(inside), "power_w_max": max(x[3] for x in inside),
                "source": "NVML, ~2 ms period, samples inside the timed regions (device-timed steps and e2e)"}


# ---------------------------------------------------------------------------------- GPU arm
def bind_to_gpu_numa_node(nvml_index):
    """Pins this rank to the CPUs NVML reports as local to its GPU, so that the pinned host buffers
    (first touch) and the H2D/D2H copies stay on the GPU's own socket.  Returns the previous
    affinity (restored for the CPU-baseline leg, which must use all host threads)."""
    try:
        before = os.sched_getaffinity(0)
    except AttributeError:
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(nvml_index))
        n_words = (max(before) + 64) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1} & before
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass
    return before


def _nvml_index(torch, cuda_index):
    """NVML enumerates physical GPUs; CUDA_VISIBLE_DEVICES may renumber them."""
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        ids = [v.strip() for v in vis.split(",") if v.strip()]
        if cuda_index < len(ids) and ids[cuda_index].isdigit():
            return int(ids[cuda_index])
    return cuda_index


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference_arm(args)
    # stdout carries exactly one JSON line: anything libraries print there meanwhile (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import importlib
    import litejpeg_b200 as lj
    synth = importlib.import_module("litejpeg_b200.synth")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the encode path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    all_cpus = bind_to_gpu_numa_node(_nvml_index(torch, local_rank))   # before the pinned buffers are allocated
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")      # host-side waits that must not occupy the GPUs
    n_gpus = max(world, 1)
    total_frames = args.frames if args.scaling == "strong" else args.frames * n_gpus
    lo, hi = shard(total_frames, n_gpus, rank)
    frames = hi - lo
    if frames < 1:
        raise RuntimeError("fewer frames (%d) than GPUs (%d)" % (total_frames, n_gpus))

    # ---- synthetic input, pinned on the host: frame i of the batch is synthetic frame i % unique.  Rank 0
    # holds the whole batch (the one-call end-to-end path starts from it), the others their shard.
    uniq = min(UNIQUE_FRAMES, total_frames)
    h_lo, h_hi = (0, total_frames) if rank == 0 else (lo, hi)
    host_all = torch.empty((h_hi - h_lo, H, W, 3), dtype=torch.uint8).pin_memory()
    made = {}
    for i in range(h_lo, h_hi):
        if i % uniq not in made:
            host_all[i - h_lo] = synth.frame(H, W, CONFIG_IDX, i % uniq, KIND)
            made[i % uniq] = i - h_lo
        else:
            host_all[i - h_lo] = host_all[made[i % uniq]]
    host = host_all[lo - h_lo:hi - h_lo]
    rgb = host.to(dev, non_blocking=True)
    torch.cuda.synchronize()

    L = lj.lib()
    desc = lj.make_desc(H, W, SUB, frames)
    geo = lj.geometry(desc)
    qy, qc = lj.quant_tables(QUALITY)
    qy_a, qc_a = lj._lib.table_arg(qy), lj._lib.table_arg(qc)
    zrl = lj._lib.ZRL[ZRL]
    dct_flag = lj._lib.DCT_MODE[args.dct]
    ent_opts = zrl | lj._lib.opt_restart(args.restart)
    cap = int(L.b2j_default_out_cap(ctypes.byref(desc)))
    nblk = frames * geo.blocks_per_frame
    coefs = torch.empty((nblk, 64), dtype=torch.int16, device=dev)
    ent_ws = torch.empty((L.b2j_entropy_workspace_bytes_opts(ctypes.byref(desc), cap, lj._lib.opt_restart(args.restart)) + 256,),
                         dtype=torch.uint8, device=dev)
    ent_ptr = (ent_ws.data_ptr() + 255) & ~255
    out = torch.empty((frames, cap), dtype=torch.uint8, device=dev)
    sizes = torch.empty((frames,), dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    dref = ctypes.byref(desc)

    if frames * H * W * 3 > 126e6 and not args.no_graph:
        args.graph = True                                       # input larger than the L2: graph replay of the step
    flush = None
    if frames * H * W * 3 <= 126e6 and not args.graph:          # input fits L2: flush it between the timed steps
        flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)

    def step(ev=None):
        """One pass of the hot path: exactly the launches b2j_encode issues, with optional stage events."""
        nonlocal stream, sp
        if flush is not None: flush.fill_(1)
        if ev: ev[0].record(stream)
        lj._lib.check(L.b2j_fused_block(rgb.data_ptr(), dref, qy_a, qc_a, coefs.data_ptr(), dct_flag, sp))
        if ev: ev[1].record(stream)
        lj._lib.check(L.b2j_entropy_encode(coefs.data_ptr(), dref, qy_a, qc_a, ent_opts, ent_ptr, out.data_ptr(), cap,
                                           sizes.data_ptr(), sp))
        if ev: ev[2].record(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness guard: the staged sequence equals the one-call path, and frame 0 equals the oracle
    step()
    torch.cuda.synchronize()
    sz = sizes.cpu()
    assert int(sz.min()) > 0, "a frame did not fit its output slot"
    enc = lj.JpegEncoder(SUB, quality=QUALITY, zrl=ZRL, dct=args.dct, restart_interval=args.restart)
    out2, sz2 = enc.encode_device(rgb)
    torch.cuda.synchronize()
    assert torch.equal(sz2.cpu(), sz) and torch.equal(out2[0, :int(sz[0])], out[0, :int(sz[0])])
    del out2, sz2
    enc._ws = {}
    torch.cuda.empty_cache()
    out_bytes = int(sz.sum())

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(_nvml_index(torch, local_rank if world > 1 else torch.cuda.current_device())) if rank == 0 else None
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    graph = None
    if args.graph:
        for k in range(args.steps):                 # stage split from an ungraphed pass (not part of `value`)
            step(evs[k])
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            sp_saved, stream_saved = sp, stream
            stream = side
            sp = side.cuda_stream
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                step()
            stream, sp = stream_saved, sp_saved
        for _ in range(args.warmup):
            graph.replay()
    barrier()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    t_start.record(stream)
    for k in range(args.steps):
        if graph is not None:
            graph.replay()
        else:
            step(evs[k])
    t_end.record(stream)
    barrier()
    if sampler:
        sampler.window(w0, time.perf_counter())
    stage_ms = [sum(e[i].elapsed_time(e[i + 1]) for e in evs) / args.steps for i in range(2)]
    # the step's own time: K1 start to stuffing end (an L2 flush between steps, when there is one, is not part of it)
    elapsed_ms = t_start.elapsed_time(t_end) if flush is None else sum(e[0].elapsed_time(e[2]) for e in evs)
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    px_per_step = total_frames * H * W
    value = px_per_step * args.steps / (elapsed_ms / 1e3) / 1e6

    # ---- end to end through the host-buffer entry point of the C ABI: ONE call on rank 0 over the devices of
    # all ranks (b2j_multi_encoder_encode_host), from the pinned host batch to JFIF bytes in pinned host memory.
    # The other ranks wait on a host-side (gloo) barrier, so their GPUs are free for that call.
    def host_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=host_group)

    e2e = None
    if not args.skip_e2e:
        host_barrier()
        if rank == 0:
            fpc = max(1, 16 * 1080 * 1920 // (H * W))
            devices = list(range(n_gpus)) if torch.cuda.device_count() >= n_gpus else [local_rank]
            menc = lj.MultiGpuJpegEncoder(devices, H, W, SUB, quality=QUALITY, zrl=ZRL, dct=args.dct, restart_interval=args.restart,
                                          out_cap=cap, frames_per_chunk=fpc)
            out_host = torch.empty((total_frames, cap), dtype=torch.uint8).pin_memory()
            sizes_host = torch.empty((total_frames,), dtype=torch.int64)
            for _ in range(2):
                menc.encode_host(host_all, out_host, sizes_host)
            assert torch.equal(sizes_host[lo:hi], sz) and int(sizes_host.min()) > 0
            assert out_host[hi - 1, :int(sz[-1])].equal(out[frames - 1, :int(sz[-1])].cpu())
            e_steps = max(3, min(args.steps, 10))
            t0 = time.perf_counter()
            for _ in range(e_steps):
                menc.encode_host(host_all, out_host, sizes_host)
            dt = time.perf_counter() - t0
            if sampler:
                sampler.window(t0, t0 + dt)
            h2d_s = menc.h2d_seconds(host_all, repeats=3)            # only the host-to-device copies of the same call
            menc.close()
            h2d_bytes = total_frames * H * W * 3
            all_out = int(sizes_host.sum())
            frac = h2d_s / (dt / e_steps)
            e2e = {"value": total_frames * H * W * e_steps / dt / 1e6, "unit": UNIT,
                   "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": all_out + total_frames * 8,
                   "steps": e_steps, "ms_per_step": dt / e_steps * 1e3, "frames_per_chunk": fpc, "devices": devices,
                   "h2d_ceiling_gbs": h2d_bytes / h2d_s / 1e9, "h2d_only_ms": h2d_s * 1e3, "frac_of_h2d_ceiling": frac,
                   "limiter": ("host-to-device copies (PCIe / host memory path): the call runs at %.0f %% of the time the copies "
                               "alone take" % (100 * frac)) if frac >= 0.7 else
                              "not the copies alone (%.0f %% of the call): kernels, device-to-host copies and host threads" % (100 * frac),
                   "api": "b2j_multi_encoder_encode_host on rank 0 over %d device(s): one call, pinned host RGB of the whole batch "
                          "in, JFIF bytes in pinned host memory out; wall clock around the call, which returns when the last "
                          "byte is on the host" % len(devices)}
        host_barrier()
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (K1)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    padded_px = frames * geo.padded_width * geo.padded_height
    coef_per_px = {"444": 3.0, "422": 2.0, "420": 1.5}[SUB]
    k1_bytes = int(padded_px * (3 + 2 * coef_per_px))   # 3 B/px RGB read + int16 coefficients written (9 / 7 / 6 B per padded px)
    k1_ms = stage_ms[0]
    achieved = k1_bytes / (k1_ms / 1e3) / 1e9
    # dram bytes of one K1 launch from the ncu --set full capture of this workload (profiles/scripts/prof_full.sh writes
    # the file together with a hash of the kernel sources; a capture of other sources is not reported)
    traffic, traffic_note = None, None
    if args.config == "2" and KIND == "photo" and frames == FRAMES_PER_GPU and args.dct == "model":   # the capture's workload
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
            if traffic.get("k1_source_sha") != k1_source_sha():
                traffic_note = "profiles/k1_traffic.json was captured from other kernel sources (%s)" % traffic.get("k1_source_sha")
                traffic = None
        except Exception:
            pass
    hv = {"444": "1,1", "422": "2,1", "420": "2,2"}[SUB]
    pair = args.dct == "model" and min(int(x) for x in list(qy) + list(qc)) >= 2        # b2j.h B2J_FUSED_PAIR: what AUTO launches
    roofline = {"kernel": "%s<%s> (K1: colour + chroma resampling + %s DCT + quantise + zig-zag)"
                          % ("k_fused_pair" if pair else "k_fused_stream", hv,
                             "float32-with-error-interval / exact float64" if pair else "float64" if args.dct == "model" else "fixed-point"),
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "frac_of_nominal_8000": achieved / 8000.0,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_launch": k1_bytes, "ms_per_launch": k1_ms,
                "traffic": traffic["bytes_per_launch"] if traffic else None,
                "traffic_source": traffic.get("source") if traffic else traffic_note,
                "k1_source_sha": k1_source_sha(),
                "share_of_step": k1_ms / sum(stage_ms),
                "note": "K1 is issue-bound, not HBM-bound: ~58 thread-instructions per pixel for the bit-exact colour "
                        "conversion, DCT and quantiser (DESIGN.md, profiles/r02)"}

    base = None
    if all_cpus:
        os.sched_setaffinity(0, all_cpus)            # the CPU leg uses every host thread
    if not args.skip_cpu_baseline and n_gpus == 1 and args.dct == "model" and not args.restart:        # reported on rank 0 at N = 1 only
        base, _ = cpu_baseline(host_all, 2)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(n_gpus, args.frames, args.scaling),
            "stage_ms": {"K1_fused_block": stage_ms[0], "K2_K3_K4_entropy_fused_and_stuffing": stage_ms[1]},
            "timing": ("the %d timed steps are replays of a CUDA graph of one step (the same launches); stage_ms and the "
                       "roofline come from an eager pass with events between the stages" % args.steps) if graph is not None
                      else "eager launches, CUDA events around the timed steps",
            "output_bytes_per_step_rank0": out_bytes, "bits_per_pixel": out_bytes * 8 / (frames * H * W),
            "roofline": roofline, "cpu_baseline": base, "e2e": e2e,
            "gpu_launches": args.steps * int(L.b2j_encode_launch_count(dref)), "clocks": clocks}
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
