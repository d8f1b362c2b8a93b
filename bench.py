#!/usr/bin/env python
"""bench.py — FFV1 encode throughput of the B200 path on BASELINE.json's headline workload.

Workload (BASELINE.json configs[2], the one the metric is quoted on; it fits one GPU): UHD 4K (3840x2160) 16-bit RGB DPX
(big endian), `-slices 24` (6x4), `-context 1 -coder 1 -slicecrc 1 -level 3 -g 1` (RAWcooked's option set,
/root/reference/Source/CLI/Global.cpp:938-989). One "step" = one batch of --frames synthetic frames per GPU through the
whole hot path (k_model -> k_range -> k_emit -> k_scan/k_pack). Frames shard one batch per GPU (weak scaling); with N > 1
the packets are gathered to rank 0 over NCCL inside the timed region (the path's only exchange step).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0). `--impl reference` times the reference's own CPU implementation of the path — FFmpeg's
ffv1 encoder (libavcodec 62.11.100 from this image, the encoder RAWcooked shells out to), all host cores, on a bounded
sample of the same workload.
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

W, H = 3840, 2160
SLICES = 24
METRIC = "MPix/s FFV1 encode, 4K 16-bit RGB DPX (3840x2160, 24 slices)"


def workload_config(frames, n_gpus):
    return {
        "workload": "BASELINE configs[2]: UHD 4K 3840x2160 16-bit RGB DPX (big endian), FFV1 v3, -slices 24 (6x4), "
                    "-context 1 -coder 1 -slicecrc 1 -level 3 -g 1; ramps + uniform noise +-256 (film-grain model), seeded",
        "frames_per_step_per_gpu": frames,
        "frame_bytes": W * H * 6,
        "sharding": "frame-parallel, one batch per GPU; packets gathered to rank 0 over NCCL" if n_gpus > 1 else "single GPU",
        "l2": "inputs of one step (%.1f GB) are larger than L2 (126 MB); no flush needed" % (frames * W * H * 6 / 1e9),
    }


# ---------------------------------------------------------------------------------------------------------------------
def synth_frames_torch(n, seed, device):
    """n DPX payloads (16-bit BE RGB) built on the GPU: per-channel ramps + uniform noise +-256. uint8 [n, W*H*6]."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    yy = torch.arange(H, device=device, dtype=torch.int64).view(H, 1)
    xx = torch.arange(W, device=device, dtype=torch.int64).view(1, W)
    out = torch.empty((n, H, W, 3, 2), dtype=torch.uint8, device=device)
    for i in range(n):
        for k in range(3):
            ramp = (xx * (k + 1) * 65535 // (3 * (W - 1)) + yy * (3 - k) * 65535 // (4 * (H - 1)) + 977 * i) % 65536
            noise = torch.randint(-256, 257, (H, W), generator=g, device=device, dtype=torch.int64)
            v = (ramp + noise).clamp_(0, 65535)
            out[i, :, :, k, 0] = (v >> 8).to(torch.uint8)
            out[i, :, :, k, 1] = (v & 255).to(torch.uint8)
    return out.view(n, H * W * 6)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_run(target_seconds, seed=4242, max_workers=None):
    """FFmpeg's ffv1 encoder (libavcodec) on host cores, frame-parallel: one encoder context (threads=1) per worker so
    that every core is busy (slice threading alone cannot use more cores than slices). Falls back to the C port."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from rawcooked_b200 import synth as S
    cores = os.cpu_count() or 1
    try:
        avail = len(os.sched_getaffinity(0))
        cores = min(cores, avail)
    except AttributeError:
        pass
    workers = max(1, min(cores, max_workers or 64))
    R, G, B = S.rgb_content(W, H, 16, seed, "grain")
    try:
        import avcodec_ffv1 as A
        A.version()
        kind = "reference"
    except Exception as e:          # noqa: BLE001
        kind = "port"
        err = str(e)
    if kind == "reference":
        encs = [A.FFV1Encoder(W, H, "gbrp16le", SLICES, threads=1) for _ in range(workers)]
        for e in encs:
            e.encode_planes([G, B, R])          # fills the frame buffer once + warm-up encode
        t0 = time.perf_counter()
        one = len(encs[0].encode_current())
        t_one = time.perf_counter() - t0
        per = max(1, int(target_seconds / max(t_one, 1e-3)))
        done = [0] * workers

        def work(i):
            for _ in range(per):
                encs[i].encode_current()
                done[i] += 1
        ths = [threading.Thread(target=work, args=(i,)) for i in range(workers)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        dt = time.perf_counter() - t0
        n = sum(done)
        for e in encs:
            e.close()
        sample = "%d frames of the workload (one seeded 4K16 frame re-encoded; intra-only, no state between frames), libavcodec %s ffv1 " \
                 "coder=1 context=1 g=1 level=3 slicecrc=1 slices=24, %d encoder contexts x 1 thread, frame already in RAM, packet discarded; " \
                 "%.1f s; packet %d bytes" % (n, A.version(), workers, dt, one)
    else:
        import ctypes as C
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import util
        payload = S.pack_payload(R, G, B, S.DPX_RGB_16_BE)
        t0 = time.perf_counter()
        util.oracle_encode(payload, W, H, S.DPX_RGB_16_BE, 6, 4)
        dt = time.perf_counter() - t0
        n, workers = 1, 1
        sample = "1 frame through oracle/ffv1_oracle.c (single thread); libavcodec unavailable: " + err
    mpix = n * W * H / dt / 1e6
    return {"value": mpix, "unit": "MPix/s", "cores": workers, "kind": kind, "sample": sample, "fps": n / dt, "seconds": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    per_step = max(4.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    vals = []
    base = None
    for i in range(args.warmup + args.steps):
        r = cpu_reference_run(per_step)
        if i >= args.warmup:
            vals.append(r)
        base = r
    v = sum(x["value"] for x in vals) / len(vals)
    ms = 1e3 * sum(x["seconds"] for x in vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "MPix/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(args.frames, args.gpus),
        "cpu_baseline": {"value": v, "unit": "MPix/s", "cores": base["cores"], "kind": base["kind"], "sample": base["sample"]},
        "e2e": {"value": v, "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "fps": v * 1e6 / (W * H), "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
class _CudaBuf:
    """Exposes a raw device pointer to torch (zero-copy) through __cuda_array_interface__."""
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from rawcooked_b200 import dist as D, ffv1, synth as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    F = args.frames
    layout = S.DPX_RGB_16_BE
    fb = W * H * 6

    enc = ffv1.FFV1Encoder(W, H, layout, slices=SLICES, max_frames=F, device=local)
    d_frames = synth_frames_torch(F, 1000 + rank, dev)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def gather_to_rank0():
        """NCCL exchange of the encoded packets (rawcooked_b200/dist.py): lengths first, then every rank's arena to rank 0."""
        arena, off, ln = enc.packets_device(F)
        total = off[-1] + ln[-1]
        mine = torch.as_tensor(_CudaBuf(arena, total), device=dev)
        lens = torch.tensor(ln, dtype=torch.int64, device=dev)
        got = D.gather_packets(mine, lens, rank, world, F)
        return sum(int(a.numel()) for a, _ in got) if got is not None else total

    def step_device():
        enc.encode_device(d_frames.data_ptr(), F, stream.cuda_stream)
        if world > 1:
            stream.synchronize()
            gather_to_rank0()

    # ---- device-resident throughput (`value`)
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    arena, off, ln = enc.packets_device(F)
    st = enc.stats()
    out_bytes = st["packet_bytes"]

    # ---- end to end through the host-buffer C-ABI call: pinned host frames -> H2D -> encode -> D2H packets
    h_frames = torch.empty((F, fb), dtype=torch.uint8, pin_memory=True)
    h_frames.copy_(d_frames)
    h_out = torch.empty(int(out_bytes * 1.05) + (1 << 20), dtype=torch.uint8, pin_memory=True)
    h_np = h_frames.numpy()
    out_np = h_out.numpy()
    import ctypes as C
    L = ffv1.load_library()
    ptrs = (C.c_void_p * F)(*[h_np[i].ctypes.data for i in range(F)])
    offs = (C.c_size_t * F)()
    lens = (C.c_size_t * F)()

    def ck(rc):
        if rc:
            raise RuntimeError(L.b200_last_error().decode())

    def run_e2e(steps):
        """`steps` batches through the host entry points, two in flight: batch i+1 is submitted (H2D band by band + kernels)
        before the packets of batch i are fetched (D2H), so both PCIe directions hide behind the kernels."""
        ck(L.b200_ffv1_submit_host(enc._h, ptrs, F))
        for _ in range(steps - 1):
            ck(L.b200_ffv1_submit_host(enc._h, ptrs, F))
            ck(L.b200_ffv1_fetch_packets(enc._h, out_np.ctypes.data, out_np.size, offs, lens, F))
        ck(L.b200_ffv1_fetch_packets(enc._h, out_np.ctypes.data, out_np.size, offs, lens, F))
    run_e2e(max(2, args.warmup // 2))
    barrier()
    t0 = time.perf_counter()
    run_e2e(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel device time of the same workload: one serial pass, every launch bracketed by CUDA events in the library
    enc.set_timing(True)
    enc.encode_device(d_frames.data_ptr(), F, stream.cuda_stream)
    stream.synchronize()
    enc.packets_device(F)
    ts = enc.stats()
    enc.set_timing(False)

    if rank == 0:
        pix_step = F * W * H * world
        value = pix_step * args.steps / (dev_ms / 1e3) / 1e6
        e2e = pix_step * args.steps / e2e_s / 1e6
        kern = {"k_model": ts["model_us"], "k_range": ts["range_us"], "k_emit": ts["emit_us"], "k_scan+k_pack": ts["pack_us"]}
        dom = max(("k_model", "k_range", "k_emit"), key=lambda k: kern[k])
        nb = enc.nbands
        alg_bytes_launch = (F * fb + out_bytes) / nb                 # payload read once + packet written once, per band launch
        dur_s = kern[dom] / 1e6 / nb
        achieved = alg_bytes_launch / dur_s / 1e9
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peak, peak_src = float(mp["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:          # noqa: BLE001
            pass
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
        except Exception:          # noqa: BLE001
            pass
        cpu = None
        if world == 1 and not args.no_cpu:
            c = cpu_reference_run(args.cpu_seconds)
            cpu = {"value": c["value"], "unit": "MPix/s", "cores": c["cores"], "kind": c["kind"], "sample": c["sample"]}
        line = {
            "metric": METRIC, "value": value, "unit": "MPix/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic", "config": workload_config(F, world),
            "e2e": {"value": e2e, "unit": "MPix/s", "h2d_bytes_per_step": F * fb * world, "d2h_bytes_per_step": int(out_bytes) * world,
                    "api": "b200_ffv1_submit_host + b200_ffv1_fetch_packets, two batches in flight (pinned host frames -> packets in pinned host memory; every step's frames cross PCIe inside the timed region, H2D band by band, D2H of batch i during batch i+1)", "fps": e2e * 1e6 / (W * H)},
            "gpu_launches": int(st["launches"]) * args.steps * world,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes_launch, "launch_ms": dur_s * 1e3,
                         "how": "serial pass of the same step, each launch bracketed by CUDA events on its stream; "
                                "per launch = one 16-row band of all frames",
                         "kernel_ms_per_step": {k: v / 1e3 for k, v in kern.items()},
                         "note": "latency-bound integer path (serial range-coder recurrence), far from the HBM roof by construction"},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "fps": value * 1e6 / (W * H), "bins_per_s": st["bins"] * args.steps * world / (dev_ms / 1e3),
            "bins_per_sample": st["bins"] / st["samples"], "compression_ratio": out_bytes / (F * fb),
        }
        print(json.dumps(line), flush=True)
    enc.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=int(os.environ.get("B200_BENCH_FRAMES", "128")), help="frames per step per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
