#!/usr/bin/env python
"""bench.py — FFV1 (+FLAC) encode throughput of the B200 path on BASELINE.json's workloads.

Default workload = BASELINE config 3, the one the metric is quoted on (it fits one GPU): UHD 4K (3840x2160) 16-bit RGB DPX
(big endian), `-slices 24` (6x4), `-context 1 -coder 1 -slicecrc 1 -level 3 -g 1` (RAWcooked's option set,
/root/reference/Source/CLI/Global.cpp:938-989). `--config 2|4|5` runs the other BASELINE configurations with the same JSON
shape (2: 2K 10-bit, 4 slices; 4: 4K 16-bit TIFF + 6-channel 96 kHz 24-bit FLAC together; 5: 8K 12-bit, 64 slices).
One "step" = one batch of --frames synthetic frames per GPU through the whole hot path (k_model -> k_range -> k_emit ->
k_scan/k_pack). Frames shard one batch per GPU (weak scaling); with N > 1 the packets are gathered to rank 0 over NCCL
inside the timed region (the path's only exchange step), the gather of step i running behind the encode of
step i+1.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--config C] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0). After the timed loops three packets of the LAST timed batch (first, middle, last frame) are
checked: byte-identical to the oracle and decoded by the unmodified reference decoder back to the input payload; the line
carries "check": "pass" and the process exits non-zero otherwise. `--impl reference` times the reference's own CPU
implementation of the path — FFmpeg's ffv1 encoder (libavcodec 62.11.100 from this image, the encoder RAWcooked shells
out to), all host cores — on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import mmap
import os
import subprocess
import sys
import threading
import time
import zlib

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs, numbered as in BASELINE.json's list counted from 1 (config 1 is the CPU plumbing case, no bench line)
CONFIGS = {
    2: dict(W=2048, H=1556, layout=2, slices=4, bits=10, noise=4, frames=128, fmt="gbrp10le", audio=False,
            name="BASELINE config 2: 2K 2048x1556 10-bit RGB DPX (Filled-A, big endian), FFV1 v3, -slices 4 (2x2)"),
    3: dict(W=3840, H=2160, layout=7, slices=24, bits=16, noise=256, frames=128, fmt="gbrp16le", audio=False,
            name="BASELINE config 3: UHD 4K 3840x2160 16-bit RGB DPX (big endian), FFV1 v3, -slices 24 (6x4)"),
    4: dict(W=3840, H=2160, layout=33, slices=24, bits=16, noise=256, frames=128, fmt="gbrp16le", audio=True,
            name="BASELINE config 4: UHD 4K 3840x2160 16-bit RGB TIFF (little endian), FFV1 v3, -slices 24 (6x4) + "
                 "24-bit 96 kHz 6-channel WAV -> FLAC, together"),
    5: dict(W=7680, H=4320, layout=5, slices=64, bits=12, noise=16, frames=32, fmt="gbrp12le", audio=False,
            name="BASELINE config 5: 8K 7680x4320 12-bit RGB DPX (Filled-A, big endian), FFV1 v3, -slices 64 (8x8)"),
}
AUDIO = dict(rate=96000, channels=6, bits=24, fps=24)


def metric_name(c):
    if c["W"] == 3840 and not c["audio"]:
        return "MPix/s FFV1 encode, 4K 16-bit RGB DPX (3840x2160, 24 slices)"
    return "MPix/s FFV1 encode, %dx%d %d-bit RGB, %d slices%s" % (c["W"], c["H"], c["bits"], c["slices"], " + FLAC" if c["audio"] else "")


def frame_bytes(c):
    return c["W"] * c["H"] * (4 if c["bits"] == 10 else 6)


def workload_config(c, frames, n_gpus):
    return {
        "workload": c["name"] + ", -context 1 -coder 1 -slicecrc 1 -level 3 -g 1; ramps + uniform noise +-%d (film-grain model), seeded" % c["noise"],
        "frames_per_step_per_gpu": frames,
        "frame_bytes": frame_bytes(c),
        "sharding": "frame-parallel, one batch per GPU; packets gathered to rank 0 over NCCL, behind the next step's encode" if n_gpus > 1 else "single GPU",
        "l2": "inputs of one step (%.1f GB) are larger than L2 (126 MB); no flush needed" % (frames * frame_bytes(c) / 1e9),
    }


# ---------------------------------------------------------------------------------------------------------------------
def synth_frames_torch(c, n, seed, device):
    """n payloads in the file layout of config c, built on the GPU: per-channel ramps + uniform noise. uint8 [n, frame_bytes]."""
    import torch
    W, H, bits, layout, amp = c["W"], c["H"], c["bits"], c["layout"], c["noise"]
    mx = (1 << bits) - 1
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    yy = torch.arange(H, device=device, dtype=torch.int64).view(H, 1)
    xx = torch.arange(W, device=device, dtype=torch.int64).view(1, W)
    out = torch.empty((n, H * W * (4 if bits == 10 else 6)), dtype=torch.uint8, device=device)
    for i in range(n):
        comp = []
        for k in range(3):
            ramp = (xx * (k + 1) * mx // (3 * (W - 1)) + yy * (3 - k) * mx // (4 * (H - 1)) + 977 * i) % (mx + 1)
            noise = torch.randint(-amp, amp + 1, (H, W), generator=g, device=device, dtype=torch.int64)
            comp.append((ramp + noise).clamp_(0, mx))
        if bits == 10:                                   # R<<22 | G<<12 | B<<2, big endian
            v = (comp[0] << 22) | (comp[1] << 12) | (comp[2] << 2)
            o = out[i].view(H, W, 4)
            for b in range(4):
                o[:, :, b] = ((v >> (8 * (3 - b))) & 255).to(torch.uint8)
        else:
            sh = 4 if bits == 12 else 0
            le = layout == 33
            o = out[i].view(H, W, 3, 2)
            for k in range(3):
                v = comp[k] << sh
                o[:, :, k, 1 if le else 0] = (v >> 8).to(torch.uint8)
                o[:, :, k, 0 if le else 1] = (v & 255).to(torch.uint8)
    return out


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
class CpuReference:
    """FFmpeg's ffv1 encoder (libavcodec) on host cores, frame-parallel: one encoder context (threads=1) per worker so
    that every core is busy (slice threading alone cannot use more cores than slices). Content and encoder contexts are
    built ONCE; run(n) then encodes n frames. Falls back to the C port of the oracle when libavcodec is missing."""

    def __init__(self, c, seed=4242, max_workers=None):
        import numpy as np  # noqa: F401
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        from rawcooked_b200 import synth as S
        self.c = c
        cores = os.cpu_count() or 1
        try:
            cores = min(cores, len(os.sched_getaffinity(0)))
        except AttributeError:
            pass
        self.workers = max(1, min(cores, max_workers or 64))
        R, G, B = S.rgb_content(c["W"], c["H"], c["bits"], seed, "grain", noise=c["noise"])
        try:
            import avcodec_ffv1 as A
            self.A = A
            self.version = A.version()
            self.kind = "reference"
        except Exception as e:          # noqa: BLE001
            self.kind = "port"
            self.err = str(e)
        if self.kind == "reference":
            planes = [G, B, R]                       # gbrp* plane order
            self.encs = [self.A.FFV1Encoder(c["W"], c["H"], c["fmt"], c["slices"], threads=1) for _ in range(self.workers)]
            for e in self.encs:
                self.packet = len(e.encode_planes(planes))      # fills the frame buffer once + warm-up encode
        else:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import util
            self.util = util
            self.payload = S.pack_payload(R, G, B, c["layout"])
            self.workers = 1

    def run(self, n_frames):
        """Encodes n_frames (spread over the workers); returns seconds."""
        c = self.c
        if self.kind == "port":
            t0 = time.perf_counter()
            for _ in range(n_frames):
                nh, nv = {4: (2, 2), 24: (6, 4), 64: (8, 8)}[c["slices"]]
                self.util.oracle_encode(self.payload, c["W"], c["H"], c["layout"], nh, nv)
            return time.perf_counter() - t0
        per = [n_frames // self.workers + (1 if i < n_frames % self.workers else 0) for i in range(self.workers)]

        def work(i):
            for _ in range(per[i]):
                self.encs[i].encode_current()
        ths = [threading.Thread(target=work, args=(i,)) for i in range(self.workers) if per[i]]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        return time.perf_counter() - t0

    def describe(self, n_frames, seconds):
        c = self.c
        if self.kind == "port":
            return "%d frame(s) through oracle/ffv1_oracle.c (single thread); libavcodec unavailable: %s" % (n_frames, self.err)
        return ("%d frames of the workload (one seeded %dx%d %d-bit frame re-encoded; intra-only, no state between frames), libavcodec %s ffv1 "
                "coder=1 context=1 g=1 level=3 slicecrc=1 slices=%d, %d encoder contexts x 1 thread, frame already in RAM, packet discarded; "
                "%.1f s; packet %d bytes" % (n_frames, c["W"], c["H"], c["bits"], self.version, c["slices"], self.workers, seconds, self.packet))

    def close(self):
        if self.kind == "reference":
            for e in self.encs:
                e.close()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = CONFIGS[args.config]
    t_wall = time.perf_counter()
    ref = CpuReference(c)
    # calibrate: one frame per worker
    t_cal = ref.run(ref.workers)
    fps_est = ref.workers / max(t_cal, 1e-3)
    total_steps = args.steps + args.warmup
    budget = 150.0                                       # seconds for all steps together
    F = args.frames
    sample = int(min(F, max(ref.workers, budget / total_steps * fps_est)))
    sample = max(ref.workers, sample - sample % ref.workers)
    secs = []
    for i in range(total_steps):
        dt = ref.run(sample)
        if i >= args.warmup:
            secs.append(dt)
    dt = sum(secs) / len(secs)
    v = sample * c["W"] * c["H"] / dt / 1e6
    desc = ref.describe(sample, dt)
    if sample != F:
        desc += "; each step = a %d-frame sample of the %d-frame step, ms_per_step scaled by %d/%d" % (sample, F, F, sample)
    line = {
        "impl": "reference", "metric": metric_name(c), "value": v, "unit": "MPix/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt * F / sample, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(c, F, args.gpus),
        "cpu_baseline": {"value": v, "unit": "MPix/s", "cores": ref.workers, "kind": ref.kind, "sample": desc},
        "e2e": {"value": v, "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "fps": v * 1e6 / (c["W"] * c["H"]), "sample_frames_per_step": sample,
    }
    ref.close()
    line["wall_s"] = time.perf_counter() - t_wall
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
class _CudaBuf:
    """Exposes a raw device pointer to torch (zero-copy) through __cuda_array_interface__."""
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}


def check_packets(c, enc, payloads, packets):
    """oracle equality + reference decode of a few packets (tests/util.py drives oracle/: the checker, not the product)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    nh, nv = enc.grid
    detail = []
    ok = True
    for idx, (pay, pkt) in enumerate(zip(payloads, packets)):
        want = util.oracle_encode(pay, c["W"], c["H"], c["layout"], nh, nv)
        same = bytes(pkt) == want
        dec = None
        if util.ref_available():
            dec = util.ref_decode(enc.config_record, bytes(pkt), c["W"], c["H"], c["layout"], threads=8) == pay.tobytes()
        ok = ok and same and (dec is not False)
        detail.append({"packet_bytes": len(pkt), "equals_oracle": same, "reference_decode_equals_input": dec})
    return ok, detail


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from rawcooked_b200 import dist as D, ffv1

    c = CONFIGS[args.config]
    W, H = c["W"], c["H"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    F = args.frames if args.frames else c["frames"]
    layout = c["layout"]
    fb = frame_bytes(c)

    enc = ffv1.FFV1Encoder(W, H, layout, slices=c["slices"], max_frames=F, device=local)
    assert enc.frame_bytes == fb
    d_frames = synth_frames_torch(c, F, 1000 + rank, dev)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()

    # ---- FLAC leg of config 4: F / 24 s of 6-channel 96 kHz 24-bit audio per step, encoded inside the timed region
    flac_enc = wav = None
    flac_stats = None
    if c["audio"]:
        from rawcooked_b200 import flac as FL, synth as S
        nsamp = F * AUDIO["rate"] // AUDIO["fps"]
        pcm = S.wav_pcm(AUDIO["channels"], AUDIO["rate"], AUDIO["bits"], nsamp, seed=77)
        wav = FL.pcm_to_wav_bytes(pcm, AUDIO["bits"])
        flac_enc = FL.FLACEncoder(AUDIO["rate"], AUDIO["channels"], AUDIO["bits"], max_blocks=256, device=local)

    def flac_step():
        if flac_enc is None:
            return 0
        frames = flac_enc.encode(wav)
        return sum(len(f) for f in frames)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # the packets of step i travel to rank 0 (NCCL over NVLink) on a side stream while step i+1 is being encoded: the arena
    # of the step is first copied to one of two staging buffers (a device-to-device copy of ~4 GB)
    side = torch.cuda.Stream(device=dev) if world > 1 else None
    staging = [None, None]
    staged_ev = [None, None]
    gathered = [0]

    def gather_async(k):
        arena, off, ln = enc.packets_device(F)               # waits for the encode of this step
        total = off[-1] + ln[-1]
        b = k & 1
        if staged_ev[b] is not None:
            staged_ev[b].synchronize()                        # the gather that last used this buffer is done
        if staging[b] is None or staging[b].numel() < total:
            staging[b] = torch.empty(int(total * 1.1) + 4096, dtype=torch.uint8, device=dev)
        staging[b][:total].copy_(torch.as_tensor(_CudaBuf(arena, total), device=dev), non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(stream)
        lens = torch.tensor(ln, dtype=torch.int64, device=dev)
        with torch.cuda.stream(side):
            side.wait_event(ready)
            got = D.gather_packets(staging[b][:total], lens, rank, world, F)
            if got is not None:
                gathered[0] = sum(int(a.numel()) for a, _ in got)
            ev = torch.cuda.Event()
            ev.record(side)
            staged_ev[b] = ev

    def step_device(k):
        enc.encode_device(d_frames.data_ptr(), F, stream.cuda_stream)
        if flac_enc is not None:
            flac_step()
        if world > 1:
            gather_async(k)

    # ---- device-resident throughput (`value`)
    for k in range(args.warmup):
        step_device(k)
    if side is not None:
        side.synchronize()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(args.steps):
        step_device(k)
    if side is not None:
        stream.wait_stream(side)                              # the last gather ends inside the timed region
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
    dev_total = off[-1] + ln[-1]
    dev_crc = zlib.crc32(torch.as_tensor(_CudaBuf(arena, dev_total), device=dev).cpu().numpy().tobytes())

    # ---- end to end through the host-buffer C-ABI call: pinned host frames -> H2D -> encode -> D2H packets.
    # N > 1: every rank's packets land in ONE host segment that rank 0 (the muxer's process) has mapped: each GPU writes its
    # share over its own PCIe link (cudaHostRegister'ed POSIX shared memory), nothing funnels through rank 0's link.
    h_frames = torch.empty((F, fb), dtype=torch.uint8, pin_memory=True)
    h_frames.copy_(d_frames)
    cap = int(out_bytes * 1.05) + (1 << 20)
    shm_path = "/dev/shm/b200_bench_%s" % os.environ.get("MASTER_PORT", str(os.getpid()))
    if world > 1:
        caps = torch.tensor([cap], dtype=torch.int64, device=dev)
        dist.all_reduce(caps, op=dist.ReduceOp.MAX)
        cap = (int(caps.item()) + 4095) & ~4095
        if rank == 0:
            with open(shm_path, "wb") as f:
                f.truncate(cap * world)
        dist.barrier()
        fd = os.open(shm_path, os.O_RDWR)
        mm = mmap.mmap(fd, cap * world)
        seg = np.frombuffer(mm, dtype=np.uint8)
        rc = torch.cuda.cudart().cudaHostRegister(seg.ctypes.data + rank * cap, cap, 0)
        if int(rc) != 0:
            raise SystemExit("bench.py: cudaHostRegister of the shared packet segment failed (%s)" % str(rc))
        out_np = seg[rank * cap:(rank + 1) * cap]
    else:
        h_out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
        out_np = h_out.numpy()
    h_np = h_frames.numpy()
    L = ffv1.load_library()
    ptrs = (C.c_void_p * F)(*[h_np[i].ctypes.data for i in range(F)])
    offs = (C.c_size_t * F)()
    lens = (C.c_size_t * F)()

    def ck(rc):
        if rc:
            raise RuntimeError(L.b200_last_error().decode())

    flac_bytes = [0]

    def run_e2e(steps):
        """`steps` batches through the host entry points, two in flight: while batch i is coded, the frames of batch i+1 go up
        (b200_ffv1_prefetch_host / b200_ffv1_submit_host) and the packets of batch i-1 come down (b200_ffv1_fetch_packets), so
        both PCIe directions hide behind the kernels; only the first upload and the last download of a run have nothing to hide behind."""
        submitted = fetched = 0
        for _ in range(min(2, steps)):
            ck(L.b200_ffv1_submit_host(enc._h, ptrs, F))
            submitted += 1
        while fetched < steps:
            if submitted < steps:
                # the upload of the next batch starts before the download of the oldest one: both directions of the link are
                # busy while the batch in between is coded
                ck(L.b200_ffv1_prefetch_host(enc._h, ptrs, F))
            flac_bytes[0] = flac_step()
            ck(L.b200_ffv1_fetch_packets(enc._h, out_np.ctypes.data, out_np.size, offs, lens, F))
            fetched += 1
            if submitted < steps:
                ck(L.b200_ffv1_submit_host(enc._h, ptrs, F))
                submitted += 1
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
    e2e_total = int(offs[F - 1] + lens[F - 1])
    e2e_crc = zlib.crc32(out_np[:e2e_total].tobytes())

    # rank 0 reads every rank's packets out of the shared host segment: the proof that the e2e result is in its memory
    gather_ok = True
    if world > 1:
        mine = torch.tensor([e2e_total, e2e_crc], dtype=torch.int64, device=dev)
        allv = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(allv, mine)
        if rank == 0:
            for r in range(world):
                n, crc = int(allv[r][0].item()), int(allv[r][1].item())
                gather_ok = gather_ok and zlib.crc32(seg[r * cap:r * cap + n].tobytes()) == crc
            gather_ok = gather_ok and gathered[0] > 0

    # ---- check of what was timed: first / middle / last packet of the last e2e batch (host path through the C ABI) against
    # the oracle and the reference decoder, and the device-resident batch byte-identical to it
    check, check_detail = "skipped", None
    if rank == 0 and not args.no_check:
        idx = sorted(set([0, F // 2, F - 1]))[: max(1, args.check_frames)]
        pays = [h_np[i] for i in idx]
        pkts = [out_np[offs[i]:offs[i] + lens[i]] for i in idx]
        ok, check_detail = check_packets(c, enc, pays, pkts)
        ok = ok and dev_crc == e2e_crc and dev_total == e2e_total and gather_ok
        check = "pass" if ok else "FAIL"
        check_detail = {"frames": idx, "packets": check_detail, "device_batch_equals_host_batch": dev_crc == e2e_crc,
                        "rank0_holds_all_ranks_packets": gather_ok if world > 1 else None}

    # ---- per-kernel device time of the same workload: one serial pass, every launch bracketed by CUDA events in the library
    enc.set_timing(True)
    enc.encode_device(d_frames.data_ptr(), F, stream.cuda_stream)
    stream.synchronize()
    enc.packets_device(F)
    ts = enc.stats()
    enc.set_timing(False)
    # ---- decode + compare leg (the `--check` side): the packets of that pass, still in the encoder's arena, are
    # decoded by k_decode and compared on the GPU with the frames they were coded from; beside it the reference's own decoder
    # (oracle/_ref: ffv1_frame::Process, slice-threaded as `rawcooked --check` runs it) on one of the packets
    decode_stats = None
    if rank == 0 and world == 1 and not args.no_decode:
        from rawcooked_b200 import ffv1dec
        arena, d_offs, d_lens = enc.packets_device(F)
        dec = ffv1dec.FFV1Decoder(W, H, layout, enc.config_record, max_frames=F, device=local)
        best = None
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dec.decode_device(arena, d_offs, d_lens, d_sources=d_frames.data_ptr(), stream=stream.cuda_stream)
            mm, stt = dec.result(F)
            dtd = time.perf_counter() - t0
            best = dtd if best is None else min(best, dtd)
        ds = dec.stats()
        dec.close()
        decode_stats = {"what": "b200_ffv1_decode_device: packets in the encoder's device arena -> k_dec_index + k_decode -> compare with the device "
                                "frames (inverse RCT + byte layout + compare on the GPU)",
                        "frames": F, "ms": best * 1e3, "fps": F / best, "MPix_per_s": F * W * H / best / 1e6, "k_decode_ms": ds["decode_us"] / 1e3,
                        "mismatching_bytes": int(sum(mm)), "status_bits": int(max(stt)), "samples": int(ds["samples"]),
                        "cycles_per_sample_per_slice": ds["decode_us"] * 1e-6 * 1.965e9 / (ds["samples"] / max(1, ds["slices"]))}
        if not args.no_cpu:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import util
            if util.ref_available():
                ncores = os.cpu_count() or 1
                pk0 = np.empty(d_lens[0], np.uint8)
                torch.cuda.synchronize()
                pk0[:] = torch.as_tensor(_CudaBuf(arena + d_offs[0], d_lens[0]), device=dev).cpu().numpy()
                src0 = d_frames[0].cpu().numpy().tobytes()
                t0 = time.perf_counter()
                same = util.ref_decode(enc.config_record, pk0.tobytes(), W, H, layout, threads=ncores) == src0
                dtr = time.perf_counter() - t0
                decode_stats["cpu_reference"] = {"what": "reference decoder (oracle/_ref, unmodified ffv1_frame::Process + Transform), one frame, "
                                                         "slice threads = %d" % ncores, "s_per_frame": dtr, "fps": 1.0 / dtr,
                                                 "equals_input": bool(same)}
                decode_stats["speedup_vs_cpu_reference"] = (F / best) * dtr
        if decode_stats["mismatching_bytes"] or decode_stats["status_bits"]:
            check = "FAIL"
    # ---- analysis leg: MD5 of the F payloads as `rawcooked --hash` hashes its input files (one lane per file),
    # and the padding-bit test of `--check-padding` where the layout has padding bits; beside it the reference's md5.c on one core
    analysis_stats = None
    if rank == 0 and world == 1 and not args.no_decode:
        from rawcooked_b200 import scan
        sc = scan.Scanner(max_items=F, max_bytes=0, device=local)
        best = None
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            digs = sc.md5_device(d_frames.data_ptr(), [i * fb for i in range(F)], [fb] * F, stream=stream.cuda_stream)
            dta = time.perf_counter() - t0
            best = dta if best is None else min(best, dta)
        import hashlib
        md5_ok = digs[0] == hashlib.md5(d_frames[0].cpu().numpy().tobytes()).digest() and digs[F - 1] == hashlib.md5(d_frames[F - 1].cpu().numpy().tobytes()).digest()
        analysis_stats = {"md5": {"what": "b200_md5_device: k_md5, one lane per file, %d payloads of %d bytes resident in HBM" % (F, fb),
                                  "ms": best * 1e3, "GB_per_s": F * fb / best / 1e9, "files_per_s": F / best, "equals_hashlib": bool(md5_ok)}}
        if layout < 32:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            cntp, firstp = sc.padding_device(W, H, layout, d_frames.data_ptr(), F, stream=stream.cuda_stream)
            dtp = time.perf_counter() - t0
            stp = sc.stats()
            analysis_stats["padding"] = {"what": "b200_padding_device: the test of DPX.cpp:500-608 on the same payloads", "ms": dtp * 1e3,
                                         "kernel_ms": stp["kernel_us"] / 1e3, "bytes_read": int(stp["bytes"]),
                                         "GB_per_s": (stp["bytes"] / (stp["kernel_us"] * 1e-6) / 1e9) if stp["kernel_us"] else None,
                                         "payloads_with_nonzero_padding": int(sum(1 for v in cntp if v))}
        sc.close()
        if not args.no_cpu:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import util
            if util.ref_available():
                R = util.ref_decoder()
                R.ref_md5.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
                one = d_frames[0].cpu().numpy()
                outb = C.create_string_buffer(16)
                t0 = time.perf_counter()
                R.ref_md5(one.ctypes.data, one.size, outb)
                dtr = time.perf_counter() - t0
                analysis_stats["md5"]["cpu_reference"] = {"what": "the reference's md5.c (oracle/_ref), one core, one payload", "GB_per_s": one.size / dtr / 1e9,
                                                          "files_per_s": 1.0 / dtr, "equals_gpu": outb.raw == digs[0]}
                analysis_stats["md5"]["speedup_vs_one_core"] = (F / best) * dtr
        if not md5_ok:
            check = "FAIL"
    if flac_enc is not None and rank == 0:
        # FLAC leg alone: whole-call time (H2D + k_flac + D2H) and the kernel share, SURVEY §8d: in = samples*ch*3 B, out = packet bytes
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            nbytes = flac_step()
        torch.cuda.synchronize()
        dtf = (time.perf_counter() - t0) / reps
        flac_stats = {"in_bytes_per_step": int(wav.size), "out_bytes_per_step": int(nbytes), "ratio": nbytes / wav.size,
                      "MB_per_s": wav.size / dtf / 1e6, "x_realtime": (wav.size / (AUDIO["rate"] * AUDIO["channels"] * 3)) / dtf,
                      "block_size": flac_enc.block_size, "what": "b200_flac_encode_host: host PCM -> FLAC frames in host memory"}

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
            ref = CpuReference(c)
            t_cal = ref.run(ref.workers)
            n = max(ref.workers, int(args.cpu_seconds * ref.workers / max(t_cal, 1e-3)))
            n -= n % ref.workers
            dtc = ref.run(n)
            cpu = {"value": n * W * H / dtc / 1e6, "unit": "MPix/s", "cores": ref.workers, "kind": ref.kind, "sample": ref.describe(n, dtc)}
            ref.close()
        line = {
            "metric": metric_name(c), "value": value, "unit": "MPix/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic", "config": workload_config(c, F, world),
            "check": check, "check_detail": check_detail,
            "e2e": {"value": e2e, "unit": "MPix/s", "h2d_bytes_per_step": F * fb * world, "d2h_bytes_per_step": int(out_bytes) * world,
                    "api": "b200_ffv1_submit_host + b200_ffv1_prefetch_host + b200_ffv1_fetch_packets, two batches in flight (pinned host frames -> packets in "
                           "pinned host memory; every step's frames cross PCIe inside the timed region: upload of batch i+1 and download of batch i-1 while batch i is coded)"
                           + ("; all ranks fetch into one host segment mapped by rank 0 (each GPU over its own PCIe link)" if world > 1 else ""),
                    "fps": e2e * 1e6 / (W * H)},
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
        if flac_stats is not None:
            line["flac"] = flac_stats
        if decode_stats is not None:
            line["decode_check"] = decode_stats
        if analysis_stats is not None:
            line["analysis"] = analysis_stats
        print(json.dumps(line), flush=True)
    enc.close()
    if flac_enc is not None:
        flac_enc.close()
    if world > 1:
        dist.barrier()
        if rank == 0:
            try:
                os.unlink(shm_path)
            except OSError:
                pass
        dist.destroy_process_group()
    if rank == 0 and check == "FAIL":
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS), help="BASELINE.json configuration (counted from 1)")
    ap.add_argument("--frames", type=int, default=int(os.environ.get("B200_BENCH_FRAMES", "0")), help="frames per step per GPU (default: per config)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-check", action="store_true", help="skip the oracle / reference-decoder check of the timed packets")
    ap.add_argument("--check-frames", type=int, default=3)
    ap.add_argument("--no-decode", action="store_true", help="skip the decode + compare leg (k_decode on the packets of the last pass)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    if args.impl == "reference":
        if not args.frames:
            args.frames = CONFIGS[args.config]["frames"]
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
