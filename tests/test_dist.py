"""World-size-2 test of the multi-GPU host logic on CPU (gloo): frame sharding + the gather of variable-length packets to
rank 0 (the path's only exchange step). The same code runs over NCCL in bench.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rawcooked_b200 import dist as D


def test_shard_frames_covers_everything_in_order():
    for n in (0, 1, 7, 8, 1000):
        for world in (1, 2, 3, 8):
            runs = [D.shard_frames(n, r, world) for r in range(world)]
            assert runs[0][0] == 0 and runs[-1][1] == n
            assert all(runs[i][1] == runs[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in runs]
            assert max(sizes) - min(sizes) <= 1


def fake_packet(frame):
    rng = np.random.default_rng(frame)
    return rng.integers(0, 256, 100 + 37 * (frame % 11), dtype=np.uint8).tobytes()


def _worker(rank, world, port, n_frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = D.shard_frames(n_frames, rank, world)
        pk = [fake_packet(f) for f in range(a, b)]
        arena = torch.frombuffer(bytearray(b"".join(pk)), dtype=torch.uint8) if pk else torch.empty(0, dtype=torch.uint8)
        lens = torch.tensor([len(p) for p in pk], dtype=torch.int64)
        got = D.gather_packets(arena, lens, rank, world, max_frames_per_rank=(n_frames + world - 1) // world)
        if rank == 0:
            frames = []
            for ar, ln in got:
                frames += [bytes(t.numpy().tobytes()) for t in D.split_packets(ar, ln)]
            q.put(frames == [fake_packet(f) for f in range(n_frames)])
        else:
            assert got is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [9, 2, 1])
def test_gather_packets_gloo_world2(n_frames):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
