"""Multi-GPU plumbing of the encode path: frames are independent (`-g 1`, every frame a keyframe;
/root/reference/Source/CLI/Global.cpp:959-960), so they shard one contiguous run per rank with no data-path collective.
The only exchange step is the gather of the encoded packets to rank 0, which owns the Matroska muxer: per-frame lengths
first (all_gather), then every rank's packet arena (send/recv, variable length). Backend-agnostic: NCCL on device tensors
in production, gloo on CPU tensors in the tests."""
import torch
import torch.distributed as dist


def shard_frames(n_frames, rank, world):
    """Contiguous run [start, stop) of frame indices for `rank` (sequential file reads per rank, frame order kept)."""
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_packets(arena, lens, rank, world, max_frames_per_rank):
    """arena: uint8 tensor holding this rank's packets back to back; lens: int64 tensor [n_local] of packet sizes.
    Returns on rank 0 a list over ranks of (arena, lens) in rank order (= frame order with shard_frames); None elsewhere."""
    dev = arena.device
    if world == 1:
        return [(arena, lens)]
    n_local = torch.tensor([lens.numel()], dtype=torch.int64, device=dev)
    padded = torch.zeros(max_frames_per_rank, dtype=torch.int64, device=dev)
    padded[:lens.numel()] = lens
    all_n = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    all_lens = [torch.zeros(max_frames_per_rank, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_n, n_local)
    dist.all_gather(all_lens, padded)
    if rank != 0:
        total = int(lens.sum().item())
        if total:
            for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, arena[:total].contiguous(), 0)]):
                q.wait()
        return None
    out = [(arena[:int(lens.sum().item())], lens)]
    bufs, ops = [], []
    for r in range(1, world):
        ln = all_lens[r][:int(all_n[r].item())]
        buf = torch.empty(int(ln.sum().item()), dtype=torch.uint8, device=dev)
        bufs.append((buf, ln))
        if buf.numel():
            ops.append(dist.P2POp(dist.irecv, buf, r))
    if ops:
        for q in dist.batch_isend_irecv(ops):      # one grouped NCCL launch (ncclGroupStart/End) for all the senders
            q.wait()
    return out + bufs


def split_packets(arena, lens):
    """arena + lens -> list of per-frame uint8 tensors (views)."""
    out, o = [], 0
    for n in lens.tolist():
        out.append(arena[o:o + n])
        o += n
    return out
