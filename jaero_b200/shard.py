"""Multi-GPU plumbing: channels are independent (no cross-channel term anywhere in writeData),
so the path shards by contiguous channel ranges, one process per GPU, with NO collective in the data path.
torch.distributed (NCCL over NVLink on the GPU box, gloo in CPU tests) is used only to
  * broadcast a shared source waveform / base envelopes from rank 0 (ncclBroadcast), and
  * reduce / gather per-rank counters and digests for reporting (max-over-ranks timing included).
"""
import os

import torch
import torch.distributed as dist


def channel_range(n_channels, world, rank):
    """Contiguous slice [lo, hi) of rank `rank`: [g*N/G, (g+1)*N/G)."""
    lo = (rank * n_channels) // world
    hi = ((rank + 1) * n_channels) // world
    return lo, hi


def weighted_ranges(costs, world):
    """Contiguous ranges balancing the summed per-channel cost (mixed-mode batches, SURVEY.md §8d cfg 5)."""
    total = float(sum(costs))
    bounds, acc, g = [0], 0.0, 1
    for i, c in enumerate(costs):
        acc += c
        while g < world and acc >= total * g / world:
            bounds.append(i + 1)
            g += 1
    while len(bounds) < world + 1:
        bounds.append(len(costs))
    bounds[-1] = len(costs)
    return [(bounds[i], bounds[i + 1]) for i in range(world)]


def init_from_env(backend=None):
    """Join the torchrun-provided process group (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def broadcast_(t, src=0):
    """In-place broadcast of a shared waveform from rank `src` (no-op for a single process)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(t, src=src)
    return t


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def reduce_max(value, device="cpu"):
    """Max over ranks of a scalar (device-timed milliseconds are combined this way)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(values, device="cpu"):
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]


def gather_rows(row, device="cpu"):
    """all_gather a small 1-D int64 digest row per rank -> [world, len]."""
    t = torch.as_tensor(row, dtype=torch.int64, device=device).reshape(1, -1)
    if dist.is_initialized() and dist.get_world_size() > 1:
        out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t)
        return torch.cat(out, 0)
    return t
