"""Sharding of a batch of independent worlds over ranks (one process per GPU).

BASELINE.json north_star: "Partitioning across the 8xB200 box applies only to the batched-independent-
worlds mode ... sharded by world, no inter-GPU traffic except an NCCL allgather of per-world state for
validation".  Worlds are the only unit that shards; the step itself needs no collective.
"""
import numpy as np


def world_range(total_worlds, rank, world_size):
    """Contiguous range [first, end) of global world indices owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(total_worlds, world_size)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def perturbation(first_world, count, seed):
    """Per-world initial velocity of the Pyramid's top box, a pure function of the GLOBAL world index so
    that any sharding of the same batch simulates the same worlds."""
    out = np.zeros((count, 2), np.float32)
    for i in range(count):
        rng = np.random.default_rng([seed, first_world + i])
        out[i, 0] = np.float32(rng.uniform(-0.5, 0.5))
    return out


def world_digests(body_state):
    """One float64 per world: sum of the world's body state (c, a, v, w, xf.p) — the validation payload."""
    return np.ascontiguousarray(body_state, np.float32).astype(np.float64).sum(axis=(1, 2))


def allgather_digests(dist, digests, device=None):
    """All-gather the per-world digests of every rank (NCCL on GPUs, gloo on CPU); returns the list of
    per-rank arrays in rank order.  Ranks may own different numbers of worlds."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(digests, np.float64))
    if device is not None:
        t = t.to(device)
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(dist.get_world_size())]
    dist.all_gather(sizes, n)
    cap = int(max(int(s.item()) for s in sizes))
    padded = torch.zeros(cap, dtype=torch.float64, device=t.device)
    padded[:t.numel()] = t
    out = [torch.zeros_like(padded) for _ in range(dist.get_world_size())]
    dist.all_gather(out, padded)
    return [o[:int(s.item())].cpu().numpy() for o, s in zip(out, sizes)]
