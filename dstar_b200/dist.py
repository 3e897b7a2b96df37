"""Multi-GPU plumbing: stars are independent, so ranks only agree on who owns which stars and gather
the light curves at the end (torch.distributed; NCCL on GPUs, gloo in the CPU tests).  No data-path
collective exists or is needed."""
import numpy as np


def shard_range(n_total, rank, world):
    """Contiguous block of star indices owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allmax(dist, x, device="cpu"):
    """max over ranks of a scalar (device timings are reported as the max over ranks)."""
    if dist is None:
        return float(x)
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_monitors(dist, local, n_total, rank, world, device="cpu"):
    """Final gather of per-rank (n_epoch, n_local) monitor blocks into (n_epoch, n_total) on rank 0."""
    local = np.ascontiguousarray(local, dtype=np.float64)
    if dist is None:
        return local
    import torch
    n_epoch = local.shape[0]
    sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    pad = max(sizes)
    buf = torch.zeros((n_epoch, pad), dtype=torch.float64, device=device)
    buf[:, :local.shape[1]] = torch.from_numpy(local).to(device)
    out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0)
    if rank != 0:
        return None
    return np.concatenate([o[:, :s].cpu().numpy() for o, s in zip(out, sizes)], axis=1)
