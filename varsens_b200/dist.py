"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink), base rows sharded
over ranks, ONE all-reduce of the O(k^2) partial-sum vector.  The reference's only
scale-out mechanism is file batching (varsens/saltelli.py:173-193 -> cluster/accre-submit.sh:27 ->
:415-472); this replaces it for registered functors.  Export mode needs no collective at all:
every rank writes its own contiguous window of flat rows."""
import os


def world():
    """(rank, world_size) from torch.distributed if initialised, else from the torchrun env, else (0, 1)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(total, rank, world_size):
    """Contiguous slice [lo, hi) of `total` units owned by `rank`; sizes differ by at most one."""
    total, rank, world_size = int(total), int(rank), int(world_size)
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(total, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_partials(partials, group=None):
    """In-place SUM all-reduce of a rank's partial-sum tensor (fp64; ~10 KB at k=20).  `partials` is a
    torch tensor on this rank's device (cuda -> NCCL, cpu -> gloo)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partials, op=dist.ReduceOp.SUM, group=group)
    return partials


def partials_layout(k, l=1):
    """Index helpers for the partial-sum vector (include/varsens_b200.h: vs_partials_len)."""
    m = (2 + 2 * k) * l

    def gram(p, q):
        if p > q:
            p, q = q, p
        return 4 * l + p * m - p * (p - 1) // 2 + (q - p)

    return dict(m=m, length=4 * l + m * (m + 1) // 2, S_A=0, S_B=l, Q_A=2 * l, Q_B=3 * l, gram=gram)
