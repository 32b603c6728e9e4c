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


class PeerExchange(object):
    """Exchange buffers for vs_allreduce_finalize_p2p: one torch symmetric-memory allocation per rank, mapped into every
    process of the group over NVLink.  Layout per rank: 2 * world * plen doubles (slots, double-buffered on the epoch parity)
    followed by 2 * world uint32 flags (padded to 16 doubles)."""

    def __init__(self, plen, device, group=None):
        import torch
        import torch.distributed as tdist
        import torch.distributed._symmetric_memory as symm
        self.group = tdist.group.WORLD if group is None else group
        self.world = tdist.get_world_size(self.group)
        self.rank = tdist.get_rank(self.group)
        self.plen = int(plen)
        nslots = 2 * self.world * self.plen
        self.buf = symm.empty(nslots + 16 + 2 * self.world, dtype=torch.float64, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, self.group)
        base = [int(p_) for p_ in self.handle.buffer_ptrs]
        self.peer_bufs = base
        self.peer_flags = [b_ + 8 * nslots for b_ in base]
        self.epoch = 0
        torch.cuda.synchronize(device)
        tdist.barrier(self.group)                     # every rank's buffer is zeroed before anybody publishes a flag

    def next_epoch(self):
        self.epoch += 1
        return self.epoch


_exchanges = {}


def peer_exchange(plen, device, group=None):
    """Cached PeerExchange for (plen, device); None if symmetric memory is not available (then use NCCL)."""
    key = (int(plen), str(device))
    if key not in _exchanges:
        try:
            _exchanges[key] = PeerExchange(plen, device, group)
        except Exception as exc:                      # no NVLink peer access / no symmetric-memory support
            if os.environ.get("VS_P2P_STRICT"):
                raise
            _exchanges[key] = None
            print("varsens_b200: peer-memory all-reduce unavailable (%s); using NCCL" % (exc,))
    return _exchanges[key]


def reduce_and_finalize(ctx, k, n, part, flags):
    """Sum a rank's partial-sum tensor over the group and compute the indices: one NCCL all_reduce followed by
    vs_finalize, or -- with VS_P2P=1 -- the fused peer-memory kernel (vs_allreduce_finalize_p2p) over torch symmetric
    memory.  Measured on 8 B200 (r01): NCCL 0.850 ms per step, peer-memory kernel 0.956 ms (equal at 2 GPUs), so NCCL is
    the default for now."""
    import torch
    rank, ws = world()
    ex = None
    if ws > 1 and os.environ.get("VS_P2P", "0") == "1":
        ex = peer_exchange(part.numel(), part.device)
    if ex is not None:
        return ctx.allreduce_finalize_p2p(k, 1, n, ex.world, ex.rank, ex.peer_bufs, ex.peer_flags, ex.next_epoch(), part, flags)
    allreduce_partials(part)
    return ctx.finalize(k, 1, n, part, flags)
