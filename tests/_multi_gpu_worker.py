"""Worker of tests/test_gpu_multi.py: launched under torchrun with one rank per GPU (NCCL).  Exercises the multi-GPU
routes of the fused step -- the one-launch peer-memory step (vs_run_fused_p2p), the NCCL arm, the stand-alone exchange
kernel, the public Varsens API -- against the single-GPU result, checks that every rank holds the same bits, and that a
missing peer ends in VS_ERR_TIMEOUT instead of a hang."""
import os
import sys

import numpy
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import varsens_b200 as vb                                       # noqa: E402
from varsens_b200 import _cabi, dist as vdist, saltelli as vsalt   # noqa: E402

NAMES = ("E_2", "var_y", "U_j", "U_nj", "sens", "sens_t", "sens_2", "sens_2n")
A20 = [0, 0.5, 3, 9, 99, 99] + [99.0] * 14


def close(a, b, rel=1e-10, abs_=1e-12):
    a, b = numpy.asarray(a), numpy.asarray(b)
    d = numpy.abs(a - b)
    assert ((d <= abs_) | (d <= rel * numpy.abs(b))).all(), "max abs %.3e" % d.max()


def same_on_all_ranks(res, world):
    blob = numpy.concatenate([numpy.ascontiguousarray(getattr(res, nm)).ravel() for nm in NAMES])
    t = torch.from_numpy(blob).cuda()
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    for o in out:
        assert torch.equal(o.view(torch.int64), t.view(torch.int64)), "ranks hold different bits"


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = vb.Context.get(local)
    os.environ["VS_P2P_STRICT"] = "1"                           # a missing symmetric-memory mapping must fail the test, not fall back
    for k, n in ((20, (1 << 20) + 5), (6, 70000), (3, 4097)):
        a = A20[:k]
        perm = vsalt._reference_permutation(n)
        single = ctx.run_fused(k, n, perm, _cabi.OBJ_GFUNCTION, a)                       # the whole design on this GPU
        # one launch per rank: exchange + estimators inside the fused kernel
        l0 = ctx.launch_count()
        multi = vdist.fused_step(ctx, k, n, perm, _cabi.OBJ_GFUNCTION, a, 0, _cabi.IDENTITY, None, _cabi.FLAG_SECOND_ORDER)
        assert ctx.launch_count() - l0 == 1, "the multi-GPU step must be one launch per rank"
        for nm in NAMES:
            close(getattr(multi, nm), getattr(single, nm), rel=1e-11, abs_=1e-13)
        same_on_all_ranks(multi, world)
        again = vdist.fused_step(ctx, k, n, perm, _cabi.OBJ_GFUNCTION, a, 0, _cabi.IDENTITY, None, _cabi.FLAG_SECOND_ORDER)
        for nm in NAMES:
            assert (getattr(again, nm) == getattr(multi, nm)).all()                       # reproducible
        # NCCL arm (north_star's one all-reduce): partial sums -> all_reduce -> finalize, stream-ordered
        os.environ["VS_NCCL"] = "1"
        nccl = vdist.fused_step(ctx, k, n, perm, _cabi.OBJ_GFUNCTION, a, 0, _cabi.IDENTITY, None, _cabi.FLAG_SECOND_ORDER)
        os.environ.pop("VS_NCCL")
        for nm in NAMES:
            close(getattr(nccl, nm), getattr(single, nm), rel=1e-11, abs_=1e-13)
        # stand-alone exchange kernel on partial sums that come from elsewhere
        lo, hi = vdist.shard_range(n, rank, world)
        part = torch.empty(vdist.partials_layout(k)["length"], dtype=torch.float64, device=dev)
        ctx.fused_partials(k, n, perm, _cabi.OBJ_GFUNCTION, a, i_begin=lo, i_end=hi, out=part)
        ctx.synchronize()
        alone = vdist.reduce_and_finalize(ctx, k, n, part, _cabi.FLAG_SECOND_ORDER)
        for nm in NAMES:
            assert (getattr(alone, nm) == getattr(multi, nm)).all()                       # same slots, same order, same bits
    # the public API under a process group (Varsens._fused must order the collective and the finalisation)
    k, n = 6, 1 << 16
    v = vb.Varsens(vb.GFunction(A20[:k]), lambda x: x, k, n, verbose=False)
    ref = ctx.run_fused(k, n, vsalt._reference_permutation(n), _cabi.OBJ_GFUNCTION, A20[:k])
    for nm in NAMES:
        close(getattr(v, nm), getattr(ref, nm), rel=1e-11, abs_=1e-13)
    os.environ["VS_NCCL"] = "1"
    v2 = vb.Varsens(vb.GFunction(A20[:k]), lambda x: x, k, n, verbose=False)
    os.environ.pop("VS_NCCL")
    for nm in NAMES:
        close(getattr(v2, nm), getattr(ref, nm), rel=1e-11, abs_=1e-13)
    # two-phase objective (RK4: no fused kernel) through the one-call API: exchange in the stand-alone kernel
    k, n = 4, 4096
    perm = vsalt._reference_permutation(n)
    sc = _cabi.Scale(_cabi.SCALE_LINEAR, numpy.full(k, 0.5), numpy.full(k, 2.0))
    single = ctx.run_fused(k, n, perm, _cabi.OBJ_RK4_CHAIN, [0.02, 50], scale=sc)
    multi = vdist.fused_step(ctx, k, n, perm, _cabi.OBJ_RK4_CHAIN, [0.02, 50], 0, sc, None, _cabi.FLAG_SECOND_ORDER)
    for nm in NAMES:
        close(getattr(multi, nm), getattr(single, nm), rel=1e-10, abs_=1e-12)
    same_on_all_ranks(multi, world)
    # a peer that never shows up: bounded wait, VS_ERR_TIMEOUT, no hang; the exchange is usable again afterwards
    dist.barrier()
    k, n = 6, 70000
    perm = vsalt._reference_permutation(n)
    ex = vdist.peer_exchange(vdist.partials_layout(k)["length"], dev)
    os.environ["VS_P2P_TIMEOUT_MS"] = "300"
    ctx.reload_env()
    lo, hi = vdist.shard_range(n, rank, world)
    if rank == 0:
        try:
            ctx.run_fused_p2p(k, n, perm, _cabi.OBJ_GFUNCTION, A20[:k], ex.world, ex.rank, ex.peer_bufs, ex.peer_flags, ex.next_epoch(), lo, hi)
            raise AssertionError("expected a time-out")
        except _cabi.VarsensError as exc:
            assert exc.status == _cabi.ERR_TIMEOUT, exc
    else:
        ex.next_epoch()                                          # this rank skips the step
    os.environ.pop("VS_P2P_TIMEOUT_MS")
    ctx.reload_env()
    dist.barrier()
    ok = vdist.fused_step(ctx, k, n, perm, _cabi.OBJ_GFUNCTION, A20[:k], 0, _cabi.IDENTITY, None, _cabi.FLAG_SECOND_ORDER)
    ref = ctx.run_fused(k, n, perm, _cabi.OBJ_GFUNCTION, A20[:k])
    for nm in NAMES:
        close(getattr(ok, nm), getattr(ref, nm), rel=1e-11, abs_=1e-13)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("multi-gpu worker ok (world %d)" % world)


if __name__ == "__main__":
    main()
