"""Small invocations of every kernel family, sized for compute-sanitizer (10-100x slow-down):
    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python tools/sanitize_target.py [single|p2p]
`single`: smoke(), the k=20 fused kernel (warp-specialised form, mbarrier hand-offs, in-kernel tail), Ishigami fused, export
(bulk-store kernel and windowed kernel), product-form and RK4 evaluation kernels, tensor-path Gram at k=20 / k=50, the
register-tile Gram (l=3), Halton / Sobol generators.  `p2p` (under torchrun, 2 ranks): the one-launch peer-memory step and the
stand-alone exchange kernel.  Prints "sanitize target ok" at the end; results are cross-checked between routes, not against the
oracle (tests/ does that)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy, torch
import varsens_b200 as vb
from varsens_b200 import _cabi, saltelli, dist as vdist

A20 = [0, .5, 3, 9, 99, 99] + [99.0] * 14
mode = sys.argv[1] if len(sys.argv) > 1 else "single"


def single():
    import __graft_entry__ as g
    g.smoke()
    ctx = vb.Context.get(0)
    dev = torch.device("cuda", 0)
    n = 148 * 32 * 3 + 17                                                # a few batches per CTA and a ragged tail
    p = saltelli._reference_permutation(n)
    pd = torch.from_numpy(p.astype(numpy.int32)).to(dev)
    r = ctx.run_fused(20, n, pd, _cabi.OBJ_GFUNCTION, A20)
    h = ctx.run_fused(20, n, p, _cabi.OBJ_GFUNCTION, A20)                # host permutation: sentinel-polled slices
    assert (r.sens == h.sens).all()
    ctx.run_fused(20, n, pd, _cabi.OBJ_GFUNCTION, A20, flags=_cabi.FLAG_SECOND_ORDER | _cabi.FLAG_SEPARABLE)
    ctx.run_fused(12, 5000, saltelli._reference_permutation(5000), _cabi.OBJ_GFUNCTION, A20[:12])
    ctx.run_fused(7, 5000, saltelli._reference_permutation(5000), _cabi.OBJ_GFUNCTION, A20[:7])
    sc = _cabi.Scale(_cabi.SCALE_LINEAR, numpy.full(3, -math.pi), numpy.full(3, math.pi))
    ctx.run_fused(3, 20000, saltelli._reference_permutation(20000), _cabi.OBJ_ISHIGAMI, [7.0, 0.1], scale=sc)
    # export mode: bulk-store kernel on a base-row shard, windowed kernel on a flat-row window
    k, n = 50, 3000
    p = saltelli._reference_permutation(n)
    shard = ctx.sample_flat_shard(k, n, p, 100, 2100)
    win = ctx.sample_flat(k, n, p, row_begin=n + 100, row_end=n + 2100)
    assert (numpy.asarray(shard)[1] == numpy.asarray(win)).all()
    # two-phase path: product-form evaluation, Gram of the values (k=50: super-tile form; k=20: whole triangle per warp)
    a50 = A20 + [99.0] * 30
    v50 = ctx.eval_values(k, n, p, _cabi.OBJ_GFUNCTION, a50)
    ctx.indices_from_values(k, 1, n, n, numpy.ascontiguousarray(v50))
    n = 4098
    p = saltelli._reference_permutation(n)
    v20 = ctx.eval_values(20, n, p, _cabi.OBJ_GFUNCTION, A20)
    ctx.indices_from_values(20, 1, n, n, numpy.ascontiguousarray(v20))
    ref = numpy.array([1.0] * 10 + [0.5] * 10)
    scp = _cabi.Scale(_cabi.SCALE_POWER, ref / 10.0, ref * 10.0)
    ctx.eval_values(20, 256, saltelli._reference_permutation(256), _cabi.OBJ_RK4_CHAIN, [0.01, 50], scale=scp)
    rng = numpy.random.RandomState(3)
    ctx.indices_from_values(6, 3, 333, 333, rng.rand(2 * 333 * 7, 3) + 1.0)       # register-tile Gram (l > 1)
    ctx.indices_from_values(6, 1, 333, 333, rng.rand(2 * 333 * 7, 1) + 1.0)       # odd row count
    ctx.halton(20, 401, 5000)
    s = vb.Sample(6, 512, lambda x: x, verbose=False)
    s.flat()
    print("sanitize target ok (single), launches=%d" % ctx.launch_count())


def p2p():
    import torch.distributed as tdist
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    tdist.init_process_group("nccl", device_id=dev)
    os.environ["VS_P2P_STRICT"] = "1"
    ctx = vb.Context.get(local)
    n = 2 * 148 * 32 * 2 + 9
    p = saltelli._reference_permutation(n)
    for _ in range(3):
        r = vdist.fused_step(ctx, 20, n, p, _cabi.OBJ_GFUNCTION, A20, 0, _cabi.IDENTITY, None, _cabi.FLAG_SECOND_ORDER)
    lo, hi = vdist.shard_range(n, rank, 2)
    part = ctx.fused_partials(20, n, p, _cabi.OBJ_GFUNCTION, A20, i_begin=lo, i_end=hi, out=torch.empty(907, dtype=torch.float64, device=dev))
    q = vdist.reduce_and_finalize(ctx, 20, n, part, _cabi.FLAG_SECOND_ORDER)
    d = numpy.abs(q.sens - r.sens).max()
    assert d < 1e-12, d
    torch.cuda.synchronize()
    tdist.barrier()
    print("sanitize target ok (p2p) rank %d" % rank)
    tdist.destroy_process_group()


if mode == "p2p":
    p2p()
else:
    single()
