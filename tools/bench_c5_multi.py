"""C5 (BASELINE.json: RK4 mass-action chain, k=20 rate constants, magnitude scaling, n=2^18, dt=0.01, 1000 steps) with
the base rows sharded over the ranks of one node -- the multi-GPU form of tools/bench_configs.py's C5 entry.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_c5_multi.py

Every rank evaluates its slice through the library's two-phase path (values to HBM scratch, tensor-path Gram), the
907-double partial-sum vectors are all-reduced (NCCL) and every rank finalises.  Rank 0 prints one JSON line; time is
the max over ranks of CUDA-event time per step, after warm-up, with a barrier before every step."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
import torch.distributed as dist
import varsens_b200 as vb
from varsens_b200 import _cabi, dist as vdist


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = vb.Context.get(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    k, n = 20, 1 << 18
    ref = numpy.array([1.0] * 10 + [0.5] * 10)
    sc = _cabi.Scale(_cabi.SCALE_POWER, ref / 10.0, ref * 10.0)       # scale.magnitude(points, ref, orders=1.0), scale.py:121-122
    perm = torch.from_numpy(numpy.random.RandomState(1).permutation(n).astype(numpy.int32)).to(dev)
    lo, hi = vdist.shard_range(n, rank, world)
    plen = vdist.partials_layout(k)["length"]
    part = torch.zeros(plen, dtype=torch.float64, device=dev)
    res_dev = torch.empty(_cabi.Result.flat_len(k, 1), dtype=torch.float64, device=dev)
    flags = _cabi.FLAG_SECOND_ORDER

    def step():
        ctx.fused_partials(k, n, perm, _cabi.OBJ_RK4_CHAIN, [0.01, 1000], scale=sc, i_begin=lo, i_end=hi, flags=flags, out=part)
        if world > 1:
            vdist.allreduce_partials(part)
        ctx.finalize_device(k, 1, n, part, res_dev, flags)
        return _cabi.Result.from_flat(k, 1, res_dev.cpu().numpy())

    for _ in range(2):
        res = step()
    tot, steps = 0.0, 5
    for _ in range(steps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = step()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    t = torch.tensor([tot / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        traj = 2 * n * (1 + k)
        print(json.dumps({"config": "C5 RK4 chain k=20 n=2^18 dt=0.01 1000 steps, magnitude scaling", "n_gpus": world, "ms_per_step": ms,
                          "trajectories": traj, "trajectories_per_s": traj / (ms * 1e-3),
                          "tflops_est": traj * 1000 * 336.0 / (ms * 1e-3) / 1e12,
                          "var_y": float(res.var_y[0]), "sens0": float(res.sens[0, 0]), "sens_t0": float(res.sens_t[0, 0])}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
