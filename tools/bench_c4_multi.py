"""C4 (BASELINE.json: sample-export mode, Halton M_a / M_b / N_j matrices, k=50, n=2^22 materialised in HBM) with the BASE ROWS
sharded over the ranks of one node: rank r materialises its contiguous range of base rows of EVERY block of Sample.flat()
(vs_sample_flat_shard) -- 171.13 GB / N per GPU, no collective, each rank generates only its own Halton points.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_c4_multi.py
    python tools/bench_c4_multi.py --split 8      # one GPU doing 1/8 of the rows: the per-rank work of an 8-GPU run

Rank 0 prints one JSON line; time = max over ranks of CUDA-event time per launch (median of 3, after warm-up).  The flat-row
window of the same byte count (the split round 1 had) is timed beside it for comparison."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
import torch.distributed as dist
import varsens_b200 as vb
from varsens_b200 import dist as vdist, saltelli as vsalt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--split", type=int, default=0, help="single process: pretend to be rank 0 of this many")
    ap.add_argument("--k", type=int, default=50)
    ap.add_argument("--logn", type=int, default=22)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = vb.Context.get(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_timing(True)
    k, n = args.k, 1 << args.logn
    parts = args.split if (world == 1 and args.split > 1) else world
    perm = torch.from_numpy(vsalt._reference_permutation(n).astype(numpy.int32)).to(dev)
    lo, hi = vdist.shard_range(n, rank, parts)
    rows = hi - lo
    nbytes = (2 + 2 * k) * rows * k * 8
    out = torch.empty((2 + 2 * k, rows, k), dtype=torch.float64, device=dev)

    def timed(fn):
        fn()
        ts, ks = [], []
        for _ in range(3):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
                torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
            ks.append(ctx.last_kernel_ms())
        t = torch.tensor([float(numpy.median(ts)), float(numpy.median(ks))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    ms, kms = timed(lambda: ctx.sample_flat_shard(k, n, perm, lo, hi, out=out))
    check = float(out[:, :: max(1, rows // 257), :].sum())
    flat = out.view(-1, k)
    w0 = rank * flat.shape[0] if world > 1 else 0
    ms_w, kms_w = timed(lambda: ctx.sample_flat(k, n, perm, row_begin=w0, row_end=w0 + flat.shape[0], out=flat))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    if rank == 0:
        total = nbytes * (world if world > 1 else 1)
        print(json.dumps({"config": "C4 export k=%d n=2^%d, base rows sharded over %d part(s), %d rank(s) running" % (k, args.logn, parts, world),
                          "bytes_per_rank": nbytes, "ms": ms, "kernel_ms": kms, "aggregate_write_gbs": total / (ms * 1e-3) / 1e9,
                          "per_gpu_frac_of_hbm_peak": nbytes / (kms * 1e-3) / 1e9 / hbm, "hbm_peak_gbs": hbm,
                          "flat_row_window_same_bytes": {"ms": ms_w, "kernel_ms": kms_w, "per_gpu_frac_of_hbm_peak": nbytes / (kms_w * 1e-3) / 1e9 / hbm},
                          "checksum": check}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
