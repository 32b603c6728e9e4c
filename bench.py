#!/usr/bin/env python
"""Benchmark of the Saltelli hot path (BASELINE.json metric: model evals/sec + index time).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = one pass of the whole pipeline over the workload BASELINE.json's metric is quoted on:
Sobol g-function, k=20, n=2^24 (704,643,072 evaluations), identity scaling, a=[0,.5,3,9,99,99]+[99]*14,
fused generation + evaluation + reduction, second order included, GENERIC functor path (every point is
evaluated; the separable prefix/suffix shortcut is reported separately under "separable_shortcut").
N>1 shards the n base rows over ranks (total work fixed -> "strong"); ONE kernel launch per rank and step: the 907 partial
sums are exchanged over NVLink peer memory inside the kernel's tail (the NCCL all-reduce arm is timed beside it, "nccl_arm").

`value`   evals/s with the permutation already resident in HBM (device-timed, CUDA events, max over ranks).
`e2e`     same metric through the public C-ABI call with HOST buffers: the permutation is copied from pinned
          host memory while the one launch polls for it, and the kernel writes the indices to mapped host memory,
          inside the timed region, every step.
`roofline` FP64-pipe roofline of the fused kernel: algorithmic flops (6008 per base row at k=20)
          / kernel time (CUDA events around the launch, inside the library) / measured DFMA peak.
`cpu_baseline` the oracle's vectorised numpy pipeline on all host cores, bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K = 20
N_ROWS = 1 << 24
A = [0, 0.5, 3, 9, 99, 99] + [99.0] * 14
METRIC = "model evals/sec (Sobol g-function k=20 n=2^24, fused generation+evaluation+index estimation)"
UNIT = "evals/s"


def evals(n, k=K):
    return 2 * n * (1 + k)


def algorithmic_flops_per_row(k=K):
    # 5k per evaluation x 2(1+k) evaluations + (8k+8) first order + (4k^2+2k) second order
    return 2 * (1 + k) * 5 * k + (8 * k + 8) + (4 * k * k + 2 * k)


# -------------------------------------------------------------------------------------------------
# CPU arm: the oracle's numpy pipeline (BASELINE.md §3.2 "CPU numpy") on all host cores
# -------------------------------------------------------------------------------------------------
_POOL_STATE = {}


def _cpu_chunk(args):
    from oracle import pipeline, objectives
    i0, i1 = args
    st = _POOL_STATE
    S = pipeline.run(K, st["n"], lambda x: x, lambda X: objectives.g_function_rows(X, A), i0=i0, i1=i1,
                     perm=st["perm"], chunk=1 << 14, acc_dtype=numpy.float64, finalize=False)
    return S


def cpu_numpy_evals_per_s(target_seconds=12.0, cores=None):
    """Times the vectorised numpy oracle on a bounded sample of the C3 workload (same n, same permutation, the
    first `rows` base rows).  Returns (evals/s, cores, sample description)."""
    import multiprocessing as mp
    from oracle import pipeline
    cores = cores or os.cpu_count() or 1
    _POOL_STATE["n"] = N_ROWS
    if "perm" not in _POOL_STATE:
        _POOL_STATE["perm"] = pipeline.permutation(N_ROWS)
    ctx = mp.get_context("fork")
    per_task = 1 << 14
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()                                     # calibration: one task per core
        pool.map(_cpu_chunk, [(c * per_task, (c + 1) * per_task) for c in range(cores)])
        cal = time.perf_counter() - t0
        rate = cores * per_task / cal                                # rows/s
        rows = int(min(N_ROWS, max(cores * per_task, rate * target_seconds)))
        rows = (rows // (cores * per_task)) * cores * per_task or cores * per_task
        tasks = [(i, i + per_task) for i in range(0, rows, per_task)]
        t0 = time.perf_counter()
        pool.map(_cpu_chunk, tasks)
        dt = time.perf_counter() - t0
    return evals(rows) / dt, cores, "first %d of %d base rows (%d evals), numpy float64 sums, %d processes" % (
        rows, N_ROWS, evals(rows), cores), dt


def cpu_extras():
    """Two more CPU numbers for context: the C/OpenMP port and the literal per-row Python loop."""
    from oracle import cport, saltelli, objectives
    out = {}
    n_s = 1 << 18
    t0 = time.perf_counter()
    cport.sums(K, N_ROWS, cport.OBJ_GFUNCTION, A, i0=0, i1=n_s)
    out["c_openmp_port_evals_per_s"] = evals(n_s) / (time.perf_counter() - t0)
    out["c_openmp_threads"] = int(cport.lib().orc_num_threads())
    n_l = 128
    t0 = time.perf_counter()
    s = saltelli.Sample(K, n_l, lambda x: x, verbose=False)
    o = saltelli.Objective(K, n_l, s, lambda x: objectives.g_function_row(x, A), verbose=False)
    saltelli.Varsens(o, verbose=False)
    out["literal_python_loop_evals_per_s"] = evals(n_l) / (time.perf_counter() - t0)
    return out


def run_reference(args, rank):
    """--impl reference: the reference cannot run here (Python 2, ghalton absent), so this times the oracle port
    of its pipeline -- vectorised numpy on every host core -- on bounded samples of the same workload."""
    if rank != 0:
        return
    vals, sample = [], ""
    for it in range(args.warmup + args.steps):
        v, cores, sample, dt = cpu_numpy_evals_per_s(target_seconds=8.0)
        if it >= args.warmup:
            vals.append((v, dt))
    value = float(numpy.mean([v for v, _ in vals]))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(numpy.mean([d for _, d in vals])), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3 g-function k=20 n=2^24 (bounded sample per step)", "k": K, "n": N_ROWS},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# -------------------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in out.splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(numpy.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def run_gpu(args, rank, world):
    import torch
    import torch.distributed as dist
    import varsens_b200 as vb
    from varsens_b200 import _cabi, dist as vdist, saltelli as vsalt

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = vb.Context.get(local)
    work_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(work_stream)
    ctx.set_stream(work_stream.cuda_stream)

    n, k = N_ROWS, K
    flags = _cabi.FLAG_SECOND_ORDER
    t0 = time.perf_counter()
    perm_np = vsalt._reference_permutation(n)                     # the reference's own RNG call (saltelli.py:100-101), product code
    perm_draw_s = time.perf_counter() - t0
    perm_host = torch.from_numpy(perm_np.astype(numpy.int32)).pin_memory()
    perm_dev = perm_host.to(dev)
    lo, hi = vdist.shard_range(n, rank, world)
    plen = vdist.partials_layout(k)["length"]
    part = torch.zeros(plen, dtype=torch.float64, device=dev)
    flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    ex = None
    if world > 1 and not vdist.use_nccl():
        ex = vdist.peer_exchange(plen, dev)                             # symmetric memory over NVLink; None -> NCCL
    exchange = "none" if world == 1 else ("one-launch step: peer-memory all-reduce over NVLink inside the fused kernel" if ex is not None
                                          else "NCCL all_reduce between vs_fused_partials and vs_finalize")

    # The step through the public API: a prepared plan (Context.fused_plan: arguments converted once) whose run() is ONE foreign
    # call = one kernel launch per rank.  N=1: vs_run_fused.  N>1: vs_run_fused_p2p (exchange inside the kernel) or, NCCL arm,
    # vs_fused_partials -> all_reduce -> vs_finalize.
    def make_step(perm, use_ex):
        if world == 1:
            return ctx.fused_plan(k, n, perm, _cabi.OBJ_GFUNCTION, A, flags=flags).run
        if use_ex is not None:
            return ctx.fused_plan(k, n, perm, _cabi.OBJ_GFUNCTION, A, flags=flags, i_begin=lo, i_end=hi, exchange=use_ex).run

        def nccl_step():
            ctx.fused_partials(k, n, perm, _cabi.OBJ_GFUNCTION, A, i_begin=lo, i_end=hi, flags=flags, out=part)
            vdist.allreduce_partials(part)                              # same stream as the ctx (set_stream above)
            return ctx.finalize(k, 1, n, part, flags)
        return nccl_step

    step_resident = make_step(perm_dev, ex)
    # public call with HOST buffers: permutation slice H2D from pinned memory, indices to the host, every step
    step_e2e = make_step(perm_host, ex)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, kernel_ms=None, wall_too=False):
        """W untimed warm-up steps, then K timed steps bracketed by a barrier + synchronize on both sides.  Each step is timed on
        the device (CUDA events on the stream the step runs on); the L2 flush between iterations is outside the events.  There
        is no host barrier BETWEEN timed steps: at N>1 the exchange inside the step is the only synchronisation the ranks need
        (a rank that is early waits in its kernel's tail, which the events include).  Returns the max over ranks of the mean."""
        for _ in range(warmup):
            fn()
            flush.zero_()
        barrier()
        tot = 0.0
        res = None
        for _ in range(steps):
            flush.zero_()                                               # L2 flush between timed iterations (stream-ordered, untimed)
            if wall_too:
                torch.cuda.synchronize()                                # the host clock below must not see the flush
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            res = fn()
            e1.record()
            e1.synchronize()
            wall = (time.perf_counter() - t0) * 1e3
            dev_ms = e0.elapsed_time(e1)
            tot += max(dev_ms, wall) if wall_too else dev_ms            # e2e includes the host-side call/return
            if kernel_ms is not None:
                kernel_ms.append(ctx.last_kernel_ms())
        barrier()
        t = torch.tensor([tot / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), res

    peak_tflops = ctx.measure_fp64_peak()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = ctx.launch_count()
    ms, res = timed(step_resident, args.steps, args.warmup)
    launches = (ctx.launch_count() - launches0) // (args.steps + args.warmup) * args.steps
    tail_ns = ctx.last_tail_ns(k).tolist()
    # the fused kernel alone: same steps once more with the library's events around the launch switched on
    kernel_ms = []
    ctx.set_timing(True)
    timed(step_resident, max(3, args.steps // 2), 2, kernel_ms)
    ctx.set_timing(False)
    ms_e2e, res2 = timed(step_e2e, args.steps, max(args.warmup, 3), wall_too=True)
    clocks = sampler.stop() if sampler else None          # sampled over both timed regions

    # the other exchange, for the record (N>1): NCCL all_reduce between the partial-sum launch and the finalize launch
    nccl_ms = None
    if world > 1 and ex is not None:
        nccl_ms, res_nccl = timed(make_step(perm_dev, None), max(3, args.steps // 2), 3)
        nccl_same = bool(numpy.allclose(res_nccl.sens, res.sens, rtol=1e-12, atol=1e-14))

    # index time: from "partial sums of every rank ready" to the indices on the host
    def step_index():
        if world > 1:
            vdist.allreduce_partials(part)
        return ctx.finalize(k, 1, n, part, flags)
    ctx.fused_partials(k, n, perm_dev, _cabi.OBJ_GFUNCTION, A, i_begin=lo, i_end=hi, flags=flags, out=part)
    part_keep = part.clone()
    index_ms, _ = timed(step_index, args.steps, 3)
    part.copy_(part_keep)
    index_values_ms = None
    if world == 1:
        try:
            vals = torch.rand((2 + 2 * k) * n, dtype=torch.float64, device=dev)          # 5.6 GB of stand-in values, flat() order
            index_values_ms, _ = timed(lambda: ctx.indices_from_values(k, 1, n, n, vals, flags=flags), max(3, args.steps // 2), 3)
            del vals
        except Exception as exc:                                                         # extra figure only; never fail the bench on it
            index_values_ms = "failed: %s" % (str(exc).splitlines()[0][:80],)

    # separable shortcut for context; the drop-in Python API, first and cached call
    sep_ms = api_first_ms = api_cached_ms = None
    if world == 1:
        ms_sep, _ = timed(lambda: ctx.run_fused(k, n, perm_dev, _cabi.OBJ_GFUNCTION, A,
                                                flags=flags | _cabi.FLAG_SEPARABLE), max(2, args.steps // 2), 2)
        sep_ms = ms_sep
        vsalt._perm_cache.clear()
        vsalt._perm_dev_cache.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        v = vb.Varsens(vb.GFunction(A), lambda x: x, k, n, verbose=False)
        api_first_ms = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        v = vb.Varsens(vb.GFunction(A), lambda x: x, k, n, verbose=False)
        api_cached_ms = (time.perf_counter() - t0) * 1e3
        api_same = bool((v.sens == res.sens).all())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    kms = float(numpy.mean(kernel_ms))
    rows_rank0 = hi - lo
    flops = algorithmic_flops_per_row(k) * rows_rank0
    achieved = flops / (kms * 1e-3) / 1e12
    executed = (algorithmic_flops_per_row(k) - (2 + 2 * k) * (k - 1)) * rows_rank0 / (kms * 1e-3) / 1e12
    # DRAM traffic of the fused kernel: from the committed ncu --set full capture of this very variant (tools/ncu_summary.py
    # writes the json next to the summary); scaled by rows.  null when no capture of this build is committed.
    traffic, traffic_src = None, None
    tj = os.path.join(ROOT, "profiles", "r02_fused_k20_traffic.json")
    if os.path.exists(tj):
        with open(tj) as fh:
            t = json.load(fh)
        traffic = float(t["dram_bytes"]) * rows_rank0 / float(t["rows"])
        traffic_src = "profiles/%s (ncu --set full, %s, n=%d)" % (t["file"], t["kernel"], t["rows"])
    # the library's default at k = 20: three E-warps per S-warp for single-GPU steps, two for peer-exchange steps (fused_impl.cuh)
    names = {"6": "3 E-warps per S-warp (16 warps, 128 registers)", "5": "2 E-warps per S-warp (12 warps, 168 registers)"}
    vsel = os.environ.get("VS_FUSED_VARIANT", "")
    variant = names.get(vsel if vsel in names else ("6" if world == 1 else "5"), "variant %s" % vsel)
    line = {
        "metric": METRIC, "value": evals(n) / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3 Sobol g-function k=20 n=2^24 identity scaling, fused generic functor, second order on",
                   "k": k, "n": n, "evals_per_step": evals(n),
                   "parallelism": "rows sharded over %d rank(s); exchange of %d doubles: %s" % (world, plen, exchange),
                   "l2": "192 MiB buffer written between timed iterations (L2 flush)",
                   "step": "ONE kernel launch per rank: generation + evaluation + Gram + fixed-order combine%s + estimators, results written to mapped host memory" % (" + peer-memory all-reduce" if ex is not None else "") if (world == 1 or ex is not None) else "vs_fused_partials (1 launch) + NCCL all_reduce + finalize kernel"},
        "index_time": {"partials_to_host_indices_ms": index_ms, "what": "%sfinalize kernel writing to mapped host memory, timed alone" % ("NCCL all-reduce of the partial sums + " if world > 1 else ""),
                       "in_kernel_tail_ns": {"combine_pack": tail_ns[0], "peer_stores": tail_ns[1], "wait_and_sum_peers": tail_ns[2], "estimators_store": tail_ns[3],
                                             "first_level_combine": tail_ns[4], "group_row_fence_ticket": tail_ns[5], "second_level_combine": tail_ns[6], "pack": tail_ns[7],
                                             "cta_prologue": tail_ns[8], "cta_main_loop": tail_ns[9], "cta_combine_ticket": tail_ns[10],
                                             "what": "globaltimer stamps inside the CTA that finished last, last resident step, rank 0"},
                       "values_to_host_indices_ms": index_values_ms, "values_what": "vs_indices_from_values on 2n(1+k) = %d resident values (bulk-copy + DMMA Gram kernel)" % evals(n)},
        "e2e": {"value": evals(n) / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(4 * (hi - lo) * world), "d2h_bytes_per_step": int(8 * (2 + 4 * k + 2 * k * k)),
                "how": "host permutation (pinned) copied in slices on a copy stream; the one fused launch polls an arrival counter per slice; indices land in mapped host memory"},
        "gpu_launches": int(launches),
        # "tensor" = compute bound; the pipe is FP64: evaluation on DFMA, Gram on DMMA (mma.sync m8n8k4.f64), which share one datapath
        "roofline": {"bound": "tensor", "pipe": "fp64 (DFMA + DMMA share one 36-37 TFLOP/s datapath on B200)", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
                     "frac_executed": executed / peak_tflops,
                     "frac_note": "frac credits SURVEY.md 8(d)'s 5k flops per evaluation; the kernel hoists the k divisions by (1+a_c) of a point into one multiply, so frac_executed counts (k-1) fewer flops per evaluation",
                     "traffic": traffic, "traffic_source": traffic_src,
                     "kernel": "vs::fused_wsd_kernel<20, GFunctionReg<20>, false, EPS, 1> (%s, paired-layout DMMA Gram, in-kernel tail)" % variant, "kernel_ms": kms,
                     "kernel_ms_note": "CUDA events around the launch inside the library; at N>1 it includes the wait for the slowest peer in the tail",
                     "algorithmic_flops_per_launch": flops,
                     "peak_source": "DFMA-chain microbenchmark run in this process (vs_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 figure"},
        "clocks": clocks,
        "check": {"var_y": float(res.var_y[0]), "E_2": float(res.E_2[0]), "sens0": float(res.sens[0, 0]),
                  "e2e_matches_resident_bitwise": bool((res.sens == res2.sens).all() and (res.sens_2 == res2.sens_2).all()),
                  "e2e_max_abs_diff_sens": float(numpy.max(numpy.abs(res.sens - res2.sens)))},
        "host_permutation": {"draw_s": perm_draw_s, "what": "numpy.random.seed(1); shuffle of n indices (saltelli.py:100-101), once, outside the timed regions"},
    }
    if nccl_ms is not None:
        line["nccl_arm"] = {"ms_per_step": nccl_ms, "value": evals(n) / (nccl_ms * 1e-3), "unit": UNIT, "matches_peer_memory_step": nccl_same,
                            "what": "vs_fused_partials + torch.distributed.all_reduce (NCCL) + vs_finalize, resident permutation"}
    if sep_ms is not None:
        line["separable_shortcut"] = {"value": evals(n) / (sep_ms * 1e-3), "unit": UNIT, "ms_per_step": sep_ms,
                                      "note": "prefix/suffix products (VS_FLAG_SEPARABLE); not used for value/roofline"}
        line["python_api"] = {"api_first_call_ms": api_first_ms, "api_cached_ms": api_cached_ms, "bitwise_equal_to_c_abi_step": api_same,
                              "what": "Varsens(GFunction(A), lambda x: x, 20, 2**24): the first call draws the seeded permutation on the host and uploads it (64 MB), later calls reuse the device-resident copy"}
    if world == 1 and not args.no_cpu:
        v, cores, sample, _ = cpu_numpy_evals_per_s(target_seconds=12.0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        line["cpu_baseline"].update(cpu_extras())
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # stdout carries exactly one JSON line (NCCL_DEBUG=VERSION prints there otherwise)
    # stdout carries exactly ONE JSON line: whatever a library prints there meanwhile (NCCL's version banner ...) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.impl == "reference":
        return run_reference(args, rank)
    if args.warmup < 3:
        args.warmup = 3
    run_gpu(args, rank, world)


if __name__ == "__main__":
    main()
