"""One table row per distinct kernel from the raw-page CSVs the profiling run leaves in gpurun_out/ (tools/run_profiles.sh):
    python tools/ncu_table.py gpurun_out/r02_ncu_*_raw.csv > profiles/r02_ncu_kernels.md
Launches of the same kernel are averaged (the first launch of a kernel in a capture is the warm-up launch unless ncu skipped it)."""
import csv, sys, collections, re

COLS = [("gpu__time_duration.sum", "time"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "fp64 pipe %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts %"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def fmt(name, v):
    if v != v:
        return "-"
    if name.startswith("dram rd") or name.startswith("dram wr"):
        return "%.4g MB" % (v / 1e6)
    if name == "time":
        return "%.4g ms" % v
    if name in ("regs", "grid", "block", "smem conflicts"):
        return "%d" % round(v)
    return "%.1f" % v


print("| capture | kernel | launches | " + " | ".join(n for _, n in COLS) + " | top stalls (per issue) |")
print("|---|---|---|" + "---|" * (len(COLS) + 1))
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("vs::", "")
        agg.setdefault(name, []).append(r)
    cap = re.sub(r".*r02_ncu_|_raw.csv", "", path)
    for name, rs in agg.items():
        cells = []
        for key, short in COLS:
            if key not in ix:
                cells.append("-")
                continue
            vals = []
            for r in rs:
                try:
                    x = float(r[ix[key]].replace(",", "")) * UNIT.get(units[ix[key]], 1.0)
                    if x == x:
                        vals.append(x)
                except ValueError:
                    pass
            cells.append(fmt(short, sum(vals) / len(vals)) if vals else "-")
        st = {}
        for h, i in ix.items():
            m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", h)
            if m and m.group(1) not in ("selected", "not_selected"):
                st[m.group(1)] = sum(float(r[i]) for r in rs) / len(rs)
        top = ", ".join("%s %.2f" % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print("| %s | `%s` | %d | %s | %s |" % (cap, name[:70], len(rs), " | ".join(cells), top))
