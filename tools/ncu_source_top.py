"""Condense the SASS view of an ncu capture (ncu -i x.ncu-rep --page source --csv [| gzip]) into: executed warp
instructions by opcode, stall samples by opcode and reason, the hottest instructions, shared-memory wavefronts by opcode.
    python tools/ncu_source_top.py gpurun_out/r02_ncu_fused_k20_n24_source.csv.gz [top_n] > profiles/r02_fused_sass_profile.txt"""
import csv, gzip, sys, collections, io

path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
allrows = list(csv.reader(io.StringIO(raw)))
# the page holds one section per captured launch: "Kernel Name" row, header row, instruction rows; keep the last launch of
# every distinct kernel (the earlier ones are warm-up launches)
sections, cur = collections.OrderedDict(), None
for r in allrows:
    if r and r[0] == "Kernel Name":
        cur = []
        sections[r[1]] = cur
        cur.append(r)
    elif cur is not None:
        cur.append(r)


def num(r, ix, key):
    try:
        return float(r[ix[key]])
    except (ValueError, KeyError):
        return 0.0


def opcode(src):
    toks = src.split()
    if toks and toks[0].startswith("@"):
        toks = toks[1:]
    return toks[0].rstrip(";") if toks else "?"


def report(rows):
    print("=" * 120)
    print("kernel:", rows[0][1])
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    inst = collections.Counter()
    samp = collections.Counter()
    wave = collections.Counter()
    wave_ideal = collections.Counter()
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    by_reason = collections.Counter()
    for r in body:
        op = opcode(r[ix["Source"]])
        inst[op] += num(r, ix, "Instructions Executed")
        samp[op] += num(r, ix, "# Samples")
        wave[op] += num(r, ix, "L1 Wavefronts Shared")
        wave_ideal[op] += num(r, ix, "L1 Wavefronts Shared Ideal")
        for h in reasons:
            by_reason[h] += num(r, ix, h)
    ti, ts = sum(inst.values()), sum(samp.values())
    print("\nSASS instructions: %d   executed warp instructions: %.4g   stall samples: %d" % (len(body), ti, ts))
    print("\n%-34s %14s %7s %10s %7s" % ("opcode", "warp insts", "%", "samples", "%"))
    for op, c in inst.most_common(topn):
        print("%-34s %14.4g %6.1f%% %10d %6.1f%%" % (op, c, 100 * c / max(ti, 1), samp[op], 100 * samp[op] / max(ts, 1)))
    print("\nstall samples by reason:")
    for h, c in by_reason.most_common(10):
        print("  %-28s %10d %6.1f%%" % (h, c, 100 * c / max(ts, 1)))
    print("\nshared-memory wavefronts by opcode (actual / ideal):")
    for op, c in wave.most_common(8):
        if c:
            print("  %-22s %14.4g / %-14.4g (%.2fx)" % (op, c, wave_ideal[op], c / max(wave_ideal[op], 1)))
    print("\nhottest instructions (stall samples):")
    for r in sorted(body, key=lambda r: -num(r, ix, "# Samples"))[:topn]:
        top = max(reasons, key=lambda h: num(r, ix, h))
        print("  %6d  %-14s %s" % (num(r, ix, "# Samples"), top.replace("stall_", ""), r[ix["Source"]].strip()[:110]))


for rows in sections.values():
    report(rows)
