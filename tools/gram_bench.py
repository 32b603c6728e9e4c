"""Time vs_partials_from_values (estimators on given values) on resident device values.

    python tools/gram_bench.py            # prints one JSON line per configuration

Algorithmic bytes = 8 * (2 + 2k) * l * rows (every value is read once); HBM peak from MEASURED_PEAKS.json.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varsens_b200 import Context  # noqa: E402

CONFIGS = [(20, 1 << 22, 1), (20, 1 << 20, 1), (50, 1 << 20, 1), (6, 1 << 22, 1), (30, 1 << 21, 1), (100, 1 << 18, 1), (6, 1 << 20, 2)]


def main():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6388.0))
    ctx = Context(0)
    ctx.set_timing(True)
    flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
    for k, rows, l in CONFIGS:
        m = (2 + 2 * k) * l
        vals = torch.rand(m * rows, dtype=torch.float64, device="cuda") + 1.0
        out = torch.empty(int(ctx.partials_len(k, l)) if hasattr(ctx, "partials_len") else 4 * l + m * (m + 1) // 2, dtype=torch.float64,
                          device="cuda")
        shift = [1.5] * l
        for mode in ("tensor", "register"):
            if mode == "register":
                os.environ["VS_GRAM_MMA"] = "0"
            else:
                os.environ.pop("VS_GRAM_MMA", None)
            ctx.reload_env()
            best = []
            for it in range(6):
                flush.zero_()
                ctx.partials_from_values(k, l, rows, vals, shift=shift, out=out)
                ctx.synchronize()
                best.append(ctx.last_kernel_ms())
            ms = sorted(best[1:])[len(best[1:]) // 2]
            gb = 8.0 * m * rows / 1e9
            print(json.dumps({"k": k, "rows": rows, "l": l, "kernel": mode, "ms": round(ms, 4), "GBps": round(gb / (ms * 1e-3), 1),
                              "hbm_frac": round(gb / (ms * 1e-3) / hbm, 3), "dmma_tflops": round(2.0 * m * m / 2 * rows / (ms * 1e-3) / 1e12, 2)}))
        os.environ.pop("VS_GRAM_MMA", None)
        del vals


if __name__ == "__main__":
    main()
