"""Brute-force check of the multiply-only digit step used by the fused kernels (csrc/fused_impl.cuh: digit_step<B,true>):
   c = ceil(2^32 / b);  w = m * c (64-bit);  q = w >> 32;  8*(m mod b) = ((w & 0xffffffff) * 8b) >> 32
is exact for every m with m * 8b < 2^32.  (Proof: c = (2^32+e)/b, 0<=e<b; m = qb+r => w = q 2^32 + (r 2^32 + m e)/b, the low
word is (r 2^32 + m e)/b, times 8b over 2^32 is 8r + 8 m e / 2^32, and 8 m e < 8 m b < 2^32.)"""
import numpy as np

def primes(n):
    out, c = [], 2
    while len(out) < n:
        if all(c % p for p in out if p * p <= c):
            out.append(c)
        c += 1
    return out

bad = 0
for b in primes(64)[1:]:
    c = -(-2 ** 32 // b)
    lim = (2 ** 32) // (8 * b)
    m = np.concatenate([np.arange(0, min(lim, 1 << 22), dtype=np.uint64),
                        np.random.RandomState(b).randint(0, lim, size=2_000_000).astype(np.uint64),
                        np.arange(max(0, lim - 100000), lim, dtype=np.uint64)])
    w = m * np.uint64(c)
    q, low = w >> np.uint64(32), w & np.uint64(0xFFFFFFFF)
    d8 = (low * np.uint64(8 * b)) >> np.uint64(32)
    ok = bool((q == m // np.uint64(b)).all() and (d8 == np.uint64(8) * (m % np.uint64(b))).all())
    bad += not ok
    print("base %3d  magic %10d  valid for m < %9d  %s" % (b, c, lim, "ok" if ok else "FAIL"))
raise SystemExit(bad)
