"""GPU parity tests: the CUDA path (through the C ABI / the host mirror) against the CPU oracle and the
golden fixtures.  Bit-exact for generated points and assembled sample matrices (identity / linear
scaling); <= 2 ulp for power scaling (libm pow is platform-defined); indices within
1e-10 relative or 1e-12 absolute (BASELINE.json)."""
import math

import numpy
import pytest

from conftest import close
from oracle import cport, exact as oexact, halton as ohalton, objectives as ob, pipeline, saltelli as osalt, scale as oscale, sobol as osobol

pytestmark = pytest.mark.gpu

A6 = [0, 0.5, 3, 9, 99, 99]
A20 = A6 + [99.0] * 14
NAMES = ("E_2", "var_y", "U_j", "U_nj", "sens", "sens_t", "sens_2", "sens_2n")


@pytest.fixture(scope="module")
def vb():
    import varsens_b200
    return varsens_b200


@pytest.fixture(scope="module")
def ctx(vb):
    return vb.Context.get(0)


def perm_of(n):
    return pipeline.permutation(n).astype(numpy.uint32)


def ulp_diff(a, b):
    a, b = numpy.asarray(a, dtype=numpy.float64), numpy.asarray(b, dtype=numpy.float64)
    return numpy.abs(a.view(numpy.int64) - b.view(numpy.int64)).max()


def assert_indices(res, ref, names=NAMES):
    for name in names:
        close(numpy.asarray(getattr(res, name)).reshape(ref[name].shape), ref[name])


# ------------------------------------------------------------------------------------------------
# K1 Halton
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,first,count", [(1, 1, 1), (6, 121, 2048), (20, 401, 5000), (50, 1001, 777),
                                           (20, 33554000, 900), (3, 4294960000, 5000), (97, 1, 64)])
def test_halton_bit_exact(ctx, k, first, count):
    got = ctx.halton(k, first, count)
    assert (got == cport.halton(k, first, count)).all()


def test_halton_golden(ctx, golden):
    assert (ctx.halton(20, 401, 64) == golden["halton_k20_first401"]).all()
    assert (ctx.halton(50, 1001, 16) == golden["halton_k50_first1001"]).all()
    assert (ctx.halton(20, 33554000, 16) == golden["halton_k20_first33554000"]).all()
    assert ctx.halton(5, 1, 1)[0].tolist() == [1 / 2, 1 / 3, 1 / 5, 1 / 7, 1 / 11]


def test_halton_scaled_and_edges(ctx, vb):
    from varsens_b200 import _cabi
    lb, ub = numpy.linspace(-5, 3, 7), numpy.linspace(4, 90, 7)
    got = ctx.halton(7, 141, 1000, _cabi.Scale(_cabi.SCALE_LINEAR, lb, ub))
    assert (got == oscale.linear(ohalton.halton_points(7, 141, 1000), lb, ub)).all()
    lo, up = numpy.linspace(0.01, 3, 7), numpy.linspace(4, 9000, 7)
    got = ctx.halton(7, 141, 1000, _cabi.Scale(_cabi.SCALE_POWER, lo, up))
    assert ulp_diff(got, oscale.power(ohalton.halton_points(7, 141, 1000), lo, up)) <= 2     # tolerance: 2 ulp
    assert ctx.halton(4, 10, 0).shape == (0, 4)                                    # empty
    with pytest.raises(vb.VarsensError):
        ctx.halton(4, 2 ** 32 - 10, 100)                                           # index range
    with pytest.raises(vb.VarsensError):
        ctx.halton(4, 0, 1)                                                        # indices are 1-based


def test_halton_device_output(ctx):
    import torch
    out = torch.empty((3000, 20), dtype=torch.float64, device="cuda:0")
    ctx.halton(20, 401, 3000, out=out)
    ctx.synchronize()
    assert (out.cpu().numpy() == cport.halton(20, 401, 3000)).all()


# ------------------------------------------------------------------------------------------------
# K2 Sobol
# ------------------------------------------------------------------------------------------------
def test_sobol_skip_ahead(ctx, golden):
    V = golden["sobol_joekuo_k8_dirnums"]
    assert (ctx.sobol(8, 4097, 32, V) == golden["sobol_joekuo_k8_from4097"]).all()
    V50 = osobol.joe_kuo_direction_numbers(50)
    for first, count in ((0, 1000), (4097, 3000), (2 ** 31 - 7, 100), (2 ** 32 - 500, 500)):
        assert (ctx.sobol(50, first, count, V50) == osobol.sobol_points(V50, first, count)).all()
    q = ctx.sobol(50, 4097, 2000, V50, quantize6=True)
    assert (q == osobol.quantize_6sig(osobol.sobol_points(V50, 4097, 2000))).all()   # sobolGen.cpp:59 round trip
    assert q[0, 0] == 0.500366
    q = ctx.sobol(3, 1, 70000, V50[:3], quantize6=True)                               # tiny values, e-notation
    assert (q == osobol.quantize_6sig(osobol.sobol_points(V50[:3], 1, 70000))).all()


def test_sobol_sample_entry(vb, golden):
    """quantlib/sobolGen.cpp -> CSV -> Sample(loadFile=...) replaced by sobol_raw -> Sample(raw=...)."""
    k, n = 8, 16
    raw = vb.sobol_raw(k, n)                                              # 6-digit quantised, first point 4097
    assert raw.shape == (2 * n, k)
    assert (raw == osobol.quantize_6sig(golden["sobol_joekuo_k8_from4097"])).all()
    assert (vb.sobol_raw(k, n, quantize6=False) == golden["sobol_joekuo_k8_from4097"]).all()
    assert (vb.joe_kuo_direction_numbers(8) == golden["sobol_joekuo_k8_dirnums"]).all()
    lb, ub = numpy.linspace(-1, 0, k), numpy.linspace(1, 9, k)
    s = vb.Sample(k, n, lambda x: vb.scale.linear(x, lb, ub), verbose=False, raw=raw)
    o = osalt.Sample(k, n, lambda x: oscale.linear(x, lb, ub), verbose=False, raw=raw.copy())
    assert (s.flat() == o.flat()).all()
    v = vb.Varsens(vb.GFunction(numpy.linspace(0, 9, k)), sample=vb.Sample(k, n, verbose=False, raw=raw), verbose=False)
    ref = cport.run(k, n, cport.OBJ_GFUNCTION, numpy.linspace(0, 9, k), raw=raw)
    close(v.sens, ref["sens"])
    close(v.sens_2, ref["sens_2"])


# ------------------------------------------------------------------------------------------------
# K3 sample assembly / export mode
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,n,discard", [(1, 2, 0), (3, 5, 0), (5, 13, 7), (6, 1024, 0), (20, 333, 11), (50, 100, 0),
                                         (130, 9, 3)])
def test_sample_flat_identity_bit_exact(ctx, k, n, discard):
    got = ctx.sample_flat(k, n, perm_of(n), discard)
    assert (got == cport.sample_flat(k, n, discard)).all()


def test_sample_flat_goldens(ctx, golden, refgold):
    assert (ctx.sample_flat(5, 13, perm_of(13), 7) == golden["flat_k5_n13_identity_discard7"]).all()
    from varsens_b200 import _cabi
    lin = _cabi.Scale(_cabi.SCALE_LINEAR, refgold["lin_lb"], refgold["lin_ub"])
    assert (ctx.sample_flat(5, 13, perm_of(13), 7, lin) == refgold["lin_flat_k5_n13_discard7"]).all()
    f = ctx.sample_flat(6, 1024, perm_of(1024))
    assert (f[:8] == refgold["c1_flat_head"]).all() and (f[-8:] == refgold["c1_flat_tail"]).all()
    assert (f[refgold["c1_flat_rows_probe"]] == refgold["c1_flat_probe"]).all()
    ref = refgold["mag_ref"]
    mag = _cabi.Scale(_cabi.SCALE_POWER, ref / 10.0, ref * 10.0)
    assert ulp_diff(ctx.sample_flat(4, 9, perm_of(9), 0, mag), refgold["mag_flat_k4_n9"]) <= 2   # tolerance: 2 ulp
    raw = refgold["raw_in"]
    lin = _cabi.Scale(_cabi.SCALE_LINEAR, numpy.full(4, -1.0), numpy.full(4, 2.0))
    assert (ctx.sample_flat(4, 10, perm_of(10), 0, lin, raw=raw) == refgold["raw_flat_k4_n10"]).all()


def test_sample_flat_windows_and_device_buffers(ctx):
    import torch
    k, n = 7, 1000
    whole = cport.sample_flat(k, n, 5)
    total = 2 * n * (1 + k)
    perm_dev = torch.from_numpy(perm_of(n).astype(numpy.int32)).cuda()
    for lo, hi in ((0, total), (0, 1), (total - 1, total), (17, 18), (999, 1001), (1990, 2050), (2500, 12345),
                   (3, total - 3), (100, 100)):
        out = torch.full((hi - lo, k), -7.0, dtype=torch.float64, device="cuda:0")
        ctx.sample_flat(k, n, perm_dev, 5, row_begin=lo, row_end=hi, out=out)
        ctx.synchronize()
        assert (out.cpu().numpy() == whole[lo:hi]).all(), (lo, hi)
    with pytest.raises(Exception):
        ctx.sample_flat(k, n, perm_of(n), 5, row_begin=5, row_end=total + 1)


@pytest.mark.parametrize("k,n,discard", [(6, 1024, 0), (50, 300, 3), (20, 333, 11), (2, 70, 0), (64, 40, 0), (5, 200, 1), (7, 64, 0)])
def test_sample_flat_shard_and_bulk_kernel(ctx, k, n, discard, monkeypatch):
    """Export mode: the TMA bulk-store kernel (even k), the scalar-store kernel (odd k, VS_NO_BULK_EXPORT=1) and the base-row
    shard addressing (vs_sample_flat_shard: a rank's share of every block) all reproduce the oracle's Sample.flat() bit for bit."""
    import torch
    from varsens_b200 import _cabi
    p = perm_of(n)
    lb, ub = numpy.linspace(-1.0, 0.5, k), numpy.linspace(1.0, 3.0, k)
    for scale, oscale_ in ((_cabi.IDENTITY, None), (_cabi.Scale(_cabi.SCALE_LINEAR, lb, ub), ("linear", lb, ub))):
        whole = cport.sample_flat(k, n, discard, scale=oscale_)
        total = 2 * n * (1 + k)
        assert (ctx.sample_flat(k, n, p, discard, scale) == whole).all()
        for lo, hi in ((0, 1), (n - 1, n + 2), (3, total - 5), (2 * n + 7, 2 * n + 7 + 3 * n + 1), (total - 1, total)):
            assert (ctx.sample_flat(k, n, p, discard, scale, row_begin=lo, row_end=hi) == whole[lo:hi]).all(), (lo, hi)
        monkeypatch.setenv("VS_NO_BULK_EXPORT", "1")
        ctx.reload_env()
        assert (ctx.sample_flat(k, n, p, discard, scale, row_begin=3, row_end=total - 5) == whole[3:total - 5]).all()
        monkeypatch.delenv("VS_NO_BULK_EXPORT")
        ctx.reload_env()
        blocks = whole.reshape(2 + 2 * k, n, k)
        for lo, hi in ((0, n), (0, 1), (n // 3, min(n, n // 3 + 33)), (n - 5, n), (7, 7)):
            got = ctx.sample_flat_shard(k, n, p, lo, hi, discard, scale)
            assert got.shape == (2 + 2 * k, hi - lo, k) and (got == blocks[:, lo:hi, :]).all(), (lo, hi)
        # device output, with an offset that is only 8-byte aligned (falls back to the scalar-store kernel)
        buf = torch.empty((2 + 2 * k) * 40 * k + 1, dtype=torch.float64, device="cuda")
        hi = min(n, 40)
        out = buf[1:1 + (2 + 2 * k) * hi * k].view(2 + 2 * k, hi, k)
        ctx.sample_flat_shard(k, n, torch.from_numpy(p.astype(numpy.int32)).cuda(), 0, hi, discard, scale, out=out)
        ctx.synchronize()                                       # device output: the call only enqueues on the ctx stream
        assert (out.cpu().numpy() == blocks[:, :hi, :]).all()


@pytest.mark.parametrize("env", [{}, {"VS_EXPORT_COPIES": "1"}, {"VS_EXPORT_COPIES": "2"}, {"VS_EXPORT_SLOW_GEN": "1"},
                                 {"VS_HALTON_MODE": "1"}])
def test_sample_flat_bulk_kernel_forms_and_guards(ctx, env, monkeypatch):
    """Every form of the bulk export kernel -- one or two store warps (tile copies), the multiply-only digit loop with
    computed terms and the generic loop, another term-table mode -- is bit-identical to the oracle for windows of every
    shape (inside one block, across the M_1/M_2/N_j/N_nj boundaries, longer than k blocks, ragged last tile), and writes
    nothing outside the window (NaN guard regions around the device output: compute-sanitizer is closed on this pool)."""
    import torch
    from varsens_b200 import _cabi
    for key, val in env.items():
        monkeypatch.setenv(key, val)
    ctx.reload_env()
    if "VS_HALTON_MODE" in env:
        ctx.set_halton_mode(int(env["VS_HALTON_MODE"]))
    try:
        for k, n, discard in ((50, 333, 3), (12, 1000, 0), (4, 97, 0)):
            mode = {"1": "reciprocal"}.get(env.get("VS_HALTON_MODE"), "divide")
            if mode == "divide":
                whole = cport.sample_flat(k, n, discard)
            else:       # the oracle's restatement of that arithmetic as the unscaled points (Sample(raw=...) semantics)
                whole = cport.sample_flat(k, n, discard, raw=ohalton.halton_points_mode(k, 20 * k + discard + 1, 2 * n, mode))
            total = 2 * n * (1 + k)
            pd = torch.from_numpy(perm_of(n).astype(numpy.int32)).cuda()
            pad = 64
            wins = [(0, total), (5, n - 3), (n - 1, n + 1), (2 * n + 11, 2 * n + 3 * n + 5), ((k + 1) * n + 7, (k + 3) * n + 9),
                    (n + 1, (k + 4) * n - 2), (total - 33, total), (2 * n + (k - 1) * n - 5, 2 * n + k * n + 40)]
            for lo, hi in wins:
                buf = torch.full(((hi - lo) * k + 2 * pad,), float("nan"), dtype=torch.float64, device="cuda")
                ctx.sample_flat(k, n, pd, discard, row_begin=lo, row_end=hi, out=buf[pad:pad + (hi - lo) * k].view(hi - lo, k))
                ctx.synchronize()
                got = buf.cpu().numpy()
                assert (got[pad:-pad].reshape(hi - lo, k) == whole[lo:hi]).all(), (k, n, lo, hi)
                assert numpy.isnan(got[:pad]).all() and numpy.isnan(got[-pad:]).all(), (k, n, lo, hi)
            blocks = whole.reshape(2 + 2 * k, n, k)
            for lo, hi in ((0, n), (n // 2, n // 2 + 45), (n - 1, n)):
                sz = (2 + 2 * k) * (hi - lo) * k
                buf = torch.full((sz + 2 * pad,), float("nan"), dtype=torch.float64, device="cuda")
                ctx.sample_flat_shard(k, n, pd, lo, hi, discard, out=buf[pad:pad + sz].view(2 + 2 * k, hi - lo, k))
                ctx.synchronize()
                got = buf.cpu().numpy()
                assert (got[pad:-pad].reshape(2 + 2 * k, hi - lo, k) == blocks[:, lo:hi, :]).all(), (k, n, lo, hi)
                assert numpy.isnan(got[:pad]).all() and numpy.isnan(got[-pad:]).all()
    finally:
        for key in env:
            monkeypatch.delenv(key, raising=False)
        ctx.reload_env()
        ctx.set_halton_mode(0)


def test_sample_flat_large_property(ctx):
    """Structure at a size the oracle cannot hold (k=50, n=2^16 -> 2.7 GB on device): column-substitution
    identities of saltelli.py:119-123 checked on the device with torch, plus a checksum of checksums."""
    import torch
    k, n = 50, 1 << 16
    out = torch.empty((2 * n * (1 + k), k), dtype=torch.float64, device="cuda:0")
    ctx.sample_flat(k, n, perm_of(n), out=out)
    ctx.synchronize()
    M1, M2 = out[:n], out[n:2 * n]
    NJ = out[2 * n:2 * n + k * n].view(k, n, k)
    NN = out[2 * n + k * n:].view(k, n, k)
    for j in (0, 1, 17, 49):
        expect = M2.clone()
        expect[:, j] = M1[:, j]
        assert torch.equal(NJ[j], expect)
        expect = M1.clone()
        expect[:, j] = M2[:, j]
        assert torch.equal(NN[j], expect)
    # every block sums to (k-1) columns of one matrix + 1 column of the other
    c1, c2 = M1.sum(0), M2.sum(0)
    tot = out.sum(0)
    want = (1 + k) * (c1 + c2)
    assert torch.allclose(tot, want, rtol=1e-12)
    head = cport.sample_flat(k, n, row_begin=n - 3, row_end=n + 5)
    assert (out[n - 3:n + 5].cpu().numpy() == head).all()
    assert float(out.min()) > 0.0 and float(out.max()) < 1.0


def test_c4_full_size_shard_properties(ctx):
    """BASELINE config 4 at its FULL design size (k = 50, n = 2^22, Halton indices up to 2^23 + 1000): one rank's base-row shard
    of an 8-GPU export (21.4 GB on the device) -- the column-substitution identities of saltelli.py:119-123 on every block
    pair checked on the device, a checksum of checksums, bit-equality with a flat-row window of the same rows, and rows
    from both ends of the shard against the oracle."""
    import torch
    k, n = 50, 1 << 22
    p = perm_of(n)
    pd = torch.from_numpy(p.astype(numpy.int32)).cuda()
    i0, i1 = 5 * (n // 8), 6 * (n // 8)
    rows = i1 - i0
    out = torch.empty((2 + 2 * k, rows, k), dtype=torch.float64, device="cuda:0")
    ctx.sample_flat_shard(k, n, pd, i0, i1, out=out)
    ctx.synchronize()
    M1, M2 = out[0], out[1]
    for j in (0, 1, 10, 11, 12, 25, 48, 49):
        nj, nn = out[2 + j], out[2 + k + j]
        mask = torch.ones(k, dtype=torch.bool, device="cuda:0")
        mask[j] = False
        assert torch.equal(nj[:, mask], M2[:, mask]) and torch.equal(nj[:, j], M1[:, j]), j
        assert torch.equal(nn[:, mask], M1[:, mask]) and torch.equal(nn[:, j], M2[:, j]), j
    c1, c2 = M1.sum(0), M2.sum(0)
    tot = out.sum(dim=(0, 1))
    assert torch.allclose(tot, (1 + k) * (c1 + c2), rtol=1e-12)
    assert float(out.min()) > 0.0 and float(out.max()) < 1.0
    # the same rows through the flat-row window addressing (block N_j[7]) -- two store warps instead of one
    t = 2 + 7
    win = torch.empty((rows, k), dtype=torch.float64, device="cuda:0")
    ctx.sample_flat(k, n, pd, row_begin=t * n + i0, row_end=t * n + i1, out=win)
    ctx.synchronize()
    assert torch.equal(win, out[t])
    del win
    # oracle rows at both ends of the shard, in M_1, M_2 and two substituted blocks
    for t in (0, 1, 2 + 13, 2 + k + 49):
        for r0 in (i0, i1 - 9):
            want = cport.sample_flat(k, n, row_begin=t * n + r0, row_end=t * n + r0 + 9, perm=p)
            assert (out[t, r0 - i0:r0 - i0 + 9].cpu().numpy() == want).all(), (t, r0)


# ------------------------------------------------------------------------------------------------
# estimators on given values
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,n,l", [(1, 2, 1), (3, 50, 1), (6, 1024, 1), (6, 256, 2), (20, 4096, 1), (50, 600, 1),
                                   (10, 333, 3), (100, 70, 1), (20, 4098, 1), (23, 66, 1), (24, 1000, 1), (31, 130, 1),
                                   (50, 1026, 1), (7, 333, 1), (100, 130, 1), (150, 64, 1)])
def test_indices_from_values(ctx, k, n, l):
    """Contract tolerance (1e-10 relative or 1e-12 absolute) against the estimators evaluated in extended precision
    (oracle/exact.py).  The data is deliberately ill-conditioned (values 10-13, variance 0.75: U_j - E_2 cancels seven
    digits), where the fp64 numpy restatement itself sits 1e-13 - 3e-13 away from the exact answer; against that
    restatement the distance may therefore reach the sum of both rounding noises (bounded at 2e-12 below)."""
    rng = numpy.random.RandomState(k * 1000 + n + l)
    vals = rng.rand(2 * n * (1 + k), l) * 3.0 + 10.0
    res = ctx.indices_from_values(k, l, n, n, vals)
    ex = oexact.indices(vals, k, n)
    for name in NAMES:
        close(getattr(res, name), ex[name].reshape(getattr(res, name).shape), rel=1e-10, abs_=1e-12)
    o = osalt.Objective(k, n, objective_vals=vals, verbose=False)
    v = osalt.Varsens(o, verbose=False)
    for name in NAMES:
        close(getattr(res, name), numpy.asarray(getattr(v, name)).reshape(getattr(res, name).shape), rel=1e-10, abs_=2e-12)


def test_indices_from_values_reference_goldens(ctx, refgold):
    res = ctx.indices_from_values(6, 1, 1024, 1024, refgold["c1_obj_flat"])
    for name in NAMES:
        close(getattr(res, name), refgold["c1_" + name].reshape(getattr(res, name).shape))
    res = ctx.indices_from_values(6, 2, 256, 256, refgold["two_obj_flat"])
    for name in NAMES:
        close(getattr(res, name), refgold["two_" + name].reshape(getattr(res, name).shape))


@pytest.mark.parametrize("k,rows", [(3, 70), (20, 1030), (28, 514), (50, 258), (100, 66)])
def test_partials_tensor_path_matches_register_path(ctx, k, rows, monkeypatch):
    """l = 1 with an even row count runs the bulk-copy + DMMA kernel; VS_GRAM_MMA=0 forces the register-tile kernel.
    Both must give the same sufficient statistics (different summation order -> tolerance), with and without the
    second-order block, and the tensor path must be bit-reproducible."""
    rng = numpy.random.RandomState(k + rows)
    vals = rng.rand(2 + 2 * k, rows) * 2.0 + 5.0
    for flags in (vb_flags_second(), 0):
        monkeypatch.delenv("VS_GRAM_MMA", raising=False)
        ctx.reload_env()                                                    # switches are read at ctx creation / on request only
        a = ctx.partials_from_values(k, 1, rows, vals, shift=[vals[0, 0]], flags=flags)
        a2 = ctx.partials_from_values(k, 1, rows, vals, shift=[vals[0, 0]], flags=flags)
        assert (a == a2).all()
        monkeypatch.setenv("VS_GRAM_MMA", "0")
        ctx.reload_env()
        b = ctx.partials_from_values(k, 1, rows, vals, shift=[vals[0, 0]], flags=flags)
        monkeypatch.delenv("VS_GRAM_MMA", raising=False)
        ctx.reload_env()
        m = 2 + 2 * k
        # compare what the estimators read: sums + Gram rows 0,1 (first order) or the whole upper triangle
        want = vals @ vals.T
        G = numpy.zeros((m, m))
        G[numpy.triu_indices(m)] = a[4:]
        Gb = numpy.zeros((m, m))
        Gb[numpy.triu_indices(m)] = b[4:]
        top = m if flags else 2
        numpy.testing.assert_allclose(G[:top], numpy.triu(want)[:top], rtol=1e-13)
        numpy.testing.assert_allclose(G[:top], Gb[:top], rtol=1e-13)
        numpy.testing.assert_allclose(a[:4], b[:4], rtol=1e-11, atol=1e-9)
        d = vals[:2] - vals[0, 0]
        numpy.testing.assert_allclose(a[:4], [d[0].sum(), d[1].sum(), (d[0] ** 2).sum(), (d[1] ** 2).sum()], rtol=1e-11, atol=1e-9)


@pytest.mark.parametrize("k,rows,l", [(6, 333, 1), (6, 1, 1), (20, 1031, 1), (50, 259, 1), (3, 70, 2), (6, 256, 3), (6, 257, 3),
                                      (2, 3, 5), (10, 333, 3), (20, 515, 2), (4, 1001, 7), (30, 130, 2), (1, 65, 16),
                                      (2, 37, 9), (1, 7, 1), (12, 4099, 3)])
def test_partials_tensor_path_general_form(ctx, k, rows, l, monkeypatch):
    """l > 1 outputs and odd row counts run the general form of the bulk-copy + DMMA kernel (one bulk copy per value block
    and chunk, runs that start 8 bytes off a 16-byte boundary copied from one element earlier); VS_GRAM_GEN=0 forces the
    register-tile kernel.  Same sufficient statistics (coordinate c = t*l + o), with and without the second-order block;
    bit-reproducible."""
    rng = numpy.random.RandomState(100 * k + rows + l)
    nt = 2 + 2 * k
    vals = rng.rand(nt * rows, l) * 2.0 + 5.0                              # Objective.flat(): row t*rows + r, column o
    shift = vals[0].copy()
    V = vals.reshape(nt, rows, l).transpose(0, 2, 1).reshape(nt * l, rows)   # coordinate-major
    want = V @ V.T
    m = nt * l
    d = V[:2 * l] - numpy.concatenate([shift, shift])[:, None]
    for flags in (vb_flags_second(), 0):
        monkeypatch.delenv("VS_GRAM_GEN", raising=False)
        ctx.reload_env()
        a = ctx.partials_from_values(k, l, rows, vals, shift=shift, flags=flags)
        a2 = ctx.partials_from_values(k, l, rows, vals, shift=shift, flags=flags)
        assert (a == a2).all()
        monkeypatch.setenv("VS_GRAM_DEBUG", "1")                            # elimination switch of the tensor-path kernel only:
        ctx.reload_env()                                                    # no Gram update -> proves that kernel ran (no fallback)
        z = ctx.partials_from_values(k, l, rows, vals, shift=shift, flags=flags)
        monkeypatch.delenv("VS_GRAM_DEBUG", raising=False)
        assert (z[4 * l:] == 0.0).all()
        monkeypatch.setenv("VS_GRAM_GEN", "0")
        ctx.reload_env()
        b = ctx.partials_from_values(k, l, rows, vals, shift=shift, flags=flags)
        monkeypatch.delenv("VS_GRAM_GEN", raising=False)
        ctx.reload_env()
        G = numpy.zeros((m, m))
        G[numpy.triu_indices(m)] = a[4 * l:]
        Gb = numpy.zeros((m, m))
        Gb[numpy.triu_indices(m)] = b[4 * l:]
        top = m if flags else 2 * l
        numpy.testing.assert_allclose(G[:top], numpy.triu(want)[:top], rtol=1e-13)
        numpy.testing.assert_allclose(G[:top], Gb[:top], rtol=1e-13)
        sums = numpy.concatenate([d.sum(axis=1), (d ** 2).sum(axis=1)])
        numpy.testing.assert_allclose(a[:4 * l], sums, rtol=1e-11, atol=1e-9)
        numpy.testing.assert_allclose(a[:4 * l], b[:4 * l], rtol=1e-11, atol=1e-9)


def test_partials_general_form_device_buffer_with_guards(ctx):
    """The general form reads one element before / after a misaligned run: with the values in the MIDDLE of a device
    allocation whose neighbours are NaN the result must not change (and nothing may be written around the outputs)."""
    import torch
    k, rows, l = 5, 77, 3
    nt = 2 + 2 * k
    rng = numpy.random.RandomState(9)
    vals = rng.rand(nt * rows, l) + 1.0
    host = ctx.partials_from_values(k, l, rows, vals, shift=vals[0])
    pad = 64
    buf = torch.full((vals.size + 2 * pad,), float("nan"), dtype=torch.float64, device="cuda")
    buf[pad:pad + vals.size] = torch.from_numpy(vals.ravel()).cuda()
    out = torch.full((host.size + 2 * pad,), float("nan"), dtype=torch.float64, device="cuda")
    ctx.partials_from_values(k, l, rows, buf[pad:pad + vals.size], shift=vals[0], out=out[pad:pad + host.size])
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert (got[pad:pad + host.size] == host).all()
    assert numpy.isnan(got[:pad]).all() and numpy.isnan(got[pad + host.size:]).all()


def vb_flags_second():
    from varsens_b200 import _cabi
    return _cabi.FLAG_SECOND_ORDER


def test_partials_shards_sum_to_whole(ctx):
    k, n = 6, 1000
    vals = cport.values(k, n, cport.OBJ_GFUNCTION, A6)                      # (2+2k, n)
    whole = ctx.partials_from_values(k, 1, n, vals, shift=[vals[0, 0]])
    acc = numpy.zeros_like(whole)
    for lo, hi in ((0, 333), (333, 334), (334, 1000)):
        acc += ctx.partials_from_values(k, 1, hi - lo, numpy.ascontiguousarray(vals[:, lo:hi]), shift=[vals[0, 0]])
    numpy.testing.assert_allclose(acc, whole, rtol=1e-13, atol=1e-12)
    res = ctx.finalize(k, 1, n, acc)
    assert_indices(res, cport.run(k, n, cport.OBJ_GFUNCTION, A6))


# ------------------------------------------------------------------------------------------------
# fused pipeline
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,n", [(2, 2), (2, 33), (3, 1000), (4, 64), (5, 31), (6, 1024), (8, 500), (10, 999), (12, 256),
                                 (16, 77), (20, 4096), (7, 300), (9, 130), (11, 257), (13, 64), (14, 100), (15, 333),
                                 (17, 96), (18, 40), (19, 1000), (24, 200), (1, 50)])
def test_fused_gfunction_matches_oracle(ctx, k, n):
    a = (A20 + [1.0, 2.0, 5.0, 99.0])[:k]
    res = ctx.run_fused(k, n, perm_of(n), cport.OBJ_GFUNCTION, a)          # k=1, 24: two-phase path inside the library
    assert_indices(res, cport.run(k, n, cport.OBJ_GFUNCTION, a))


def test_fused_c1_reference_golden(ctx, refgold):
    res = ctx.run_fused(6, 1024, perm_of(1024), cport.OBJ_GFUNCTION, A6)
    for name in NAMES:
        close(getattr(res, name), refgold["c1_" + name].reshape(getattr(res, name).shape))


def test_fused_variants_agree(ctx):
    from varsens_b200 import _cabi
    k, n = 20, 5000
    p = perm_of(n)
    ref = cport.run(k, n, cport.OBJ_GFUNCTION, A20)
    full = ctx.run_fused(k, n, p, cport.OBJ_GFUNCTION, A20)
    sep = ctx.run_fused(k, n, p, cport.OBJ_GFUNCTION, A20, flags=_cabi.FLAG_SECOND_ORDER | _cabi.FLAG_SEPARABLE)
    first = ctx.run_fused(k, n, p, cport.OBJ_GFUNCTION, A20, flags=0)
    assert_indices(full, ref)
    assert_indices(sep, ref)
    assert_indices(first, ref, names=("E_2", "var_y", "U_j", "U_nj", "sens", "sens_t"))
    assert first.sens_2 is None
    again = ctx.run_fused(k, n, p, cport.OBJ_GFUNCTION, A20)
    for name in NAMES:                                                      # fixed-order reductions: bit-reproducible
        assert (getattr(again, name) == getattr(full, name)).all()


def test_fused_scaled_discard_raw(ctx):
    from varsens_b200 import _cabi
    k, n = 6, 700
    lb, ub = numpy.linspace(0.1, 0.3, k), numpy.linspace(0.6, 0.95, k)
    res = ctx.run_fused(k, n, perm_of(n), cport.OBJ_GFUNCTION, A6, discard=9, scale=_cabi.Scale(_cabi.SCALE_LINEAR, lb, ub))
    assert_indices(res, cport.run(k, n, cport.OBJ_GFUNCTION, A6, discard=9, scale=("linear", lb, ub)))
    lo, up = numpy.linspace(0.05, 0.2, k), numpy.linspace(0.7, 1.0, k)
    res = ctx.run_fused(k, n, perm_of(n), cport.OBJ_GFUNCTION, A6, scale=_cabi.Scale(_cabi.SCALE_POWER, lo, up))
    assert_indices(res, cport.run(k, n, cport.OBJ_GFUNCTION, A6, scale=("power", lo, up)))
    raw = numpy.random.RandomState(3).rand(2 * n, k)
    res = ctx.run_fused(k, n, perm_of(n), cport.OBJ_GFUNCTION, A6, raw=raw)
    assert_indices(res, cport.run(k, n, cport.OBJ_GFUNCTION, A6, raw=raw))


def test_fused_shards_equal_whole(ctx):
    k, n = 20, 3000
    p = perm_of(n)
    whole = ctx.fused_partials(k, n, p, cport.OBJ_GFUNCTION, A20)
    acc = numpy.zeros_like(whole)
    for lo, hi in ((0, 1000), (1000, 1001), (1001, 2990), (2990, 3000), (3000, 3000)):
        acc += ctx.fused_partials(k, n, p, cport.OBJ_GFUNCTION, A20, i_begin=lo, i_end=hi)
    numpy.testing.assert_allclose(acc, whole, rtol=1e-13, atol=1e-12)
    assert_indices(ctx.finalize(k, 1, n, acc), cport.run(k, n, cport.OBJ_GFUNCTION, A20))


def test_fused_host_permutation_is_pipelined_and_agrees(ctx, monkeypatch):
    """vs_run_fused / vs_fused_partials with a HOST permutation of >= 2^19 rows cut the H2D copy into slices whose arrival
    the ONE fused launch polls (abi.cu: fused_step).  The row -> warp assignment does not depend on how the permutation
    arrives, so the result is bit-identical to the all-device call and to the unpipelined host path."""
    import torch
    k, n = 6, 1 << 22
    perm = perm_of(n)
    l0 = ctx.launch_count()
    dev = ctx.run_fused(k, n, torch.from_numpy(perm.astype(numpy.int32)).cuda(), cport.OBJ_GFUNCTION, A6)
    assert ctx.launch_count() - l0 == 1                                     # the whole step is one kernel launch
    l0 = ctx.launch_count()
    host = ctx.run_fused(k, n, perm, cport.OBJ_GFUNCTION, A6)
    assert ctx.launch_count() - l0 == 1
    host2 = ctx.run_fused(k, n, torch.from_numpy(perm.astype(numpy.int32)).pin_memory(), cport.OBJ_GFUNCTION, A6)
    for name in NAMES:
        assert (getattr(host, name) == getattr(dev, name)).all()
        assert (getattr(host, name) == getattr(host2, name)).all()
    # a shard with a host permutation
    lo = n // 4 + 3
    a = ctx.fused_partials(k, n, perm, cport.OBJ_GFUNCTION, A6, i_begin=lo, i_end=n)
    monkeypatch.setenv("VS_NO_PIPELINE", "1")
    ctx.reload_env()
    b = ctx.fused_partials(k, n, perm, cport.OBJ_GFUNCTION, A6, i_begin=lo, i_end=n)
    monkeypatch.delenv("VS_NO_PIPELINE", raising=False)
    ctx.reload_env()
    assert (a == b).all()


def test_fused_ishigami(ctx):
    from varsens_b200 import _cabi
    pi = math.pi
    n = 1 << 14
    sc = _cabi.Scale(_cabi.SCALE_LINEAR, numpy.full(3, -pi), numpy.full(3, pi))
    res = ctx.run_fused(3, n, perm_of(n), cport.OBJ_ISHIGAMI, [7.0, 0.1], scale=sc)
    assert_indices(res, cport.run(3, n, cport.OBJ_ISHIGAMI, [7.0, 0.1], scale=("linear", [-pi] * 3, [pi] * 3)))


def test_eval_values_and_rk4(ctx):
    from varsens_b200 import _cabi
    k, n = 20, 64
    ref = numpy.array([1.0] * 10 + [0.5] * 10)
    sc = _cabi.Scale(_cabi.SCALE_POWER, ref / 10.0, ref * 10.0)
    got = ctx.eval_values(k, n, perm_of(n), cport.OBJ_RK4_CHAIN, [0.01, 300], scale=sc)
    want = cport.values(k, n, cport.OBJ_RK4_CHAIN, [0.01, 300], scale=("power", ref / 10.0, ref * 10.0))
    # rate constants within 2 ulp (pow) -> trajectories agree to rounding level; the RK4 arithmetic itself is pinned
    numpy.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-300)
    res = ctx.run_fused(k, n, perm_of(n), cport.OBJ_RK4_CHAIN, [0.01, 300], scale=sc)
    assert_indices(res, cport.run(k, n, cport.OBJ_RK4_CHAIN, [0.01, 300], scale=("power", ref / 10.0, ref * 10.0)))
    for kk in (2, 6, 20, 22, 40):                                            # templated and run-time link counts
        # linear scaling is bit-exact, so device and oracle integrate identical rate constants: with the frozen RK4
        # arithmetic (device.cuh RK4Chain / oracle.c f_rk4_chain) the trajectories must be BIT-IDENTICAL
        r = numpy.linspace(0.5, 2.0, kk)
        got = ctx.eval_values(kk, 16, perm_of(16), cport.OBJ_RK4_CHAIN, [0.02, 250], scale=_cabi.Scale(_cabi.SCALE_LINEAR, r * 0.5, r * 2))
        want = cport.values(kk, 16, cport.OBJ_RK4_CHAIN, [0.02, 250], scale=("linear", r * 0.5, r * 2))
        assert (got == want).all(), "k=%d: %d of %d trajectories differ" % (kk, int((got != want).sum()), got.size)
    got = ctx.eval_values(6, 100, perm_of(100), cport.OBJ_GFUNCTION, A6, i_begin=10, i_end=77)
    want = cport.values(6, 100, cport.OBJ_GFUNCTION, A6, i0=10, i1=77)
    numpy.testing.assert_allclose(got, want, rtol=1e-13)


@pytest.mark.parametrize("k,n", [(24, 200), (50, 700), (11, 64), (12, 33), (100, 70), (3, 1000)])
def test_eval_values_product_form_kernel(ctx, k, n, monkeypatch):
    """Two-phase path, product-form kernel (prefix * term * suffix per row, computed Halton terms for bases >= 37) against the
    oracle's point-by-point values: same products, different association -> a few ulp; and against the point-by-point kernel."""
    from varsens_b200 import _cabi
    a = ([0, 0.5, 3, 9, 99, 99] * 20)[:k]
    p = perm_of(n)
    lb, ub = numpy.linspace(0.1, 0.3, k), numpy.linspace(0.6, 0.95, k)
    for scale, osc, disc in ((_cabi.IDENTITY, None, 0), (_cabi.Scale(_cabi.SCALE_LINEAR, lb, ub), ("linear", lb, ub), 13)):
        want = cport.values(k, n, cport.OBJ_GFUNCTION, a, discard=disc, scale=osc)
        got = ctx.eval_values(k, n, p, cport.OBJ_GFUNCTION, a, discard=disc, scale=scale)
        numpy.testing.assert_allclose(got, want, rtol=1e-13, atol=0)
        lo, hi = n // 3, n - 1
        part = ctx.eval_values(k, n, p, cport.OBJ_GFUNCTION, a, discard=disc, scale=scale, i_begin=lo, i_end=hi)
        assert (part == got[:, lo:hi]).all()
        monkeypatch.setenv("VS_NO_PF_EVAL", "1")
        ctx.reload_env()
        old = ctx.eval_values(k, n, p, cport.OBJ_GFUNCTION, a, discard=disc, scale=scale)
        monkeypatch.delenv("VS_NO_PF_EVAL")
        ctx.reload_env()
        numpy.testing.assert_allclose(got, old, rtol=1e-13, atol=0)
    raw = numpy.random.RandomState(k).rand(2 * n, k)
    numpy.testing.assert_allclose(ctx.eval_values(k, n, p, cport.OBJ_GFUNCTION, a, raw=raw),
                                  cport.values(k, n, cport.OBJ_GFUNCTION, a, raw=raw), rtol=1e-13, atol=0)
    if k > 20:                                                       # indices through the whole two-phase path
        res = ctx.run_fused(k, n, p, cport.OBJ_GFUNCTION, a)
        assert_indices(res, cport.run(k, n, cport.OBJ_GFUNCTION, a))


def test_c5_rk4_frozen_spec_all_indices(ctx):
    """BASELINE config 5 at its frozen spec (k=20 rate constants, scale.magnitude(ref, orders=1), dt=0.01, 1000 steps,
    objective X_10(T)) at n = 2^14: ALL eight outputs against the long-double C oracle, contract tolerance."""
    from varsens_b200 import _cabi
    k, n = 20, 1 << 14
    ref = numpy.array([1.0] * 10 + [0.5] * 10)
    lo, up = ref / 10.0, ref * 10.0                                          # scale.py:121-122 with orders=1, base=10
    sc = _cabi.Scale(_cabi.SCALE_POWER, lo, up)
    p = perm_of(n)
    res = ctx.run_fused(k, n, p, cport.OBJ_RK4_CHAIN, [0.01, 1000], scale=sc)
    want = cport.run(k, n, cport.OBJ_RK4_CHAIN, [0.01, 1000], scale=("power", lo, up))
    assert_indices(res, want)
    # a window of the 1000-step trajectories themselves (rate constants within 2 ulp -> rounding-level agreement)
    got = ctx.eval_values(k, n, p, cport.OBJ_RK4_CHAIN, [0.01, 1000], scale=sc, i_begin=5000, i_end=5064)
    ref_v = cport.values(k, n, cport.OBJ_RK4_CHAIN, [0.01, 1000], scale=("power", lo, up), i0=5000, i1=5064)
    numpy.testing.assert_allclose(got, ref_v, rtol=1e-12, atol=1e-300)


def test_c5_rk4_full_size_shards_and_oracle_rows(ctx):
    """BASELINE config 5 at its FULL size (n = 2^18: 11,010,048 trajectories of 1000 RK4 steps): three ragged shards add up to
    the single launch within the contract tolerance, and the trajectories of the last 64 base rows of the design agree with
    the oracle to rounding level."""
    from varsens_b200 import _cabi
    k, n = 20, 1 << 18
    ref = numpy.array([1.0] * 10 + [0.5] * 10)
    lo, up = ref / 10.0, ref * 10.0
    sc = _cabi.Scale(_cabi.SCALE_POWER, lo, up)
    p = perm_of(n)
    whole = ctx.run_fused(k, n, p, cport.OBJ_RK4_CHAIN, [0.01, 1000], scale=sc)
    acc = None
    for a, b in ((0, 100001), (100001, 100002), (100002, n)):
        part = ctx.fused_partials(k, n, p, cport.OBJ_RK4_CHAIN, [0.01, 1000], scale=sc, i_begin=a, i_end=b)
        acc = part if acc is None else acc + part
    split = ctx.finalize(k, 1, n, acc)
    for name in NAMES:
        close(getattr(split, name), getattr(whole, name))
    assert float(whole.var_y[0]) > 0.0 and numpy.isfinite(whole.sens_2).all()
    got = ctx.eval_values(k, n, p, cport.OBJ_RK4_CHAIN, [0.01, 1000], scale=sc, i_begin=n - 64, i_end=n)
    want = cport.values(k, n, cport.OBJ_RK4_CHAIN, [0.01, 1000], scale=("power", lo, up), i0=n - 64, i1=n)
    numpy.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-300)


# ------------------------------------------------------------------------------------------------
# host mirror: the reference's API
# ------------------------------------------------------------------------------------------------
def test_api_readme_example(vb, refgold):
    """README.md:28-40 with the reference's own Python objective (one call per row)."""
    def gi_function(xi, ai): return (numpy.abs(4.0 * xi - 2.0) + ai) / (1.0 + ai)
    def g_function(x, a): return numpy.prod([gi_function(xi, a[i]) for i, xi in enumerate(x)])
    def g_scaling(x): return x
    def g_objective(x): return g_function(x, [0, 0.5, 3, 9, 99, 99])
    v = vb.Varsens(g_objective, g_scaling, 6, 1024, verbose=False)
    for name in NAMES:
        got = numpy.asarray(getattr(v, name))
        close(got, refgold["c1_" + name].reshape(got.shape))
    assert v.sens.shape == (6, 1) and v.sens_2.shape == (6, 1, 6, 1) and v.E_2.shape == (1,)
    assert (v.objective.flat() == refgold["c1_obj_flat"]).all()              # same rows -> same Python results
    assert (v.sample.flat()[refgold["c1_flat_rows_probe"]] == refgold["c1_flat_probe"]).all()


def test_api_three_routes_agree(vb, refgold):
    import torch
    s = vb.Sample(6, 1024, lambda x: x, verbose=False)
    v_functor = vb.Varsens(vb.GFunction(A6), sample=s, verbose=False)

    @vb.vectorized
    def g_torch(X):
        a = torch.tensor(A6, dtype=torch.float64, device=X.device)
        return torch.prod((torch.abs(4.0 * X - 2.0) + a) / (1.0 + a), dim=1)

    v_torch = vb.Varsens(g_torch, sample=s, verbose=False)
    o = vb.Objective(6, 1024, objective_vals=refgold["c1_obj_flat"], verbose=False)
    v_vals = vb.Varsens(o, verbose=False)
    for v in (v_functor, v_torch, v_vals):
        for name in NAMES:
            got = numpy.asarray(getattr(v, name))
            close(got, refgold["c1_" + name].reshape(got.shape))
    # lazily materialised attributes have the reference's shapes
    assert s.M_1.shape == (1024, 6) and s.N_j.shape == (6, 1024, 6) and s.N_nj.shape == (6, 1024, 6)
    fo = vb.Objective(6, 1024, s, vb.GFunction(A6), verbose=False)
    assert fo.fM_1.shape == (1024, 1) and fo.fN_j.shape == (6, 1024, 1)
    numpy.testing.assert_allclose(fo.flat(), refgold["c1_obj_flat"], rtol=1e-13)


def test_api_two_outputs_and_scalings(vb, refgold):
    s = vb.Sample(6, 256, lambda x: x, verbose=False)
    v = vb.Varsens(lambda x: [ob.g_function_row(x, A6), ob.g_function_row(x, A6[::-1])], sample=s, verbose=False)
    for name in NAMES:
        got = numpy.asarray(getattr(v, name))
        assert got.shape == refgold["two_" + name].shape
        close(got, refgold["two_" + name])
    lb, ub = refgold["lin_lb"], refgold["lin_ub"]
    s = vb.Sample(5, 13, lambda x: vb.scale.linear(x, lb, ub), 7, False)
    assert (s.flat() == refgold["lin_flat_k5_n13_discard7"]).all()
    s = vb.Sample(3, 11, lambda x: vb.scale.percentage(x, numpy.array([1.0, 10.0, 1000.0]), 33.0), verbose=False)
    assert (s.flat() == refgold["pct_flat_k3_n11"]).all()
    s = vb.Sample(5, 13, lambda x: numpy.sqrt(x) * 2.0 - 0.25, verbose=False)            # untraceable -> host callable
    o = osalt.Sample(5, 13, lambda x: numpy.sqrt(x) * 2.0 - 0.25, verbose=False)
    assert (s.flat() == o.flat()).all()
    raw = refgold["raw_in"].copy()
    s = vb.Sample(4, 10, lambda x: vb.scale.linear(x, -1.0, 2.0), verbose=False, raw=raw)
    assert (s.flat() == refgold["raw_flat_k4_n10"]).all() and (raw == refgold["raw_in"]).all()
    assert (s.generate_N_j(s.M_1, s.M_2) == s.N_j).all() and (s.generate_N_j(s.M_2, s.M_1) == s.N_nj).all()


def test_api_reference_unit_tests(vb):
    """varsens/tests/test_sample.py and test_objective.py restated against the GPU-backed classes."""
    x = vb.Sample(11, 13, lambda x: x, 0, False)
    assert x.k == 11 and x.n == 13 and x.M_1.shape == (13, 11) and x.M_2.shape == (13, 11)
    assert x.N_j.shape == (11, 13, 11) and x.N_nj.shape == (11, 13, 11)
    x = vb.Sample(7, 11, lambda x: x, 0, False)
    for m in (x.M_1, x.M_2, x.N_j, x.N_nj):
        assert (m >= 0).all() and (m <= 1).all()
    x = vb.Sample(3, 5, lambda x: x, 0, False)
    for i in range(3):
        assert (x.M_1[:, i] == x.N_j[i][:, i]).all() and (x.M_2[:, i] == x.N_nj[i][:, i]).all()
        for j in range(3):
            if j != i:
                assert (x.M_1[:, j] == x.N_nj[i][:, j]).all() and (x.M_2[:, j] == x.N_j[i][:, j]).all()
    x = vb.Sample(17, 1024, lambda x: x, 0, False)
    assert abs(numpy.sum(x.M_1) / x.n / x.k - 0.5) < 5e-3 and abs(numpy.sum(x.N_nj) / x.n / x.k / x.k - 0.5) < 5e-3
    s = vb.Sample(5, 13, lambda x: x, 0, False)
    assert s.flat().shape == (13 * 12, 5)
    assert numpy.sum(s.M_1[0]) == numpy.sum(s.flat()[0]) and numpy.sum(s.N_nj[4][-1]) == numpy.sum(s.flat()[-1])
    s = vb.Sample(8, 23, lambda x: x, 0, False)
    o = vb.Objective(8, 23, s, lambda x: 1.0 - x, verbose=False)
    assert o.fM_1.shape == (23, 8) and o.fN_j.shape == (8, 23, 8)
    assert abs(numpy.sum(1.0 - o.fM_1 - s.M_1)) < 1e-12 and abs(numpy.sum(1.0 - o.fN_nj - s.N_nj)) < 1e-9
    with pytest.raises(ValueError):
        vb.Varsens(lambda x: 0.0)
    with pytest.raises(Exception, match="requires that a 'scaling'"):
        vb.Sample(3, 5)
    with pytest.raises(Exception, match="Raw sample dimensions"):
        vb.Sample(3, 5, raw=numpy.zeros((9, 3)))


def test_api_export_load_round_trip(vb, tmp_path, refgold, capsys):
    """varsens/tests/test_import_export.py:69-96: export batches -> evaluate -> load == direct run; plus NaN trimming."""
    k, n = 6, 1024
    s = vb.Sample(k, n, lambda x: x, verbose=False)
    v = vb.Varsens(lambda x: ob.g_function_row(x, A6), sample=s, verbose=False)
    s.export(str(tmp_path), "batch", ".csv", 2000)
    nfiles = int(math.ceil(2 * n * (1 + k) / 2000.0))
    for i in range(nfiles):
        rows = numpy.loadtxt(str(tmp_path / ("batch_%d.csv" % (i + 1))))
        numpy.savetxt(str(tmp_path / ("obj_%d.csv" % (i + 1))), ob.g_function_rows(rows, A6))
    o = vb.Objective(k, n, verbose=False, indir=str(tmp_path), prefix="obj", postfix=".csv", nFiles=nfiles)
    v2 = vb.Varsens(o, sample=s, verbose=False)
    for name in NAMES:
        numpy.testing.assert_allclose(getattr(v2, name), getattr(v, name), rtol=0, atol=5e-8)    # 7 places, as the reference test
    s2 = vb.Sample(k, n, verbose=False, indir=str(tmp_path), prefix="batch", postfix=".csv", nFiles=nfiles)
    assert (s2.M_2 == s.M_2).all() and (s2.N_nj == s.N_nj).all()
    # binary batches (.npy): same file naming and row windows, bit-exact both ways
    s.export(str(tmp_path), "bin", ".npy", 3000)
    nbin = int(math.ceil(2 * n * (1 + k) / 3000.0))
    first = numpy.load(str(tmp_path / "bin_1.npy"))
    assert first.shape == (3000, k) and (first == s.flat()[:3000]).all()
    s3 = vb.Sample(k, n, verbose=False, indir=str(tmp_path), prefix="bin", postfix=".npy", nFiles=nbin)
    assert (s3.M_1 == s.M_1).all() and (s3.N_j == s.N_j).all()
    v.objective.export(str(tmp_path), "objbin", ".npy", 5000)
    o4 = vb.Objective(k, n, verbose=False, indir=str(tmp_path), prefix="objbin", postfix=".npy",
                      nFiles=int(math.ceil(2 * n * (1 + k) / 5000.0)))
    assert (o4.flat() == v.objective.flat()).all()
    o3 = vb.Objective(6, 256, objective_vals=refgold["nan_obj_in"], verbose=False)
    assert tuple(o3.fM_1.shape) == tuple(refgold["nan_fM_1_shape"])
    v3 = vb.Varsens(o3, verbose=False)
    for name in ("E_2", "var_y", "sens", "sens_t", "sens_2"):
        got = numpy.asarray(getattr(v3, name))
        close(got, refgold["nan_" + name].reshape(got.shape))
    assert "WARNING: 2 of 3584 objectives were NaN" in capsys.readouterr().out


# ------------------------------------------------------------------------------------------------
# larger sizes: oracle at n = 2^18, analytic truths and invariances at BASELINE's full size
# ------------------------------------------------------------------------------------------------
def test_fused_k20_n2p18_against_c_oracle(ctx):
    n = 1 << 18
    res = ctx.run_fused(20, n, perm_of(n), cport.OBJ_GFUNCTION, A20)
    assert_indices(res, cport.run(20, n, cport.OBJ_GFUNCTION, A20))


def test_fused_full_size_c3(ctx):
    """BASELINE config 3 at FULL size (k=20, n=2^24, 704,643,072 evaluations): all eight outputs against the long-double
    C/OpenMP oracle within the contract tolerance (1e-10 relative or 1e-12 absolute); closed-form truths
    (test_g_function.py:20-50) to 3e-3; the generic and separable kernels agree; a 2-shard split reproduces the single launch;
    the host-permutation (e2e) call is bit-identical to the resident one."""
    from varsens_b200 import _cabi
    import torch
    k, n = 20, 1 << 24
    p = perm_of(n)
    pd = torch.from_numpy(p.astype(numpy.int32)).cuda()
    res = ctx.run_fused(k, n, pd, cport.OBJ_GFUNCTION, A20)
    want = cport.run(k, n, cport.OBJ_GFUNCTION, A20, perm=p)
    assert_indices(res, want)
    var = float(res.var_y[0])
    assert abs(var - ob.g_var(A20)) < 3e-3 and abs(float(res.E_2[0]) - 1.0) < 3e-3
    truth = ob.g_truth(A20)
    for i in range(k):
        assert abs(res.sens[i, 0] * var - truth[i]) < 3e-3
        assert abs(res.sens_t[i, 0] * var - ob.g_truth_t(A20, i)) < 3e-3
    for i, j in ((0, 1), (0, 2), (1, 3), (4, 19)):
        assert abs(res.sens_2[i, 0, j, 0] * var - ob.g_truth_2(A20, i, j)) < 3e-3
        assert abs(res.sens_2n[i, 0, j, 0] * var - ob.g_truth_vnc(A20, [i, j])) < 3e-3
    assert numpy.allclose(res.sens_2[:, 0, :, 0], res.sens_2[:, 0, :, 0].T, rtol=0, atol=1e-15)
    host = ctx.run_fused(k, n, p, cport.OBJ_GFUNCTION, A20)
    for name in NAMES:
        assert (getattr(host, name) == getattr(res, name)).all()
    sep = ctx.run_fused(k, n, pd, cport.OBJ_GFUNCTION, A20, flags=_cabi.FLAG_SECOND_ORDER | _cabi.FLAG_SEPARABLE)
    assert_indices(sep, want)
    acc = ctx.fused_partials(k, n, pd, cport.OBJ_GFUNCTION, A20, i_begin=0, i_end=n // 2 + 7)
    acc = acc + ctx.fused_partials(k, n, pd, cport.OBJ_GFUNCTION, A20, i_begin=n // 2 + 7, i_end=n)
    two = ctx.finalize(k, 1, n, acc)
    assert_indices(two, want)


def test_fused_full_size_c2_ishigami(ctx):
    """BASELINE config 2 at FULL size (Ishigami, k=3, n=2^22, scale.linear(-pi, pi)): all eight outputs against the
    long-double C oracle within the contract tolerance, and the analytic indices to 2e-3."""
    from varsens_b200 import _cabi
    pi = math.pi
    n = 1 << 22
    p = perm_of(n)
    sc = _cabi.Scale(_cabi.SCALE_LINEAR, numpy.full(3, -pi), numpy.full(3, pi))
    res = ctx.run_fused(3, n, p, cport.OBJ_ISHIGAMI, [7.0, 0.1], scale=sc)
    assert_indices(res, cport.run(3, n, cport.OBJ_ISHIGAMI, [7.0, 0.1], scale=("linear", [-pi] * 3, [pi] * 3), perm=p))
    assert abs(float(res.var_y[0]) - 13.8446) < 2e-2
    for got, truth in zip(res.sens[:, 0], (0.3139, 0.4424, 0.0)):
        assert abs(got - truth) < 2e-3
    for got, truth in zip(res.sens_t[:, 0], (0.5576, 0.4424, 0.2437)):
        assert abs(got - truth) < 2e-3
    assert abs(res.sens_2[0, 0, 2, 0] - 0.5576) < 2e-3 and abs(res.sens_2[0, 0, 1, 0] - 0.7563) < 2e-3
