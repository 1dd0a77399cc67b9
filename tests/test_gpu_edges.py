"""GPU: edge cases of the batched C ABI -- empty batches, ragged sizes around CTA / wave
boundaries, scalars 0 / n / > n, infinity operands, every launch shape."""
import numpy as np
import pytest

import bls_oracle as O

pytestmark = pytest.mark.gpu


def ser1(p):
    return p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big")


def ser2(p):
    return b"".join(c.to_bytes(48, "big") for c in (p[0][0], p[0][1], p[1][0], p[1][1]))


def test_empty_batches():
    from bls_b200 import engine
    assert engine.pairing_batch(b"", b"").size == 0
    assert engine.hash_to_g2(b"").size == 0
    assert engine.verify_batch(b"", b"", b"").size == 0
    assert engine.scalar_mul(b"", b"", True).size == 0
    out, ok = engine.decompress(b"", False)
    assert out.size == 0 and ok.size == 0


def test_scalar_edge_values():
    """k = 0, 1, n - 1, n, n + 5, 2^256 - 1 on both groups vs the oracle"""
    from bls_b200 import engine
    ks = [0, 1, O.N - 1, O.N, O.N + 5, (1 << 256) - 1]
    sc = b"".join(k.to_bytes(32, "big") for k in ks)
    for g2, G, ser in ((False, O.G1, ser1), (True, O.G2, ser2)):
        w = 192 if g2 else 96
        out = engine.scalar_mul(ser(G) * len(ks), sc, g2).tobytes()
        for i, k in enumerate(ks):
            want = O.to_aff(O.jac_mul(k % O.N, O.to_jac(G)))
            want_b = bytes(w) if want[2] else ser(want)
            assert out[w * i:w * (i + 1)] == want_b, (g2, k)
    # infinity times anything is infinity
    assert engine.scalar_mul(bytes(96), (5).to_bytes(32, "big"), False).tobytes() == bytes(96)


def test_scalar_mul_small_order_and_off_subgroup_points():
    """the windowed ladder keeps P, 2P, 3P in a table: (0, 2) has order 3 on E(Fq) (3P is infinity,
    2P = -P), so every table entry and the infinity bookkeeping is exercised; off-subgroup points
    (decoded from arbitrary x) must behave like any other point"""
    from bls_b200 import engine
    p3 = (0, 2, False)
    assert O.aff_mul(3, p3)[2] and not O.aff_mul(2, p3)[2]
    ks = list(range(0, 14)) + [O.N - 1, O.N, (1 << 256) - 1, 0x5555555555555555 << 190, 0xaaaaaaaa << 224 | 0xffff]
    sc = b"".join(k.to_bytes(32, "big") for k in ks)
    out = engine.scalar_mul(ser1(p3) * len(ks), sc, False).tobytes()
    for i, k in enumerate(ks):
        want = O.aff_mul(k % 3, p3)
        assert out[96 * i:96 * (i + 1)] == (bytes(96) if want[2] else ser1(want)), k
    # off-subgroup points on both curves
    rng = np.random.default_rng(17)
    for g2 in (False, True):
        w = 192 if g2 else 96
        pts = []
        x = 5
        while len(pts) < 3:
            x += 1
            try:
                cand = (O.g2_y_for_x((x, 1)) if g2 else O.g1_y_for_x(x))
            except ValueError:
                continue
            pts.append(((x, 1), cand[0], False) if g2 else (x, cand[0], False))
        ks = [int.from_bytes(rng.bytes(32), "big") for _ in pts]
        ser = ser2 if g2 else ser1
        out = engine.scalar_mul(b"".join(ser(p) for p in pts), b"".join(k.to_bytes(32, "big") for k in ks), g2).tobytes()
        for i, (p, k) in enumerate(zip(pts, ks)):
            want = O.to_aff(O.jac_mul(k, O.to_jac(p)))
            assert out[w * i:w * (i + 1)] == (bytes(w) if want[2] else ser(want)), (g2, i)


@pytest.mark.parametrize("ctas", [1, 2, 3, 4, 5])
def test_ragged_batches_every_shape(ctas):
    """batch sizes around the CTA (128, 384 for the wide shape 4, 512 for shape 5) and wave boundaries, one known
    pairing repeated"""
    from bls_b200 import _lib, engine
    _lib.init()
    _lib.check(_lib.lib.b200bls_set_ctas_per_sm(ctas))
    try:
        p, q = O.aff_mul(9, O.G1), O.aff_mul(4, O.G2)
        want = O.f12_serialize(O.ate_pairing(p, q))
        wave = _lib.lib.b200bls_sm_count() * 128 * {4: 3, 5: 4}.get(ctas, ctas)
        for n in (1, 127, 129, 383, 385, 511, 513, wave - 1, wave + 1):
            out = engine.pairing_batch(ser1(p) * n, ser2(q) * n).tobytes()
            assert out[:576] == want and out[-576:] == want
            assert out == want * n
    finally:
        _lib.check(_lib.lib.b200bls_set_ctas_per_sm(0))


def test_streams_overlap_and_stay_independent():
    """the same batch enqueued on every library stream gives identical results"""
    from bls_b200 import _lib, engine
    from bls_b200._lib import check, lib
    _lib.init()
    n = 1000
    p, q = O.aff_mul(2, O.G1), O.aff_mul(3, O.G2)
    dP = engine.DeviceBuffer(96 * n).upload(ser1(p) * n)
    dQ = engine.DeviceBuffer(192 * n).upload(ser2(q) * n)
    outs = [engine.DeviceBuffer(576 * n) for _ in range(lib.b200bls_stream_count())]
    for k, o in enumerate(outs):
        check(lib.b200bls_set_stream(k))
        check(lib.b200bls_pairing_batch_dev(dP.ptr, dQ.ptr, o.ptr, n))
    check(lib.b200bls_set_stream(0))
    check(lib.b200bls_sync())
    want = O.f12_serialize(O.ate_pairing(p, q)) * n
    for o in outs:
        assert o.download().tobytes() == want


def test_every_launch_shape_gives_the_same_bytes():
    """each shape runs differently assembled programs (workspace size, spills, Tensor-Memory use, CTA
    width): hash-to-G2, scalar multiplication, Miller loop + final exponentiation, verification and
    decompression must produce identical bytes under all of them"""
    from bls_b200 import _lib, engine, synth
    _lib.init()
    n = 517
    hs = synth.message_hashes(31, n)
    sc = synth.scalars(32, n)
    g1 = np.frombuffer(ser1(O.G1), dtype=np.uint8)
    results = []
    try:
        for ctas in (1, 2, 3, 4, 5):
            _lib.check(_lib.lib.b200bls_set_ctas_per_sm(ctas))
            H = engine.hash_to_g2(hs)
            sig = engine.scalar_mul(H, sc, True)
            pk = engine.scalar_mul(np.tile(g1, n), sc, False)
            sig_bad = sig.copy().reshape(n, 192)
            sig_bad[5], sig_bad[100] = sig_bad[6].copy(), sig_bad[101].copy()
            ok = engine.verify_batch(pk, hs, sig_bad.reshape(-1))
            e = engine.pairing_batch(pk[:96 * 40], H[:192 * 40])
            fe = engine.final_exp_batch(engine.miller_loop_batch(pk[:96 * 40], H[:192 * 40]))
            comp = engine.compress(sig, True)
            back, dok = engine.decompress(comp, True)
            results.append((H.tobytes(), sig.tobytes(), pk.tobytes(), ok.tobytes(), e.tobytes(), fe.tobytes(),
                            comp.tobytes(), back.tobytes(), dok.tobytes()))
    finally:
        _lib.check(_lib.lib.b200bls_set_ctas_per_sm(0))
    for r in results[1:]:
        assert r == results[0]
    ok = np.frombuffer(results[0][3], dtype=np.uint8)
    assert ok.sum() == n - 2 and ok[5] == 0 and ok[100] == 0
    assert results[0][4] == results[0][5]                       # pairing == final_exp(miller_loop)
    assert results[0][7] == results[0][1] and all(results[0][8])
