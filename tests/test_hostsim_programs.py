"""CPU checks of the DEVICE code paths: the assembled VM programs run on the host build of the
interpreter (tests/hostsim: csrc/fp.cuh + vm_exec.cuh with every PTX instruction emulated, the
kernel's cell / cold / Tensor-Memory-slot layouts) and must reproduce the golden vectors.  This
is what lets formulas, the allocator and the limb arithmetic be validated without a GPU."""
import ctypes
import os

import numpy as np
import pytest

import bls_oracle as O
import hostsim
from conftest import load_golden

from bls_b200.programs import curve, fieldops, hashg2, pairing, registry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pq(cases):
    P = b"".join(bytes.fromhex(c["p"]["x"]) + bytes.fromhex(c["p"]["y"]) for c in cases)
    Q = b"".join(bytes.fromhex(c["q"]["x"]) + bytes.fromhex(c["q"]["y"]) for c in cases)
    return np.frombuffer(P, dtype=np.uint8).copy(), np.frombuffer(Q, dtype=np.uint8).copy()


@pytest.mark.parametrize("ctas", sorted(registry.SHAPES))
def test_pairing_program_every_launch_shape(ctas):
    n_slots, n_tmem = registry.SHAPES[ctas]
    g = load_golden("pairing_kat.json")
    cases = g["pairs"][:5] + g["degenerate"]
    asm = pairing.build_pairing().assemble(n_slots, n_cold=4096, n_tmem=n_tmem)
    P, Q = _pq(cases)
    out = np.zeros(576 * len(cases), dtype=np.uint8)
    hostsim.run(asm, {0: P, 1: Q, 2: out}, {0: 96, 1: 192, 2: 576}, len(cases), n_blocks=2, nt=3)
    raw = out.tobytes()
    for i, c in enumerate(cases):
        assert raw[576 * i:576 * (i + 1)].hex() == c["out"], (ctas, i)


def test_sha_stage_matches_hashlib():
    lib = hostsim.lib()
    hs = b"".join(O.hash256(bytes([i])) for i in range(5))
    out = np.zeros(256 * 5, dtype=np.uint8)
    lib.hs_sha_stage(hs, out.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(5))
    raw = out.tobytes()
    for i in range(5):
        h = hs[32 * i:32 * (i + 1)]
        want = b"".join(O.hash512(h + b"G2_" + j + b"_c" + k) for j in (b"0", b"1") for k in (b"0", b"1"))
        assert raw[256 * i:256 * (i + 1)] == want


def test_hash_pks_exponents_device_code():
    """the per-key part of hash_pks (util.py:46-49) as the CUDA kernel computes it, against the
    oracle and the reference's own exponents (agg_kat / sig_kat were produced with them)"""
    lib = hostsim.lib()
    pk_hash = O.hash256(b"some public keys")
    n = 300
    out = np.zeros(32 * n, dtype=np.uint8)
    lib.hs_hash_pks(pk_hash, ctypes.c_uint32(0), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(n))
    raw = out.tobytes()
    for i in range(n):
        want = int.from_bytes(O.hash256(i.to_bytes(4, "big") + pk_hash), "big") % O.N
        assert int.from_bytes(raw[32 * i:32 * (i + 1)], "big") == want
    # a digest >= 2n needs both subtractions, one in [n, 2n) one, one below n none: all three occur
    vals = [int.from_bytes(O.hash256(i.to_bytes(4, "big") + pk_hash), "big") // O.N for i in range(n)]
    assert set(vals) == {0, 1, 2}
    # non-zero first index
    lib.hs_hash_pks(pk_hash, ctypes.c_uint32(0xfffffff0), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(8))
    for i in range(8):
        want = int.from_bytes(O.hash256((0xfffffff0 + i).to_bytes(4, "big") + pk_hash), "big") % O.N
        assert int.from_bytes(out.tobytes()[32 * i:32 * (i + 1)], "big") == want


def test_hash_and_verify_programs():
    g = load_golden("hash_kat.json")
    cases = g["hash_to_g2_prehashed"][:6]
    lib = hostsim.lib()
    hs = b"".join(bytes.fromhex(c["h"]) for c in cases)
    sha = np.zeros(256 * len(cases), dtype=np.uint8)
    lib.hs_sha_stage(hs, sha.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(len(cases)))
    asm = hashg2.build_hash_to_g2().assemble(6, n_cold=4096, n_tmem=5)
    H = np.zeros(192 * len(cases), dtype=np.uint8)
    hostsim.run(asm, {0: sha, 1: H}, {0: 256, 1: 192}, len(cases), n_blocks=1, nt=4)
    raw = H.tobytes()
    for i, c in enumerate(cases):
        assert raw[192 * i:192 * (i + 1)].hex() == c["out"]["x"] + c["out"]["y"], i
    # verification core on the reference's own signature vectors
    sg = load_golden("sig_kat.json")
    tab = sg["verify_table"][:4]

    def ser1(p):
        return p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big")

    def ser2(p):
        return b"".join(c.to_bytes(48, "big") for c in (p[0][0], p[0][1], p[1][0], p[1][1]))
    pk = np.frombuffer(b"".join(ser1(O.g1_deserialize(bytes.fromhex(t["pk"]))) for t in tab), dtype=np.uint8).copy()
    Hm = np.frombuffer(b"".join(ser2(O.hash_to_g2_prehashed(bytes.fromhex(t["h"]))) for t in tab), dtype=np.uint8).copy()
    sig = np.frombuffer(b"".join(ser2(O.g2_deserialize(bytes.fromhex(t["sig"]))) for t in tab), dtype=np.uint8).copy()
    ok = np.full(len(tab), 9, dtype=np.uint8)
    asm = pairing.build_verify_pair().assemble(9, n_cold=4096, n_tmem=10)
    hostsim.run(asm, {0: pk, 1: Hm, 2: sig, 3: ok}, {0: 96, 1: 192, 2: 192, 3: 1}, len(tab), n_blocks=1, nt=4)
    assert [bool(x) for x in ok] == [t["ok"] for t in tab]
    # the fused program (hash-to-G2 feeding the Miller loop as a projective point, no inversion)
    tab = sg["verify_table"][:6]
    pk = np.frombuffer(b"".join(ser1(O.g1_deserialize(bytes.fromhex(t["pk"]))) for t in tab), dtype=np.uint8).copy()
    sig = np.frombuffer(b"".join(ser2(O.g2_deserialize(bytes.fromhex(t["sig"]))) for t in tab), dtype=np.uint8).copy()
    hs = b"".join(bytes.fromhex(t["h"]) for t in tab)
    sha = np.zeros(256 * len(tab), dtype=np.uint8)
    lib.hs_sha_stage(hs, sha.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(len(tab)))
    ok = np.full(len(tab), 9, dtype=np.uint8)
    asm = pairing.build_verify_full().assemble(6, n_cold=4096, n_tmem=7)
    hostsim.run(asm, {0: pk, 1: sha, 2: sig, 3: ok}, {0: 96, 1: 256, 2: 192, 3: 1}, len(tab), n_blocks=2, nt=3)
    assert [bool(x) for x in ok] == [t["ok"] for t in tab]
    assert any(t["ok"] for t in tab) and not all(t["ok"] for t in tab)


@pytest.mark.parametrize("g2", [False, True])
def test_sum_programs_with_skipped_and_executed_regions(g2):
    """two-pass point sums incl. P + P, P + (-P), infinity; SKIPZ regions skipped and not"""
    G = O.G2 if g2 else O.G1
    w = 192 if g2 else 96

    def ser(p):
        if g2:
            return b"".join(c.to_bytes(48, "big") for c in (p[0][0], p[0][1], p[1][0], p[1][1]))
        return p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big")
    pts = [O.aff_mul(k, G) for k in (3, 5, 7, 11, 13)]
    plist = pts + [pts[1], pts[1], O.aff_neg(pts[2])]
    data = b"".join(ser(p) for p in plist) + bytes(w)
    n = len(data) // w
    want = ser((O.g2_sum if g2 else O.g1_sum)(plist))
    for honor, (ns, ntm) in ((True, (18, 0)), (False, (18, 0)), (True, (9, 10))):
        # (9, 10): cross-thread reads with part of the workspace in (thread-private) tensor memory
        a1 = curve.build_sum_pass1(g2)().assemble(ns, n_tmem=ntm)
        a2 = curve.build_sum_pass2(g2)().assemble(ns, n_tmem=ntm)
        nb = 2
        # raw (Montgomery-form, weakly reduced) intermediates are equal mod q, not byte for byte, between the two
        # executors: each one is run end to end
        for paired in (False, True):
            raw = np.zeros((3 if g2 else 2) * 6 * nb * 16, dtype=np.uint8)
            hostsim.run(a1, {0: np.frombuffer(data, dtype=np.uint8).copy(), 1: raw}, {0: w, 1: nb}, n,
                        n_blocks=nb, nt=128, honor_skips=honor, paired=paired)
            out = np.zeros(w, dtype=np.uint8)
            hostsim.run(a2, {0: raw, 1: out}, {0: nb, 1: w}, nb, n_blocks=1, nt=128, honor_skips=honor, paired=paired)
            assert out.tobytes() == want, paired


@pytest.mark.parametrize("g2", [False, True])
def test_one_launch_small_sum_program(g2):
    """g?_sums: fold, CTA tree and to_affine in one program on one block -- fewer points than threads, more points
    than threads (several item blocks), P + P, P + (-P), infinity"""
    G = O.G2 if g2 else O.G1
    w = 192 if g2 else 96

    def ser(p):
        if p[2]:
            return bytes(w)
        if g2:
            return b"".join(c.to_bytes(48, "big") for c in (p[0][0], p[0][1], p[1][0], p[1][1]))
        return p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big")
    pts = [O.aff_mul(k, G) for k in (3, 5, 7, 11, 13)]
    asm = curve.build_sum_small(g2)().assemble(9, n_tmem=10)
    for plist in (pts[:1], pts[:2] + [O.aff_neg(pts[1])], pts + [pts[1], pts[1], O.aff_neg(pts[2])] + pts * 30):
        data = b"".join(ser(p) for p in plist) + bytes(w)                 # + the point at infinity
        n = len(data) // w
        want = ser((O.g2_sum if g2 else O.g1_sum)(plist))
        for paired in (False, True):
            out = np.zeros(w, dtype=np.uint8)
            hostsim.run(asm, {0: np.frombuffer(data, dtype=np.uint8).copy(), 1: out}, {0: w, 1: w}, n, n_blocks=1, nt=128,
                        paired=paired)
            assert out.tobytes() == want, (g2, len(plist), paired)


@pytest.mark.parametrize("g2", [False, True])
def test_three_pass_sum_programs(g2):
    """large sums: per-thread fold (g?_sumf, no cross-thread step, workspace partly in tensor memory), Jacobian fold +
    CTA tree (g?_sum1j), one CTA (g?_sum2) -- incl. P + P, P + (-P), infinity, more points than threads (several
    item blocks per thread) and fewer (threads whose partial stays the identity)"""
    G = O.G2 if g2 else O.G1
    w = 192 if g2 else 96

    def ser(p):
        if g2:
            return b"".join(c.to_bytes(48, "big") for c in (p[0][0], p[0][1], p[1][0], p[1][1]))
        return p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big")
    pts = [O.aff_mul(k, G) for k in (3, 5, 7, 11, 13, 17, 19)]
    for plist, nt_a, nb_a in ((pts + [pts[1], pts[1], O.aff_neg(pts[2])] + pts[3:], 3, 2),      # 14 points + infinity on 6 threads
                              (pts[:3], 4, 2)):                                                # 3 points + infinity on 8 threads
        data = b"".join(ser(p) for p in plist) + bytes(w)
        n = len(data) // w
        want = ser((O.g2_sum if g2 else O.g1_sum)(plist))
        af = curve.build_sum_fold(g2)().assemble(6, n_cold=4096, n_tmem=7)
        aj = curve.build_sum_pass1j(g2)().assemble(9, n_tmem=10)
        a2 = curve.build_sum_pass2(g2)().assemble(9, n_tmem=10)
        elems = 3 if g2 else 2
        parts = nt_a * nb_a
        raw_a = np.zeros(elems * 6 * parts * 16, dtype=np.uint8)
        hostsim.run(af, {0: np.frombuffer(data, dtype=np.uint8).copy(), 1: raw_a}, {0: w, 1: parts}, n,
                    n_blocks=nb_a, nt=nt_a, paired=False)
        n_parts = min(n, parts)                     # threads beyond the item count store nothing (sum_dev: n >= parts)
        nb = 2
        raw_b = np.zeros(elems * 6 * nb * 16, dtype=np.uint8)
        hostsim.run(aj, {0: raw_a, 1: raw_b}, {0: parts, 1: nb}, n_parts, n_blocks=nb, nt=128, paired=False)
        out = np.zeros(w, dtype=np.uint8)
        hostsim.run(a2, {0: raw_b, 1: out}, {0: nb, 1: w}, nb, n_blocks=1, nt=128, paired=False)
        assert out.tobytes() == want, (g2, len(plist))


def test_field_programs_on_weakly_reduced_edge_values():
    """operands 0, 1, q-1 and values >= q on input (reduced on load): mul / inv / sub at level 12"""
    import random
    rnd = random.Random(3)
    Q = O.Q
    vals = [(0,) * 12, (1,) + (0,) * 11, (Q - 1,) * 12] + [tuple(rnd.randrange(Q) for _ in range(12)) for _ in range(3)]

    def ser(e):
        return b"".join(int(c).to_bytes(48, "big") for c in e)
    a = np.frombuffer(b"".join(ser(v) for v in vals), dtype=np.uint8).copy()
    b = np.frombuffer(b"".join(ser(v) for v in reversed(vals)), dtype=np.uint8).copy()
    for op, fn in (("mul", O.f12_mul), ("sub", O.f12_sub)):
        asm = fieldops.build_field_op(12, op)().assemble(6, n_cold=4096, n_tmem=5)
        out = np.zeros(576 * len(vals), dtype=np.uint8)
        hostsim.run(asm, {0: a, 1: b, 2: out}, {0: 576, 1: 576, 2: 576}, len(vals), n_blocks=1, nt=2)
        raw = out.tobytes()
        for i, v in enumerate(vals):
            assert raw[576 * i:576 * (i + 1)] == ser(fn(v, vals[len(vals) - 1 - i])), (op, i)


def test_host_simulation_detects_cross_thread_hazards():
    """Regression guard for a race that once slipped through (a spill/fill between the barrier
    and the cross-thread read of a tree reduction): the host simulation aborts when a cell
    written after the last barrier is read by another thread.  Inject exactly that into a
    correct program and expect the simulator process to die."""
    import subprocess
    import sys
    import textwrap
    code = textwrap.dedent('''
        import sys
        import numpy as np
        sys.path[:0] = %r
        import hostsim
        from bls_b200.programs import curve
        from bls_b200.vm import isa
        asm = curve.build_sum_pass1(True)().assemble(18)
        rows = asm.code.tolist()
        k = next(i for i, r in enumerate(rows) if (r[0] & 0xff) == isa.OPCODE["XMOV2"])
        if sys.argv[1] == "inject":      # a write to the source cell between barrier and read
            rows.insert(k, [isa.OPCODE["MOV2"], rows[k][2], rows[k][2], 0])
            if asm.epilogue_start > k: asm.epilogue_start += 1
        asm.code = np.array(rows, dtype=np.uint16)
        pts = np.zeros(192 * 4, dtype=np.uint8)
        raw = np.zeros(3 * 6 * 1 * 16, dtype=np.uint8)
        hostsim.run(asm, {0: pts, 1: raw}, {0: 192, 1: 1}, 4, n_blocks=1, nt=128, paired=(sys.argv[2] == "paired"))
        print("finished")
    ''') % ([os.path.join(ROOT, d) for d in ('python-bls_b200', 'tests', 'oracle')],)
    for mode in ("single", "paired"):
        ok = subprocess.run([sys.executable, "-c", code, "clean", mode], capture_output=True, text=True)
        assert ok.returncode == 0 and "finished" in ok.stdout, ok.stderr[-500:]
        bad = subprocess.run([sys.executable, "-c", code, "inject", mode], capture_output=True, text=True)
        assert bad.returncode != 0 and "finished" not in bad.stdout


@pytest.mark.parametrize("g2", [False, True])
def test_windowed_scalar_mul_program(g2):
    """the two-bit-window ladder (programs/curve.py: scalar_mul) on the host build of the device
    code: generator, the order-3 point (0, 2) of E(Fq) (3P = infinity, 2P = -P), infinity, and
    scalars whose two-bit digits cover 0..3 in every position class"""
    G = O.G2 if g2 else O.G1
    w = 192 if g2 else 96

    def ser(p):
        if p[2]:
            return bytes(w)
        if g2:
            return b"".join(c.to_bytes(48, "big") for c in (p[0][0], p[0][1], p[1][0], p[1][1]))
        return p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big")

    pts = [G, G, G, G, (G[0], G[1], True)]
    ks = [0, 1, 0x1b, (0xe4 << 248) | 0x93, 12345]
    if not g2:
        pts += [(0, 2, False)] * 4
        ks += [1, 2, 3, 0x7d]
    asm = curve.build_scalar_mul(g2)().assemble(18, n_cold=4096, n_tmem=21)
    P = np.frombuffer(b"".join(ser((p[0], p[1], p[2] if len(p) > 2 else False)) for p in pts), dtype=np.uint8).copy()
    S = np.frombuffer(b"".join(k.to_bytes(32, "big") for k in ks), dtype=np.uint8).copy()
    out = np.zeros(w * len(pts), dtype=np.uint8)
    hostsim.run(asm, {0: P, 1: S, 2: out}, {0: w, 1: 32, 2: w}, len(pts), n_blocks=3, nt=3)
    raw = out.tobytes()
    for i, (p, k) in enumerate(zip(pts, ks)):
        base = (p[0], p[1], p[2] if len(p) > 2 else False)
        want = O.to_aff(O.jac_mul(k, O.to_jac(base))) if not base[2] else (0, 0, True)
        assert raw[w * i:w * (i + 1)] == ser(want), (g2, i, k)


@pytest.mark.parametrize("g2", [False, True])
def test_bucket_scale_program(g2):
    """(digit << 11 w) * P of the multi-scalar multiplication's bucket sums (programs/curve.py: build_bucket_scale):
    digits 0, 1, 2047 and mixed, windows 0, 1, 23, mixed windows inside one block of lanes (the skip regions must be
    no-ops for the lanes that do not take them), infinity and the order-3 point"""
    G = O.G2 if g2 else O.G1
    w = 192 if g2 else 96

    def ser(p):
        if p[2]:
            return bytes(w)
        if g2:
            return b"".join(c.to_bytes(48, "big") for c in (p[0][0], p[0][1], p[1][0], p[1][1]))
        return p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big")

    inf = (G[0], G[1], True)
    base = (G[0], G[1], False)
    cases = [(base, 1, 0), (base, 0, 5), (base, 2047, 0), (base, 1365, 23), (base, 3, 1), (inf, 77, 3), (base, 1024, 11),
             (O.aff_mul(7, G), 682, 22), (base, 1, 23)]
    if not g2:
        cases += [((0, 2, False), 1, 0), ((0, 2, False), 2, 1), ((0, 2, False), 3, 2)]
    asm = curve.build_bucket_scale(g2)().assemble(6, n_cold=4096, n_tmem=7)
    P = np.frombuffer(b"".join(ser(p) for p, _, _ in cases), dtype=np.uint8).copy()
    rec = bytearray()
    for _, d, win in cases:
        rec += d.to_bytes(2, "big") + bytes(1 if win > k else 0 for k in range(curve.MSM_W - 1)) + bytes(32 - 2 - (curve.MSM_W - 1))
    S = np.frombuffer(bytes(rec), dtype=np.uint8).copy()
    out = np.zeros(w * len(cases), dtype=np.uint8)
    hostsim.run(asm, {0: P, 1: S, 2: out}, {0: w, 1: 32, 2: w}, len(cases), n_blocks=3, nt=4)
    raw = out.tobytes()
    for i, (p, d, win) in enumerate(cases):
        k = d << (curve.MSM_C * win)
        want = O.to_aff(O.jac_mul(k, O.to_jac(p))) if not p[2] else (0, 0, True)
        assert raw[w * i:w * (i + 1)] == ser(want), (g2, i, d, win)


def test_instruction_fusion_patterns():
    """builder.fuse_pairs: every post-operation (z+c, z-c, c-z, xi*z, 2z) behind every primary, on the
    host build of the device code against plain modular arithmetic; and the cases that must NOT fuse
    (intermediate used twice, persistent destination, producer and consumer in different sections)"""
    from bls_b200.vm import isa
    from bls_b200.vm.builder import Program, fuse_pairs

    def f2(a, b, op):
        if op == "mul":
            return ((a[0] * b[0] - a[1] * b[1]) % O.Q, (a[0] * b[1] + a[1] * b[0]) % O.Q)
        if op == "add":
            return ((a[0] + b[0]) % O.Q, (a[1] + b[1]) % O.Q)
        if op == "sub":
            return ((a[0] - b[0]) % O.Q, (a[1] - b[1]) % O.Q)
        raise ValueError(op)

    xi = lambda a: ((a[0] - a[1]) % O.Q, (a[0] + a[1]) % O.Q)
    prog = Program("fusion")
    prog.begin_body()
    x = prog.load2_be48(0, 0)
    y = prog.load2_be48(0, 96)
    c = prog.load2_be48(0, 192)
    outs = [
        (x * y) - c, c - (x * y), (x * y) + c, (x * y).mul_xi(), (x * y).dbl(),
        x.sqr() - c, c - x.sqr(), x.sqr() + c, x.sqr().mul_xi(), x.sqr().dbl(),
        x.mul_xi() + c, (x - y) + c, (x - y).mul_xi(), (x + y) - c, c - (x + y), x.mul_xi() - c,
    ]
    t = x * y                       # used twice: must stay a separate instruction
    outs += [t + c, t - c]
    for k, v in enumerate(outs):
        prog.store2_be48(1, 96 * k, v)
    ops, _ = fuse_pairs(prog)
    n_post = sum(1 for o in ops if o.post)
    assert n_post == 16, n_post
    assert sum(1 for o in ops if o.name == "MUL2") == 6 and sum(1 for o in ops if o.name == "MUL2" and o.post) == 5
    rng = np.random.default_rng(99)
    vals = [tuple(int.from_bytes(rng.bytes(48), "big") % O.Q for _ in range(2)) for _ in range(3)]
    vals[2] = (O.Q - 1, 0)          # an edge operand
    X, Y, C = vals
    xy, xx = f2(X, Y, "mul"), f2(X, X, "mul")
    dbl = lambda a: f2(a, a, "add")
    want = [f2(xy, C, "sub"), f2(C, xy, "sub"), f2(xy, C, "add"), xi(xy), dbl(xy),
            f2(xx, C, "sub"), f2(C, xx, "sub"), f2(xx, C, "add"), xi(xx), dbl(xx),
            f2(xi(X), C, "add"), f2(f2(X, Y, "sub"), C, "add"), xi(f2(X, Y, "sub")), f2(f2(X, Y, "add"), C, "sub"),
            f2(C, f2(X, Y, "add"), "sub"), f2(xi(X), C, "sub"), f2(xy, C, "add"), f2(xy, C, "sub")]
    inp = np.frombuffer(b"".join(v.to_bytes(48, "big") for p in (X, Y, C) for v in p), dtype=np.uint8).copy()
    for shape in ((18, 21), (6, 7), (4, 2)):
        asm = prog.assemble(shape[0], n_cold=4096, n_tmem=shape[1])
        out = np.zeros(96 * len(outs), dtype=np.uint8)
        hostsim.run(asm, {0: np.tile(inp, 3), 1: np.tile(out, 3)}, {0: 288, 1: 96 * len(outs)}, 3, n_blocks=1, nt=3)
        got = hostsim.run(asm, {0: inp, 1: out}, {0: 288, 1: 96 * len(outs)}, 1, n_blocks=1, nt=1)[1].tobytes()
        for k, w in enumerate(want):
            assert got[96 * k:96 * (k + 1)] == w[0].to_bytes(48, "big") + w[1].to_bytes(48, "big"), (shape, k)
    # no fusion across a section boundary or into / out of a persistent variable
    prog2 = Program("nofuse")
    acc = prog2.var2(prog2.const2((0, 0)))
    prog2.begin_body()
    a = prog2.load2_be48(0, 0)
    s = a.sqr()
    prog2.assign(acc, acc + s)      # acc + s: ADD2 producer feeds a MOV2 (not add-like); nothing to fuse with s? (s feeds ADD2)
    prog2.begin_epilogue()
    prog2.store2_be48(1, 0, acc.dbl())
    ops2, marks2 = fuse_pairs(prog2)
    assert all(o.d is not acc or not o.post for o in ops2)
    assert marks2["body"] is not None and marks2["epilogue"] is not None
    assert ops2[marks2["epilogue"]].name in ("DBL2", "STBE48")


def test_final_exp_check_program():
    """final_exp_check (the boolean, cubed final exponentiation of aggregate verification) on the host
    build: miller(P, Q) * miller(-P, Q) exponentiates to one, miller(P, Q)^2 does not, and one does"""
    g = load_golden("pairing_kat.json")
    c = g["pairs"][1]
    P, Q = _pq([c])
    negP = P.copy()
    y = (-int.from_bytes(bytes(P[48:96]), "big")) % O.Q
    negP[48:96] = np.frombuffer(y.to_bytes(48, "big"), dtype=np.uint8)
    ml = pairing.build_miller_only().assemble(18, n_cold=4096, n_tmem=21)
    f = np.zeros(576 * 2, dtype=np.uint8)
    hostsim.run(ml, {0: np.concatenate([P, negP]), 1: np.concatenate([Q, Q]), 2: f}, {0: 96, 1: 192, 2: 576}, 2,
                n_blocks=1, nt=2)
    mul = fieldops.build_field_op(12, "mul")().assemble(18, n_cold=4096, n_tmem=21)
    a = np.concatenate([f[:576], f[:576]])
    b = np.concatenate([f[576:], f[:576]])
    prod = np.zeros(576 * 2, dtype=np.uint8)
    hostsim.run(mul, {0: a, 1: b, 2: prod}, {0: 576, 1: 576, 2: 576}, 2, n_blocks=1, nt=2)
    one = np.frombuffer((1).to_bytes(48, "big") + bytes(528), dtype=np.uint8)
    chk = pairing.build_final_exp_check().assemble(6, n_cold=4096, n_tmem=7)
    inp = np.concatenate([prod, one]).copy()
    ok = np.full(3, 9, dtype=np.uint8)
    hostsim.run(chk, {0: inp, 1: ok}, {0: 576, 1: 1}, 3, n_blocks=1, nt=3)
    assert list(ok) == [1, 0, 1]


def test_legendre_symbol_instruction():
    """FSQR1 (binary Jacobi algorithm on the limbs, csrc/fp.cuh fp_is_square) against Euler's
    criterion: 0, 1, q - 1, small values, powers of two (whole-limb shifts), squares and their
    non-residue multiples, inputs >= q on the wire (reduced on load)"""
    import random
    from bls_b200.vm.builder import Program
    rnd = random.Random(11)
    Q = O.Q
    vals = [0, 1, 2, 3, 4, 5, Q - 1, Q - 2, Q, Q + 4, 1 << 32, 1 << 64, 1 << 352, 3 << 320, (1 << 383) - 1]
    for _ in range(40):
        r = rnd.randrange(1, Q)
        vals += [r, r * r % Q, (Q - 1) * (r * r % Q) % Q, r >> rnd.randrange(380)]
    prog = Program("fsqr1_test")
    prog.begin_body()
    prog.store_flag(1, 0, prog.load1_be48(0, 0).is_square())
    asm = prog.assemble(6, n_cold=64, n_tmem=0)
    a = np.frombuffer(b"".join(v.to_bytes(48, "big") for v in vals), dtype=np.uint8).copy()
    out = np.full(len(vals), 9, dtype=np.uint8)
    hostsim.run(asm, {0: a, 1: out}, {0: 48, 1: 1}, len(vals), n_blocks=1, nt=8)
    want = [int(pow(v % Q, (Q - 1) // 2, Q) == 1) for v in vals]
    assert list(out) == want
    assert 0 < sum(want) < len(want)


def test_inversion_instruction():
    """INV1 (binary almost-inverse on the limbs + two Montgomery products, csrc/fp.cuh fp_inv) against
    pow(x, q - 2, q): 0 -> 0, 1 (fewest shift rounds: the k < 385 fix-up), q - 1, powers of two
    (whole-limb shifts), small and random values, inputs >= q on the wire"""
    import random
    from bls_b200.vm.builder import Program
    rnd = random.Random(12)
    Q = O.Q
    vals = [0, 1, 2, 3, Q - 1, Q - 2, Q, Q + 1, 1 << 32, 1 << 64, 1 << 352, 3 << 320, (1 << 383) - 1, (Q + 1) // 2]
    for _ in range(40):
        r = rnd.randrange(1, Q)
        vals += [r, r >> rnd.randrange(380), (r << 40) % Q]
    prog = Program("inv1_test")
    prog.begin_body()
    x = prog.load1_be48(0, 0)
    prog.store1_be48(1, 0, x.inv())
    prog.store1_be48(1, 48, x.inv() * x)
    asm = prog.assemble(6, n_cold=64, n_tmem=0)
    a = np.frombuffer(b"".join(v.to_bytes(48, "big") for v in vals), dtype=np.uint8).copy()
    out = np.zeros(96 * len(vals), dtype=np.uint8)
    hostsim.run(asm, {0: a, 1: out}, {0: 48, 1: 96}, len(vals), n_blocks=1, nt=8)
    raw = out.tobytes()
    for i, v in enumerate(vals):
        want = pow(v % Q, Q - 2, Q)
        assert int.from_bytes(raw[96 * i:96 * i + 48], "big") == want, (i, hex(v))
        assert int.from_bytes(raw[96 * i + 48:96 * i + 96], "big") == (0 if v % Q == 0 else 1), i


def test_compressed_decompression_branches():
    """tower.decompress_many on the host simulation: the general branch, the z2 = 0 branch and the
    all-zero case (x = 1), two items sharing one inversion, against the formulas evaluated with the
    oracle's Fq2 arithmetic (inputs need not lie in the subgroup for this check)"""
    import random
    from bls_b200.vm.builder import Program
    from bls_b200.programs.tower import CompressedCyc, decompress_many, fp_inverter
    rnd = random.Random(21)
    Q = O.Q
    r2 = lambda: (rnd.randrange(Q), rnd.randrange(Q))
    zero = (0, 0)
    rows = [(r2(), r2(), r2(), r2(), r2(), r2(), r2(), r2()),           # both general
            (zero, r2(), r2(), r2(), r2(), r2(), r2(), r2()),           # first item: z2 = 0
            (r2(), r2(), r2(), r2(), zero, zero, zero, zero),           # second item: x = 1
            (zero, zero, zero, zero, zero, r2(), r2(), r2())]
    prog = Program("decompress_test")
    prog.begin_body()
    vals = [prog.load2_be48(0, 96 * k) for k in range(8)]
    outs = decompress_many(prog, [CompressedCyc(*vals[:4]), CompressedCyc(*vals[4:])], fp_inverter(prog))
    for j, f in enumerate(outs):
        prog.store2_be48(1, 192 * j, f.c0.a0)
        prog.store2_be48(1, 192 * j + 96, f.c1.a1)
    asm = prog.assemble(6, n_cold=256, n_tmem=5)
    ser = lambda a: a[0].to_bytes(48, "big") + a[1].to_bytes(48, "big")
    inp = np.frombuffer(b"".join(ser(v) for row in rows for v in row), dtype=np.uint8).copy()
    out = np.zeros(384 * len(rows), dtype=np.uint8)
    hostsim.run(asm, {0: inp, 1: out}, {0: 768, 1: 384}, len(rows), n_blocks=1, nt=4)
    raw = out.tobytes()
    mul, add, sub, xi, k = O.f2_mul, O.f2_add, O.f2_sub, O.f2_mul_xi, O.f2_scale

    def want(z2, z3, z4, z5):
        if z2 != zero:
            num, den = sub(add(k(mul(z4, z4), 3), xi(mul(z5, z5))), k(z3, 2)), k(z2, 4)
        else:
            num, den = k(mul(z4, z5), 2), z3
        z1 = mul(num, O.f2_inv(den)) if den != zero else zero
        z0 = add(xi(sub(add(k(mul(z1, z1), 2), mul(z2, z5)), k(mul(z3, z4), 3))), (1, 0))
        return z0, z1
    for i, row in enumerate(rows):
        for j in range(2):
            z0, z1 = want(*row[4 * j:4 * j + 4])
            got = raw[384 * i + 192 * j:384 * i + 192 * j + 192]
            assert got == ser(z0) + ser(z1), (i, j)


def test_miller_squaring_and_compressed_power_against_oracle():
    """F12.sqr_x2 (Chung-Hasan SQR3 over Fq4) == 2 f^2 on arbitrary f, and pairing._pow_x_compressed
    (Karabina chain, shared decompression, top bits on Granger-Scott squarings) == m^|x| on elements
    of the cyclotomic subgroup, both against the oracle's plain Fq12 arithmetic"""
    import random
    from bls_b200.vm.builder import Program
    from bls_b200.programs.tower import F12, fp_inverter
    rnd = random.Random(31)
    Q = O.Q
    fs = [tuple(rnd.randrange(Q) for _ in range(12)) for _ in range(3)]
    ms = []
    for f in fs:
        t = O.f12_mul(O.f12_frob(f, 6), O.f12_inv(f))
        ms.append(O.f12_mul(O.f12_frob(t, 2), t))
    ser = lambda e: b"".join(int(c).to_bytes(48, "big") for c in e)
    prog = Program("sqr_pow_test")
    prog.begin_body()
    f = F12.from_coeffs([prog.load2_be48(0, 96 * k) for k in range(6)])
    m = F12.from_coeffs([prog.load2_be48(1, 96 * k) for k in range(6)])
    pairing.store_f12(prog, 2, f.sqr_x2())
    for k, c in enumerate(pairing._pow_x_compressed(prog, m, fp_inverter(prog)).coeffs()):
        prog.store2_be48(3, 96 * k, c)
    asm = prog.assemble(6, n_cold=4096, n_tmem=7)
    a = np.frombuffer(b"".join(ser(x) for x in fs), dtype=np.uint8).copy()
    b = np.frombuffer(b"".join(ser(x) for x in ms), dtype=np.uint8).copy()
    o1 = np.zeros(576 * len(fs), dtype=np.uint8)
    o2 = np.zeros(576 * len(fs), dtype=np.uint8)
    hostsim.run(asm, {0: a, 1: b, 2: o1, 3: o2}, {0: 576, 1: 576, 2: 576, 3: 576}, len(fs), n_blocks=1, nt=4)
    for i in range(len(fs)):
        assert o1.tobytes()[576 * i:576 * (i + 1)] == ser(O.f12_scale(O.f12_mul(fs[i], fs[i]), 2)), i
        assert o2.tobytes()[576 * i:576 * (i + 1)] == ser(O.f12_pow(ms[i], pairing.X_ABS)), i


# ---- single-function parity programs (programs/extras.py): the reference's golden vectors for functions the hot
# programs only use inside larger computations, replayed on the device code --------------------------------------
def _u8(b):
    return np.frombuffer(bytes(b), dtype=np.uint8).copy()


def test_sw_encode_program_all_golden_vectors_incl_edge_branches():
    """ec.py:449-507 on all 29 golden vectors from the live reference, among them the branch no hash ever
    reaches: t = 0 -> infinity (the Fq2 analogue of tests.py:104)"""
    from bls_b200.programs import extras
    g = load_golden("hash_kat.json")
    cases = list(g["sw_encode"])
    assert any(c["out"]["inf"] for c in cases)
    # w0 = t^2 + b + 1 = 0 (-> generator, ec.py:466-470) has NO solution over Fq2: -(5 + 4u) has norm 41, a
    # non-residue mod q, so it is not a square -- the branch exists in the reference and in the program but no
    # input reaches it (the reference's own KATs for it, tests.py:106-107, are for the G1 map over Fq)
    assert pow(41, (O.Q - 1) // 2, O.Q) == O.Q - 1
    with pytest.raises(ValueError):
        O.f2_sqrt(((-5) % O.Q, (-4) % O.Q))
    T = _u8(b"".join(bytes.fromhex(c["t"]) for c in cases))
    for ns, ntm in ((18, 21), (6, 7)):
        asm = extras.build_sw_encode().assemble(ns, n_cold=4096, n_tmem=ntm)
        out = np.full(192 * len(cases), 7, dtype=np.uint8)
        hostsim.run(asm, {0: T, 1: out}, {0: 96, 1: 192}, len(cases), n_blocks=2, nt=4)
        raw = out.tobytes()
        for i, c in enumerate(cases):
            want = bytes(192) if c["out"]["inf"] else bytes.fromhex(c["out"]["x"] + c["out"]["y"])
            assert raw[192 * i:192 * (i + 1)] == want, (ns, i)


def test_frobenius_pow_sqrt_programs_golden():
    """fields_t.py:344-364 (pow, qi_pow) and fields.py:199-205, 463-482 (modsqrt): the 20 + 3 + 16 golden cases"""
    from bls_b200.programs import extras
    g = load_golden("field_kat.json")
    seen = {"frob": 0, "pow": 0, "sqrt": 0}
    for c in g["cases"]:
        level, op = c["level"], c["op"]
        if op not in seen:
            continue
        w = 48 * level
        a = bytes.fromhex(c["a"]) if len(c["a"]) == 2 * w else None
        if a is None:                                  # operands stored by name
            a = b"".join(int(x).to_bytes(48, "big") for x in g["operands"][str(level)][c["a"]])
        a = _u8(a)
        if op == "frob":
            asm = extras.build_frob(level, c["i"])().assemble(18, n_cold=64, n_tmem=21)
            out = np.zeros(w, dtype=np.uint8)
            hostsim.run(asm, {0: a, 1: out}, {0: w, 1: w}, 1, nt=2)
            assert out.tobytes().hex() == c["out"], (level, c["i"])
        elif op == "pow":
            asm = extras.build_pow(level)().assemble(18, n_cold=64, n_tmem=21)
            e = int(c["e"], 16)
            out = np.zeros(2 * w, dtype=np.uint8)
            # second item: exponent 0 -> one; same base
            hostsim.run(asm, {0: np.concatenate([a, a]), 1: _u8(e.to_bytes(48, "big") + bytes(48)), 2: out},
                        {0: w, 1: 48, 2: w}, 2, nt=2)
            assert out.tobytes()[:w].hex() == c["out"], level
            assert out.tobytes()[w:] == (1).to_bytes(48, "big") + bytes(w - 48)
        else:
            asm = extras.build_sqrt(level)().assemble(18, n_cold=64, n_tmem=21)
            out = np.full(w, 9, dtype=np.uint8)
            ok = np.full(1, 9, dtype=np.uint8)
            hostsim.run(asm, {0: a, 1: out, 2: ok}, {0: w, 1: w, 2: 1}, 1, nt=2)
            if c["out"] is None:
                assert ok[0] == 0 and not out.any(), level
            else:
                assert ok[0] == 1 and out.tobytes().hex() == c["out"], level
        seen[op] += 1
    assert seen == {"frob": 20, "pow": 3, "sqrt": 16}
    # Fq2.modsqrt on a real element (a1 = 0) goes through the Fq root (fields.py:467-469); zero -> zero
    asm = extras.build_sqrt(2)().assemble(18, n_cold=64, n_tmem=21)
    sq = pow(12345, 2, O.Q)
    nonres = next(x for x in range(2, 50) if pow(x, (O.Q - 1) // 2, O.Q) != 1)
    ins = [(sq, 0), (nonres, 0), (0, 0)]
    a = _u8(b"".join(x.to_bytes(48, "big") + y.to_bytes(48, "big") for x, y in ins))
    out = np.full(96 * 3, 9, dtype=np.uint8)
    ok = np.full(3, 9, dtype=np.uint8)
    hostsim.run(asm, {0: a, 1: out, 2: ok}, {0: 96, 1: 96, 2: 1}, 3, nt=4)
    assert list(ok) == [1, 0, 1]
    assert out.tobytes()[:96] == O.fq_sqrt(sq).to_bytes(48, "big") + bytes(48) and not out[96:].any()


def test_twist_maps_and_psi_programs():
    """fields_t.py:936-943, 1018-1031, ec.py:402-444 against the oracle's Fq12 arithmetic and the golden psi vectors"""
    from bls_b200.programs import extras

    def f12b(t):
        return b"".join(int(c).to_bytes(48, "big") for c in t)
    g = load_golden("hash_kat.json")
    pts = [O.aff_mul(k, O.G2) for k in (1, 7, 1234567)]
    P = _u8(b"".join(b"".join(c.to_bytes(48, "big") for c in (p[0][0], p[0][1], p[1][0], p[1][1])) for p in pts))
    out = np.zeros(1152 * len(pts), dtype=np.uint8)
    hostsim.run(extras.build_untwist().assemble(6, n_cold=64, n_tmem=7), {0: P, 1: out}, {0: 192, 1: 1152}, len(pts), nt=2)
    for i, p in enumerate(pts):
        ux, uy, _ = O.untwist((p[0], p[1], False))
        assert out.tobytes()[1152 * i:1152 * (i + 1)] == f12b(ux) + f12b(uy), i
    # twist of arbitrary Fq12 coordinates, and twist(untwist(P)) = P embedded
    import random
    rnd = random.Random(77)
    xs = [tuple(rnd.randrange(O.Q) for _ in range(12)) for _ in range(4)]
    inp = np.concatenate([_u8(f12b(xs[0]) + f12b(xs[1]) + f12b(xs[2]) + f12b(xs[3])), out[:1152]])
    res = np.zeros(1152 * 3, dtype=np.uint8)
    hostsim.run(extras.build_twist12().assemble(18, n_cold=64, n_tmem=0), {0: inp, 1: res}, {0: 1152, 1: 1152}, 3, nt=2)
    for i in range(2):
        tx, ty, _ = O.twist((xs[2 * i], xs[2 * i + 1], False))
        assert res.tobytes()[1152 * i:1152 * (i + 1)] == f12b(tx) + f12b(ty), i
    p = pts[0]
    assert res.tobytes()[2304:] == f12b(tuple(p[0]) + (0,) * 10) + f12b(tuple(p[1]) + (0,) * 10)
    # psi: golden vectors from the live reference + the oracle on more points
    cases = [(bytes.fromhex(c["p"]["x"] + c["p"]["y"]), bytes.fromhex(c["out"]["x"] + c["out"]["y"])) for c in g["psi"]]
    for p in pts[1:]:
        q = O.psi((p[0], p[1], False))
        cases.append((b"".join(c.to_bytes(48, "big") for c in (p[0][0], p[0][1], p[1][0], p[1][1])),
                      b"".join(c.to_bytes(48, "big") for c in (q[0][0], q[0][1], q[1][0], q[1][1]))))
    out = np.zeros(192 * len(cases), dtype=np.uint8)
    hostsim.run(extras.build_psi().assemble(6, n_cold=64, n_tmem=7), {0: _u8(b"".join(c[0] for c in cases)), 1: out},
                {0: 192, 1: 192}, len(cases), nt=4)
    for i, c in enumerate(cases):
        assert out.tobytes()[192 * i:192 * (i + 1)] == c[1], i
