"""GPU parity: batched field operations through the C ABI vs the reference's known answers
(bls_py/tdata.py via tests.py:434-1036, replayed in tests/golden/field_kat.json) and vs the
oracle on seeded random operands."""
import random

import numpy as np
import pytest

import bls_oracle as O
from conftest import load_golden, unhex_elems

pytestmark = pytest.mark.gpu
Q = O.Q


def ser(elems):
    return b"".join(int(c).to_bytes(48, "big") for c in elems)


def _operand(g, ref, level):
    lv, idx = ref
    e = unhex_elems(g["operands"][str(lv)][idx])
    return e + (0,) * (level - len(e))


def test_tdata_known_answers():
    from bls_b200 import engine
    g = load_golden("field_kat.json")
    groups = {}
    for c in g["cases"]:
        if c["op"] in ("add", "sub", "mul", "sqr", "neg", "inv") and isinstance(c["a"], list):
            groups.setdefault((c["level"], c["op"]), []).append(c)
    assert len(groups) == 24
    total = 0
    for (level, op), cases in groups.items():
        a = b"".join(ser(_operand(g, c["a"], level)) for c in cases)
        b = b"".join(ser(_operand(g, c["b"], level)) for c in cases) if op in ("add", "sub", "mul") else None
        out = engine.field_op(level, op, a, b).tobytes()
        w = 48 * level
        for i, c in enumerate(cases):
            assert out[i * w:(i + 1) * w].hex() == c["out"], (level, op, i)
            total += 1
    assert total >= 300


@pytest.mark.parametrize("level", [1, 2, 6, 12])
def test_random_vs_oracle(level):
    from bls_b200 import engine
    rnd = random.Random(1000 + level)
    n = 300          # more than one CTA, ragged
    a = [tuple(rnd.randrange(Q) for _ in range(level)) for _ in range(n)]
    b = [tuple(rnd.randrange(Q) for _ in range(level)) for _ in range(n)]
    # edge operands: 0, 1, q-1
    a[0] = (0,) * level
    a[1] = (1,) + (0,) * (level - 1)
    a[2] = (Q - 1,) * level
    b[2] = (Q - 1,) * level
    mul = {1: lambda x, y: (x[0] * y[0] % Q,), 2: O.f2_mul, 6: O.f6_mul, 12: O.f12_mul}[level]
    inv = {1: lambda x: (O.fq_inv(x[0]),), 2: O.f2_inv, 6: O.f6_inv, 12: O.f12_inv}[level]
    A, B = b"".join(map(ser, a)), b"".join(map(ser, b))
    w = 48 * level
    got = engine.field_op(level, "mul", A, B).tobytes()
    for i in range(n):
        assert got[i * w:(i + 1) * w] == ser(mul(a[i], b[i])), i
    got = engine.field_op(level, "inv", A[:40 * w]).tobytes()
    for i in range(3, 40):
        assert got[i * w:(i + 1) * w] == ser(inv(a[i])), i
    assert got[:w] == bytes(w)                       # inverse of 0 is 0 (fields_t.py:47-55)
    got = engine.field_op(level, "sub", A, B).tobytes()
    for i in range(n):
        assert got[i * w:(i + 1) * w] == ser(tuple((x - y) % Q for x, y in zip(a[i], b[i])))


def test_inputs_are_reduced_mod_q():
    """Fq(Q, int) reduces (bls_py/fields.py:59-61); so do the loads"""
    from bls_b200 import engine
    big = (1 << 384) - 1
    out = engine.field_op(1, "add", big.to_bytes(48, "big"), (Q + 5).to_bytes(48, "big")).tobytes()
    assert int.from_bytes(out, "big") == (big + Q + 5) % Q


def test_empty_batch():
    from bls_b200 import engine
    assert engine.field_op(12, "mul", b"", b"").size == 0


def test_inversion_edge_values():
    """the binary almost-inverse behind every inversion (csrc/fp.cuh fp_inv): values with the fewest
    and the most shift rounds, whole-limb shifts, lanes of one warp finishing at different rounds"""
    from bls_b200 import engine
    rnd = random.Random(77)
    vals = [0, 1, 2, 3, Q - 1, Q - 2, (Q + 1) // 2, 1 << 32, 1 << 64, 1 << 352, 3 << 320, (1 << 380) + 1]
    for _ in range(500):
        r = rnd.randrange(1, Q)
        vals += [r, r >> rnd.randrange(380)]
    got = engine.field_op(1, "inv", b"".join(v.to_bytes(48, "big") for v in vals)).tobytes()
    for i, v in enumerate(vals):
        assert int.from_bytes(got[48 * i:48 * i + 48], "big") == pow(v, Q - 2, Q), (i, hex(v))


def test_legendre_symbol_edge_values():
    """FSQR1 (binary Jacobi algorithm, csrc/fp.cuh fp_is_square) vs Euler's criterion"""
    from bls_b200 import engine
    rnd = random.Random(78)
    vals = [0, 1, 2, 3, 4, 5, Q - 1, Q - 2, Q, Q + 4, 1 << 32, 1 << 64, 1 << 352, 3 << 320, (1 << 383) - 1]
    for _ in range(400):
        r = rnd.randrange(1, Q)
        vals += [r, r * r % Q, (Q - 1) * (r * r % Q) % Q, r >> rnd.randrange(380)]
    n = len(vals)
    a = engine.DeviceBuffer(48 * n).upload(np.frombuffer(b"".join(v.to_bytes(48, "big") for v in vals), dtype=np.uint8))
    out = engine.DeviceBuffer(n)
    engine.run_program_dev("f1_is_square", n, [a, out], [48, 1])
    got = out.download()
    want = [int(pow(v % Q, (Q - 1) // 2, Q) == 1) for v in vals]
    assert list(got[:n]) == want
