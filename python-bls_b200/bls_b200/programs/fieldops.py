"""Batched field operations on every tower level (parity with bls_py/tdata.py, microbenches).

Buffers: 0 = a, 1 = b (binary ops), 2 = out; elements are level * 48 bytes big-endian in the
reference's flat ZT order (bls_py/fields.py:273-278).
"""
from ..vm.builder import Program
from .tower import F6, F12, f2_inv, fp_inverter


def _load(prog, buf, level):
    if level == 1:
        return prog.load1_be48(buf, 0)
    c = [prog.load2_be48(buf, 96 * k) for k in range(level // 2)]
    if level == 2:
        return c[0]
    if level == 6:
        return F6(*c)
    return F12.from_coeffs(c)


def _store(prog, buf, level, x):
    if level == 1:
        prog.store1_be48(buf, 0, x)
    elif level == 2:
        prog.store2_be48(buf, 0, x)
    else:
        cs = [x.a0, x.a1, x.a2] if level == 6 else x.coeffs()
        for k, c in enumerate(cs):
            prog.store2_be48(buf, 96 * k, c)


def build_field_op(level, op):
    def build():
        prog = Program("f%d_%s" % (level, op))
        prog.begin_body()
        a = _load(prog, 0, level)
        if op in ("add", "sub", "mul"):
            b = _load(prog, 1, level)
            r = a + b if op == "add" else (a - b if op == "sub" else a * b)
        elif op == "sqr":
            r = a.sqr()
        elif op == "neg":
            r = -a
        elif op == "inv":
            fp_inv = fp_inverter(prog)
            r = fp_inv(a) if level == 1 else (f2_inv(a, fp_inv) if level == 2 else a.inv(fp_inv))
        else:
            raise ValueError(op)
        _store(prog, 2, level, r)
        return prog
    return build


def build_fq2_mul_chain(n_mul):
    """microbenchmark: n_mul dependent Fq2 products per item (a <- a * b), for the
    integer-pipe roofline measurement of the MUL2 body in isolation"""
    def build():
        prog = Program("fq2_mul_chain%d" % n_mul)
        prog.begin_body()
        a = prog.load2_be48(0, 0)
        b = prog.load2_be48(1, 0)
        for _ in range(n_mul):
            a = a * b
        prog.store2_be48(2, 0, a)
        return prog
    return build


def build_is_square():
    """buffers: 0 = Fq element (48 B), 1 = one byte: 1 iff the element is a nonzero square (the
    Legendre-symbol instruction FSQR1 on its own; the product uses it inside hash-to-G2, where the
    reference's y_for_x raises for a non-residue, bls_py/ec.py:255-269)"""
    prog = Program("f1_is_square")
    prog.begin_body()
    prog.store_flag(1, 0, prog.load1_be48(0, 0).is_square())
    return prog
