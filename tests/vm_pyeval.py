"""Reference evaluator for *virtual* VM programs (before cell assignment) on Python ints.

TEST INFRASTRUCTURE: used by the CPU tests to check the formulas that
``bls_b200.programs`` generates against the oracle, independently of the allocator and of
the CUDA op bodies.  The product never imports this.
"""
from bls_b200.vm.builder import Half, Val, Flag, Q

HALF_Q = (Q - 1) // 2


def run(prog, bufs, n_items, n_threads=1, strides=None, out_bufs=None):
    """bufs: {id: bytearray}; strides: {id: item stride in bytes}.  Executes prologue once,
    the body ceil(n_items / n_threads) times, the epilogue once; returns bufs."""
    strides = strides or {}
    body0 = prog.section_marks["body"]
    epi0 = prog.section_marks["epilogue"]
    if epi0 is None:
        epi0 = len(prog.ops)
    state = [dict() for _ in range(n_threads)]      # value id -> int | [c0, c1] | bool
    raw = {}                                        # internal SoA buffers: (buf, elem, item) -> (c0, c1)

    def get(st, x):
        if isinstance(x, Half):
            return st[x.parent.id][x.half]
        return st[x.id]

    def put(st, x, v):
        if isinstance(x, Half):
            st.setdefault(x.parent.id, [None, None])[x.half] = v
        else:
            st[x.id] = v

    def rd(buf, item, off, n):
        base = item * strides.get(buf, 0) + off
        return int.from_bytes(bytes(bufs[buf][base:base + n]), "big")

    def exec_range(lo, hi, it):
        for k in range(lo, hi):
            op = prog.ops[k]
            nm = op.name
            if nm in ("SYNC", "SKIPZ", "SKIP_END"):
                continue
            if nm == "XCHG":
                for dv, sv in zip(op.d, op.a):
                    vals = [list(get(state[(t + op.b) % n_threads], sv)) for t in range(n_threads)]
                    for t in range(n_threads):
                        put(state[t], dv, vals[t])
                continue
            for t in range(n_threads):
                st = state[t]
                item_raw = it * n_threads + t
                active = item_raw < n_items
                item = min(item_raw, max(n_items - 1, 0))
                a = get(st, op.a) if isinstance(op.a, (Val, Half)) else op.a
                b = get(st, op.b) if isinstance(op.b, (Val, Half)) else op.b
                if nm == "MUL2":
                    r = [(a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q]
                elif nm == "SQR2":
                    r = [(a[0] * a[0] - a[1] * a[1]) % Q, 2 * a[0] * a[1] % Q]
                elif nm == "ADD2":
                    r = [(a[0] + b[0]) % Q, (a[1] + b[1]) % Q]
                elif nm == "SUB2":
                    r = [(a[0] - b[0]) % Q, (a[1] - b[1]) % Q]
                elif nm == "NEG2":
                    r = [-a[0] % Q, -a[1] % Q]
                elif nm == "DBL2":
                    r = [2 * a[0] % Q, 2 * a[1] % Q]
                elif nm == "MULXI2":
                    r = [(a[0] - a[1]) % Q, (a[0] + a[1]) % Q]
                elif nm == "CONJ2":
                    r = [a[0], -a[1] % Q]
                elif nm == "MOV2":
                    r = list(a)
                elif nm == "MULFP2":
                    r = [a[0] * b % Q, a[1] * b % Q]
                elif nm == "TRI2":
                    sg = 1 if op.aux else -1
                    r = [(3 * a[0] + sg * 2 * b[0]) % Q, (3 * a[1] + sg * 2 * b[1]) % Q]
                elif nm == "MUL1":
                    r = a * b % Q
                elif nm == "SQR1":
                    r = a * a % Q
                elif nm == "ADD1":
                    r = (a + b) % Q
                elif nm == "SUB1":
                    r = (a - b) % Q
                elif nm == "NEG1":
                    r = -a % Q
                elif nm == "DBL1":
                    r = 2 * a % Q
                elif nm == "MOV1":
                    r = a
                elif nm == "INV1":
                    r = pow(a, Q - 2, Q)
                elif nm == "LDC1":
                    r = prog.consts[op.a]
                elif nm == "LDC2":
                    r = [prog.consts[op.a], prog.consts[op.a + 1]]
                elif nm == "FZERO1":
                    r = a == 0
                elif nm == "FZERO2":
                    r = a[0] == 0 and a[1] == 0
                elif nm == "FGTHALF":
                    r = a > HALF_Q
                elif nm == "FSQR1":
                    r = pow(a, (Q - 1) // 2, Q) == 1
                elif nm == "FEQ1":
                    r = a == b
                elif nm == "FEQ2":
                    r = list(a) == list(b)
                elif nm == "FAND":
                    r = a and b
                elif nm == "FOR":
                    r = a or b
                elif nm == "FXOR":
                    r = a != b
                elif nm == "FNOT":
                    r = not a
                elif nm == "FSET":
                    r = bool(op.a & 1)
                elif nm == "FBIT":
                    sc = rd(op.a, item, 0, (op.aux + 1) if op.aux else 32)
                    r = bool((sc >> op.b) & 1)
                elif nm == "FLDB":
                    r = rd(op.a, item, op.b, 1) != 0
                elif nm == "FACTIVE":
                    r = active
                elif nm in ("CSEL2", "CSEL1"):
                    f = get(st, op.aux)
                    r = (list(a) if nm == "CSEL2" else a) if f else (list(b) if nm == "CSEL2" else b)
                elif nm == "LDBE48":
                    r = rd(op.a, item, 16 * op.b, 48)
                    if op.aux:
                        r &= (1 << 381) - 1
                    r %= Q
                elif nm == "LDBE32":
                    r = rd(op.a, item, 16 * op.b, 32) % Q
                elif nm == "STBE48":
                    if op.aux:
                        if t == 0:
                            bufs[op.d][16 * op.b:16 * op.b + 48] = int(a).to_bytes(48, "big")
                    elif active:
                        base = item * strides.get(op.d, 0) + 16 * op.b
                        bufs[op.d][base:base + 48] = int(a).to_bytes(48, "big")
                    continue
                elif nm == "STFLAG":
                    if op.aux:
                        if t == 0:
                            bufs[op.d][op.b] = 1 if a else 0
                    elif active:
                        bufs[op.d][item * strides.get(op.d, 0) + op.b] = 1 if a else 0
                    continue
                elif nm == "LDRAW2":
                    r = list(raw[(op.a, op.b, item)])
                elif nm == "STRAW2":
                    if active:
                        raw[(op.d, op.b, item_raw)] = tuple(a)
                    continue
                elif nm == "STRAWB2":
                    if t == 0:
                        raw[(op.d, op.b, 0)] = tuple(a)
                    continue
                elif nm == "NOP":
                    continue
                else:
                    raise NotImplementedError(nm)
                put(st, op.d, r)

    exec_range(0, body0, 0)
    iters = (n_items + n_threads - 1) // n_threads
    for it in range(iters):
        exec_range(body0, epi0, it)
    exec_range(epi0, len(prog.ops), max(iters - 1, 0))
    bufs["_raw"] = raw
    return bufs
