"""Program builder for the b200-bls field VM.

Formulas are written in ordinary Python on symbolic values (``V2`` = an Fq2 element,
``V1`` = an Fq element, ``Flag`` = a per-thread predicate); every operator appends one
virtual instruction to the program.  ``Program.assemble`` then assigns shared-memory
cells with a Belady-style linear scan (values are single-assignment, the code is straight
line), inserting SPILL2/FILL2 moves to the global-memory cold area when the per-thread
workspace is exceeded, and emits the 8-byte instructions the CUDA kernel interprets.
"""
import numpy as np

from . import isa

Q = int("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f624"
        "1eabfffeb153ffffb9feffffffffaaab", 16)


class Val:
    __slots__ = ("prog", "id", "width")

    def __init__(self, prog, width):
        self.prog = prog
        self.width = width
        self.id = prog._new_id(self)


class V1(Val):
    """Fq value (one cell)."""
    __slots__ = ()

    def __init__(self, prog):
        super().__init__(prog, 1)

    def _bin(self, op, other):
        r = V1(self.prog)
        self.prog.emit(op, r, self, other)
        return r

    def __add__(self, o):
        return self._bin("ADD1", o)

    def __sub__(self, o):
        return self._bin("SUB1", o)

    def __mul__(self, o):
        if isinstance(o, (V2,)):
            return o * self
        return self._bin("MUL1", o)

    def __neg__(self):
        r = V1(self.prog)
        self.prog.emit("NEG1", r, self)
        return r

    def sqr(self):
        r = V1(self.prog)
        self.prog.emit("SQR1", r, self)
        return r

    def dbl(self):
        r = V1(self.prog)
        self.prog.emit("DBL1", r, self)
        return r

    def inv(self):
        """1 / self in Fq (0 -> 0)"""
        r = V1(self.prog)
        self.prog.emit("INV1", r, self)
        return r

    def is_zero(self):
        f = Flag(self.prog)
        self.prog.emit("FZERO1", f, self)
        return f

    def gt_half(self):
        f = Flag(self.prog)
        self.prog.emit("FGTHALF", f, self)
        return f

    def is_square(self):
        """nonzero quadratic residue: the Legendre symbol without an exponentiation"""
        f = Flag(self.prog)
        self.prog.emit("FSQR1", f, self)
        return f

    def eq(self, o):
        f = Flag(self.prog)
        self.prog.emit("FEQ1", f, self, o)
        return f


class Half:
    """Read (or initialising write) view of one coefficient of a V2."""
    __slots__ = ("parent", "half", "prog", "width")

    def __init__(self, parent, half):
        self.parent = parent
        self.half = half
        self.prog = parent.prog
        self.width = 1

    # a Half behaves like a V1 source
    __add__ = V1.__add__
    __sub__ = V1.__sub__
    __mul__ = V1.__mul__
    __neg__ = V1.__neg__
    _bin = V1._bin
    sqr = V1.sqr
    dbl = V1.dbl
    is_zero = V1.is_zero
    gt_half = V1.gt_half
    eq = V1.eq
    inv = V1.inv
    is_square = V1.is_square


class V2(Val):
    """Fq2 value (an even/odd cell pair)."""
    __slots__ = ()

    def __init__(self, prog):
        super().__init__(prog, 2)

    @property
    def c0(self):
        return Half(self, 0)

    @property
    def c1(self):
        return Half(self, 1)

    def _bin(self, op, other):
        r = V2(self.prog)
        self.prog.emit(op, r, self, other)
        return r

    def _un(self, op):
        r = V2(self.prog)
        self.prog.emit(op, r, self)
        return r

    def __add__(self, o):
        return self._bin("ADD2", o)

    def __sub__(self, o):
        return self._bin("SUB2", o)

    def __mul__(self, o):
        if isinstance(o, (V1, Half)):
            return self._bin("MULFP2", o)
        return self._bin("MUL2", o)

    def __neg__(self):
        return self._un("NEG2")

    def sqr(self):
        return self._un("SQR2")

    def dbl(self):
        return self._un("DBL2")

    def mul_xi(self):
        return self._un("MULXI2")

    def conj(self):
        return self._un("CONJ2")

    def is_zero(self):
        f = Flag(self.prog)
        self.prog.emit("FZERO2", f, self)
        return f

    def eq(self, o):
        f = Flag(self.prog)
        self.prog.emit("FEQ2", f, self, o)
        return f


class Flag(Val):
    __slots__ = ()

    def __init__(self, prog):
        super().__init__(prog, 0)

    def _bin(self, op, o):
        f = Flag(self.prog)
        self.prog.emit(op, f, self, o)
        return f

    def __and__(self, o):
        return self._bin("FAND", o)

    def __or__(self, o):
        return self._bin("FOR", o)

    def __xor__(self, o):
        return self._bin("FXOR", o)

    def __invert__(self):
        f = Flag(self.prog)
        self.prog.emit("FNOT", f, self)
        return f


class VOp:
    __slots__ = ("name", "d", "a", "b", "aux", "post")

    def __init__(self, name, d, a, b, aux, post=0):
        self.name, self.d, self.a, self.b, self.aux, self.post = name, d, a, b, aux, post


def _is_val(x):
    return isinstance(x, (Val, Half))


class Program:
    """A straight-line VM program in three sections (prologue / per-item body / epilogue)."""

    def __init__(self, name):
        self.name = name
        self.vals = []
        self.ops = []                 # virtual ops
        self.consts = []              # python ints (standard form)
        self._const_idx = {}
        self.section_marks = {"body": 0, "epilogue": None}
        self.n_in = 0
        self.n_out = 0
        self.keep = []                # values that must survive to the end (epilogue use)
        self.persistent = set()       # ids of multi-assignment variables

    # ---- infrastructure -------------------------------------------------------------
    def _new_id(self, v):
        self.vals.append(v)
        return len(self.vals) - 1

    def emit(self, name, d=None, a=None, b=None, aux=None):
        self.ops.append(VOp(name, d, a, b, aux))

    def begin_body(self):
        self.section_marks["body"] = len(self.ops)

    def begin_epilogue(self):
        self.section_marks["epilogue"] = len(self.ops)

    # ---- constants --------------------------------------------------------------------
    def _cidx(self, ints):
        key = tuple(int(x) % Q for x in ints)
        if key not in self._const_idx:
            self._const_idx[key] = len(self.consts)
            self.consts.extend(key)
        return self._const_idx[key]

    def const1(self, x):
        r = V1(self)
        self.emit("LDC1", r, self._cidx((x,)))
        return r

    def const2(self, x):
        r = V2(self)
        self.emit("LDC2", r, self._cidx(x))
        return r

    # ---- I/O ----------------------------------------------------------------------------
    def load1_be48(self, buf, off_bytes, mask_top=False):
        """mask_top: clear the three flag bits of byte 0 (buffer[0] & 0x1f, keys.py:31)"""
        assert off_bytes % 16 == 0
        r = V1(self)
        self.emit("LDBE48", r, buf, off_bytes // 16, aux=int(mask_top))
        return r

    def load2_be48(self, buf, off_bytes, mask_top=False):
        """two consecutive 48-byte big-endian coefficients -> Fq2"""
        r = V2(self)
        self.emit("LDBE48", Half(r, 0), buf, off_bytes // 16, aux=int(mask_top))
        self.emit("LDBE48", Half(r, 1), buf, off_bytes // 16 + 3)
        return r

    def load1_be32(self, buf, off_bytes):
        assert off_bytes % 16 == 0
        r = V1(self)
        self.emit("LDBE32", r, buf, off_bytes // 16)
        return r

    # block_only=True: only thread 0 of the CTA stores, into the record of item = CTA index
    def store1_be48(self, buf, off_bytes, v, block_only=False):
        self.emit("STBE48", buf, v, off_bytes // 16, aux=int(block_only))

    def store2_be48(self, buf, off_bytes, v, block_only=False):
        self.emit("STBE48", buf, v.c0, off_bytes // 16, aux=int(block_only))
        self.emit("STBE48", buf, v.c1, off_bytes // 16 + 3, aux=int(block_only))

    def store_flag(self, buf, off_bytes, f, block_only=False):
        self.emit("STFLAG", buf, f, off_bytes, aux=int(block_only))

    def load_raw2(self, buf, elem):
        r = V2(self)
        self.emit("LDRAW2", r, buf, elem)
        return r

    def store_raw2(self, buf, elem, v, block_only=False):
        self.emit("STRAWB2" if block_only else "STRAW2", buf, v, elem)

    # ---- misc -----------------------------------------------------------------------------
    def pack(self, a, b):
        r = V2(self)
        self.emit("MOV1", Half(r, 0), a)
        self.emit("MOV1", Half(r, 1), b)
        return r

    def mov1(self, a):
        r = V1(self)
        self.emit("MOV1", r, a)
        return r

    def sel2(self, f, a, b):
        r = V2(self)
        self.emit("CSEL2", r, a, b, aux=f)
        return r

    def sel1(self, f, a, b):
        r = V1(self)
        self.emit("CSEL1", r, a, b, aux=f)
        return r

    def flag_const(self, bit):
        f = Flag(self)
        self.emit("FSET", f, int(bit))
        return f

    def flag_bit(self, buf, bit, nbytes=32):
        """bit `bit` of the item's big-endian scalar of `nbytes` bytes in buffer `buf`"""
        f = Flag(self)
        self.emit("FBIT", f, buf, bit, aux=(nbytes - 1 if nbytes != 32 else None))
        return f

    def flag_byte(self, buf, off):
        f = Flag(self)
        self.emit("FLDB", f, buf, off)
        return f

    def flag_active(self):
        f = Flag(self)
        self.emit("FACTIVE", f)
        return f

    def var2(self, init):
        """persistent (multi-assignment) Fq2 variable: keeps one fixed cell pair for the whole
        program, e.g. an accumulator carried across item-loop iterations"""
        r = V2(self)
        self.emit("MOV2", r, init)
        self.persistent.add(r.id)
        return r

    def assign(self, var, value):
        assert var.id in self.persistent
        self.emit("MOV2", var, value)

    def tri2(self, a, b, plus):
        """3a + 2b (plus) or 3a - 2b in one instruction"""
        r = V2(self)
        self.emit("TRI2", r, a, b, aux=int(bool(plus)))
        return r

    def update_sel(self, var, f, a):
        """var = f ? a : var, in place (var keeps its cells)"""
        self.emit("CSEL2" if var.width == 2 else "CSEL1", var, a, var, aux=f)

    def skip_unless(self, f):
        """``with prog.skip_unless(f):`` -- the enclosed instructions are skipped by a warp in
        which no thread has flag f.  Purely an optimisation: the region must only change
        state through ``update_sel(..., f, ...)`` so that executing it with f clear is a
        no-op for that thread.  The allocator refuses spills inside a region."""
        prog = self

        class _Region:
            def __enter__(self_inner):
                prog.emit("SKIPZ", f, None)

            def __exit__(self_inner, *exc):
                prog.emit("SKIP_END")
                return False
        return _Region()

    def sync(self):
        self.emit("SYNC")

    def exchange(self, values, lane_off):
        """cross-thread read: returns, for every thread, the values held by thread
        (tid + lane_off) mod CTA size.  Assembled as ONE unit: every source is first made
        resident in shared memory and every destination cell is reserved (all spills, fills and
        relocations happen here), THEN barrier, the XMOV2 reads, barrier.  No memory traffic of
        any thread can fall between a barrier and the reads that depend on it."""
        outs = [V2(self) for _ in values]
        self.emit("XCHG", outs, list(values), int(lane_off))
        return outs

    # ---- assembly ---------------------------------------------------------------------------
    def assemble(self, n_slots, n_cold=1024, n_tmem=0):
        """n_slots Fq2 slots in shared memory plus n_tmem slots in Tensor Memory (slot indices
        n_slots .. n_slots + n_tmem - 1).  TMEM lanes are private to their thread, so the operands of
        cross-thread reads (XMOV2) are kept in / moved to shared-memory slots."""
        asm = _assemble(self, n_slots + n_tmem, n_cold, n_smem=n_slots)
        asm.n_slots, asm.n_tmem = n_slots, n_tmem
        return asm


class Assembled:
    """Concrete program: ``code`` is an (n, 4) uint16 array [op|aux<<8, d, a, b]."""

    def __init__(self, name, code, consts, marks, n_slots, n_cold, stats):
        self.name = name
        self.code = code
        self.consts = consts
        self.body_start, self.epilogue_start = marks
        self.n_slots = n_slots      # shared-memory slots
        self.n_tmem = 0             # Tensor Memory slots (indices after the shared ones)
        self.n_cold = n_cold
        self.stats = stats

    def const_limbs(self):
        """constant table as Montgomery-form little-endian u32 limbs, shape (n, 12)"""
        out = np.zeros((max(1, len(self.consts)), 12), dtype=np.uint32)
        for i, c in enumerate(self.consts):
            m = (c << 384) % Q
            for j in range(12):
                out[i, j] = (m >> (32 * j)) & 0xffffffff
        return out


def fuse_pairs(prog):
    """Peephole over the virtual ops: an Fq2 producer whose ONLY consumer is the add-like op that
    directly follows it becomes one instruction with a post-operation (isa.POST_*): the
    intermediate is never stored or reloaded and one fetch / decode / dispatch round is saved
    (a fifth of the pairing program's instructions).  Returns (ops, section marks)."""
    ops = prog.ops
    n = len(ops)

    def root(x):
        return x.parent if isinstance(x, Half) else x

    n_uses = {}
    for op in ops:
        operands = (list(op.d) + list(op.a)) if op.name == "XCHG" else (op.a, op.b, op.aux)
        for x in operands:
            if _is_val(x):
                n_uses[root(x).id] = n_uses.get(root(x).id, 0) + 1
    for v in prog.keep:
        n_uses[v.id] = n_uses.get(v.id, 0) + 1
    n_defs = {}
    for op in ops:
        if op.name != "XCHG" and _is_val(op.d):
            n_defs[root(op.d).id] = n_defs.get(root(op.d).id, 0) + 1
    body0, epi0 = prog.section_marks["body"], prog.section_marks["epilogue"]
    out, marks = [], {"body": None, "epilogue": None}
    depth = 0
    i = 0
    while i < n:
        if i == body0:
            marks["body"] = len(out)
        if i == epi0:
            marks["epilogue"] = len(out)
        op = ops[i]
        if op.name == "SKIPZ":
            depth += 1
        elif op.name == "SKIP_END":
            depth -= 1
        nxt = ops[i + 1] if i + 1 < n else None
        fused = None
        if (nxt is not None and depth == 0 and op.name in isa.POST_PRIMARY and op.aux is None
                and isinstance(op.d, V2) and nxt.name in isa.POST_SECONDARY and isinstance(nxt.d, V2)
                and i + 1 != body0 and i + 1 != epi0
                and op.d.id not in prog.persistent and n_defs.get(op.d.id, 0) == 1
                and n_uses.get(op.d.id, 0) == 1):
            x = op.d
            if nxt.name == "ADD2" and (nxt.a is x) != (nxt.b is x) and isinstance(nxt.a, V2) and isinstance(nxt.b, V2):
                fused = (isa.POST_ADD, nxt.b if nxt.a is x else nxt.a)
            elif nxt.name == "SUB2" and (nxt.a is x) != (nxt.b is x) and isinstance(nxt.a, V2) and isinstance(nxt.b, V2):
                fused = (isa.POST_SUB, nxt.b) if nxt.a is x else (isa.POST_RSUB, nxt.a)
            elif nxt.name == "MULXI2" and nxt.a is x:
                fused = (isa.POST_XI, None)
            elif nxt.name == "DBL2" and nxt.a is x:
                fused = (isa.POST_DBL, None)
        if fused is not None:
            out.append(VOp(op.name, nxt.d, op.a, op.b, fused[1], post=fused[0]))
            i += 2
            continue
        out.append(op)
        i += 1
    if marks["body"] is None:
        marks["body"] = len(out) if body0 is not None and body0 >= n else 0
    if epi0 is not None and marks["epilogue"] is None:
        marks["epilogue"] = len(out)
    return out, marks


def _assemble(prog, n_slots, n_cold, n_smem=None):
    if n_smem is None:
        n_smem = n_slots
    ops, section_marks = fuse_pairs(prog)
    n = len(ops)
    INF = 1 << 60

    def root(x):
        return x.parent if isinstance(x, Half) else x

    # use positions per value
    uses = {}
    for i, op in enumerate(ops):
        operands = (list(op.d) + list(op.a)) if op.name == "XCHG" else (op.d, op.a, op.b, op.aux)
        for x in operands:
            if _is_val(x):
                uses.setdefault(root(x).id, []).append(i)
    for v in prog.keep:
        uses.setdefault(v.id, []).append(n)
    use_ptr = {k: 0 for k in uses}
    # values that cross the prologue/body boundary (or are persistent variables) keep a fixed
    # cell for the whole program: the body is executed many times with the same assignment
    body0 = section_marks["body"]
    fixed = set(prog.persistent)
    for vid, lst in uses.items():
        if lst[0] < body0 <= lst[-1] and body0 > 0:
            fixed.add(vid)

    def next_use(vid, after):
        lst = uses[vid]
        p = use_ptr[vid]
        while p < len(lst) and lst[p] <= after:
            p += 1
        use_ptr[vid] = p
        return lst[p] if p < len(lst) else INF

    free_slots = list(range(n_slots - 1, -1, -1))
    slot_of = {}          # value id -> slot (resident)
    cold_of = {}          # value id -> cold index (has a valid cold copy)
    free_cold = list(range(n_cold - 1, -1, -1))
    free_flags = list(range(isa.N_FLAGS - 1, -1, -1))
    flag_of = {}
    out = []
    stats = {"spills": 0, "fills": 0, "max_slots": 0, "max_cold": 0}

    pending = []          # dead cold slots whose lines have not been discarded yet

    def emit(name, d=0, a=0, b=0, aux=0):
        # a dead cold copy rides on the next SPILL2 / FILL2 (aux = 0x80 | slot: discard that slot first) -- a
        # stand-alone DISCARD2 costs a whole fetch / decode round, and a slot that is re-spilled at once needs none
        if name == "SPILL2" and d in pending:
            pending.remove(d)                           # re-spilled: the store overwrites the lines in the L2
        if name in ("SPILL2", "FILL2") and aux == 0 and pending and not skip_stack:
            aux = 0x80 | pending.pop(0)
        out.append((isa.OPCODE[name] | (aux << 8), d, a, b))

    def flush_discards():
        while pending:
            c = pending.pop(0)
            out.append((isa.OPCODE["DISCARD2"], c, 0, 0))

    def alloc_slot(i, pinned, smem_only=False):
        cands = [x for x in free_slots if x < n_smem] if smem_only else free_slots
        if not cands and skip_stack:
            raise RuntimeError("spill needed inside a skip region (op %d)" % i)
        if cands:
            s = min(cands) if smem_only else free_slots[-1]
            free_slots.remove(s)
            stats["max_slots"] = max(stats["max_slots"], n_slots - len(free_slots))
            return s
        # evict the resident value with the farthest next use
        best, best_use = None, -1
        for vid in slot_of:
            if vid in pinned or vid in fixed or vid in region_pins:
                continue
            if smem_only and slot_of[vid] >= n_smem:
                continue
            nu = next_use(vid, i - 1)
            if nu > best_use:
                best, best_use = vid, nu
        if best is None:
            raise RuntimeError("workspace too small: all %d slots pinned at op %d (%s)"
                               % (n_slots, i, ops[i].name))
        s = slot_of.pop(best)
        if best_use < INF and best not in cold_of:
            if not free_cold:
                raise RuntimeError("cold area exhausted")
            c = free_cold.pop()
            cold_of[best] = c
            stats["max_cold"] = max(stats["max_cold"], n_cold - len(free_cold))
            emit("SPILL2", c, 2 * s)
            stats["spills"] += 1
        return s

    epi0 = section_marks["epilogue"]
    cur_op = [0]

    def release(vid):
        # fixed-cell values live for the whole body loop; in the epilogue (which runs once,
        # after the last iteration) they die like any other value
        if vid in fixed and (epi0 is None or cur_op[0] < epi0):
            return
        if vid in slot_of:
            free_slots.append(slot_of.pop(vid))
        if vid in cold_of:
            c = cold_of.pop(vid)
            free_cold.append(c)
            if vid in discarded:
                discarded.discard(vid)          # its last FILL2 already dropped the lines
            elif c < 0x80:
                pending.append(c)               # the value died in the workspace: its cold copy is dead too
                stats["discards"] = stats.get("discards", 0) + 1

    marks = [None, None]
    skip_stack = []
    region_pins = set()
    discarded = set()     # values whose cold copy the kernel has already discarded (FILL2 with aux = 1)
    for i, op in enumerate(ops):
        cur_op[0] = i
        if i == section_marks["body"]:
            flush_discards()
            emit("END")
            marks[0] = len(out)
        if i == section_marks["epilogue"]:
            flush_discards()
            emit("END")
            marks[1] = len(out)
        if op.name == "XCHG":
            if skip_stack:
                raise RuntimeError("exchange inside a skip region")
            srcs_x, dsts_x = list(op.a), list(op.d)
            hold = set(v.id for v in srcs_x)
            for v in srcs_x:
                if v.id in slot_of and slot_of[v.id] >= n_smem:
                    if v.id in fixed:
                        raise RuntimeError("exchange of a fixed-cell value that lives in tensor memory")
                    s2 = alloc_slot(i, hold, smem_only=True)
                    emit("MOV2", 2 * s2, 2 * slot_of[v.id])
                    free_slots.append(slot_of[v.id])
                    slot_of[v.id] = s2
                elif v.id not in slot_of:
                    if v.id not in cold_of:
                        raise RuntimeError("use of undefined value at op %d (XCHG)" % i)
                    s2 = alloc_slot(i, hold, smem_only=True)
                    slot_of[v.id] = s2
                    emit("FILL2", 2 * s2, cold_of[v.id])
                    stats["fills"] += 1
            for dv in dsts_x:
                slot_of[dv.id] = alloc_slot(i, hold, smem_only=True)
                hold.add(dv.id)
            emit("SYNC")
            for dv, v in zip(dsts_x, srcs_x):
                emit("XMOV2", 2 * slot_of[dv.id], 2 * slot_of[v.id], int(op.b))
            emit("SYNC")
            for v in srcs_x:
                if next_use(v.id, i) == INF:
                    release(v.id)
            for dv in dsts_x:
                if next_use(dv.id, i) == INF:
                    release(dv.id)
            continue
        if op.name == "SKIP_END":
            at = skip_stack.pop()
            w0, d, _, b = out[at]
            out[at] = (w0, d, len(out) - at - 1, b)
            region_pins.clear()
            continue
        if op.name == "SKIPZ":
            # A skipped region must not contain spills or fills (they would be skipped too).
            # Make every outside value it reads resident now, and free enough slots for the
            # values it defines, before the branch.
            end = i + 1
            while ops[end].name != "SKIP_END":
                end += 1
            inside_defs, outside_used = [], []
            for k in range(i + 1, end):
                o = ops[k]
                for x in (o.a, o.b):
                    if _is_val(x) and not isinstance(root(x), Flag):
                        vid = root(x).id
                        if uses[vid][0] <= i and vid not in outside_used:
                            outside_used.append(vid)
                if _is_val(o.d) and not isinstance(root(o.d), Flag):
                    vid = root(o.d).id
                    if uses[vid][0] <= i:
                        if vid not in outside_used:
                            outside_used.append(vid)
                    elif vid not in inside_defs:
                        inside_defs.append(vid)
            region_pins.update(outside_used)
            for vid in outside_used:
                if vid not in slot_of:
                    sl = alloc_slot(i, region_pins)
                    slot_of[vid] = sl
                    emit("FILL2", 2 * sl, cold_of[vid])
                    stats["fills"] += 1
            # peak number of simultaneously live region-defined values
            live, peak = set(), 0
            for k in range(i + 1, end):
                o = ops[k]
                if _is_val(o.d) and not isinstance(root(o.d), Flag) and root(o.d).id in inside_defs:
                    live.add(root(o.d).id)
                    peak = max(peak, len(live))
                for x in (o.a, o.b):
                    if _is_val(x) and root(x).id in live and uses[root(x).id][-1] == k:
                        live.discard(root(x).id)
            held = []
            while len(free_slots) + len(held) < peak + 1:
                held.append(alloc_slot(i, region_pins))
            free_slots.extend(held)
        fields = [op.d, op.a, op.b]
        srcs = [x for x in (op.a, op.b) + ((op.aux,) if op.post else ()) if _is_val(x) and not isinstance(root(x), Flag)]
        # STBE48/STRAW2/SPILL-like ops read their 'a'; dst-position value operands that are
        # data sources do not exist in this ISA (buffer ids are ints)
        pinned = set(root(x).id for x in srcs)
        d_is_data = _is_val(op.d) and not isinstance(root(op.d), Flag)
        partial_def = d_is_data and isinstance(op.d, Half)
        if partial_def and root(op.d).id in slot_of:
            pinned.add(root(op.d).id)
        if op.name == "XMOV2":
            raise RuntimeError("XMOV2 is emitted by Program.exchange() only")
        # make sources resident
        for x in srcs:
            vid = root(x).id
            if vid not in slot_of:
                if vid not in cold_of:
                    raise RuntimeError("use of undefined value at op %d (%s)" % (i, op.name))
                if skip_stack:
                    raise RuntimeError("fill needed inside a skip region (op %d)" % i)
                s = alloc_slot(i, pinned)
                slot_of[vid] = s
                # aux = 1: this fill serves the value's LAST use, so its cold copy is dead once it has been read --
                # the kernel then discards the lines from the L2 instead of letting them be written back to DRAM
                last = uses[vid][-1] == i and vid not in fixed and cold_of[vid] < 0x80
                if last:
                    discarded.add(vid)
                emit("FILL2", 2 * s, cold_of[vid], 0, (0x80 | cold_of[vid]) if last else 0)
                stats["fills"] += 1
                stats["fills_last"] = stats.get("fills_last", 0) + (1 if last else 0)
        # concrete source operands
        conc = [0, 0, 0]
        for k, x in enumerate(fields):
            if k == 0:
                continue
            if _is_val(x):
                r = root(x)
                if isinstance(r, Flag):
                    conc[k] = flag_of[r.id]
                else:
                    conc[k] = 2 * slot_of[r.id] + (x.half if isinstance(x, Half) else 0)
            elif x is not None:
                conc[k] = int(x)
        aux = 0
        if op.aux is not None:
            if op.post:
                aux = 2 * slot_of[op.aux.id]            # third source cell of a fused instruction
            else:
                aux = flag_of[op.aux.id] if _is_val(op.aux) else int(op.aux)
        # free sources that die here (so the destination can reuse their cells), except for
        # cross-thread reads where other threads still read the source after we write
        dying = []
        for x in list(srcs) + [y for y in (op.a, op.b, op.aux) if _is_val(y) and isinstance(root(y), Flag)]:
            vid = root(x).id
            if vid not in dying and next_use(vid, i) == INF and (vid not in fixed or (epi0 is not None and i >= epi0)):
                dying.append(vid)
        if True:
            for vid in dying:
                if vid in flag_of:
                    pass
                else:
                    release(vid)
        # destination
        if _is_val(op.d):
            r = root(op.d)
            if isinstance(r, Flag):
                if r.id not in flag_of:
                    if not free_flags:
                        raise RuntimeError("out of flags")
                    flag_of[r.id] = free_flags.pop()
                conc[0] = flag_of[r.id]
            else:
                if r.id not in slot_of:
                    if r.id in cold_of:          # partial def of a spilled value: bring it back
                        s = alloc_slot(i, pinned)
                        slot_of[r.id] = s
                        emit("FILL2", 2 * s, cold_of[r.id])
                        stats["fills"] += 1
                    else:
                        slot_of[r.id] = alloc_slot(i, pinned)
                if r.id in cold_of:              # cold copy goes stale on a (partial) write
                    c = cold_of.pop(r.id)
                    free_cold.append(c)
                    discarded.discard(r.id)
                    if c < 0x80:
                        pending.append(c)
                conc[0] = 2 * slot_of[r.id] + (op.d.half if isinstance(op.d, Half) else 0)
        elif op.d is not None:
            conc[0] = int(op.d)
        if op.name == "SKIPZ":
            skip_stack.append(len(out))
        emit(op.name, conc[0] | (op.post << isa.POST_SHIFT), conc[1], conc[2], aux)
        for vid in dying:
            if vid in flag_of:
                free_flags.append(flag_of.pop(vid))
        # a destination that is never used is dead immediately
        if _is_val(op.d):
            r = root(op.d)
            if r.id not in fixed and next_use(r.id, i) == INF:
                if isinstance(r, Flag):
                    if r.id in flag_of:
                        free_flags.append(flag_of.pop(r.id))
                else:
                    release(r.id)
    # every section ends with END: [prologue END | body END | epilogue END]
    if marks[0] is None:
        flush_discards()
        emit("END")
        marks[0] = len(out)
    if marks[1] is None:
        flush_discards()
        emit("END")
        marks[1] = len(out)
    flush_discards()
    emit("END")
    code = np.array(out, dtype=np.uint16).reshape(-1, 4)
    stats["n_ins"] = len(out)
    return Assembled(prog.name, code, list(prog.consts), tuple(marks), n_slots, n_cold, stats)
