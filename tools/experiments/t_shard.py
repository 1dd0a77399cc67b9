import os, sys
import numpy as np
sys.path.insert(0, "python-bls_b200"); sys.path.insert(0, "oracle")
import bls_oracle as O
from bls_b200 import engine, synth, distributed as D
from bls_b200.programs.hashg2 import G2_GEN
g2 = np.frombuffer(b"".join(c.to_bytes(48, "big") for c in (G2_GEN[0] + G2_GEN[1])), dtype=np.uint8)
n3 = 1_000_000
sc_all = synth.scalars(synth.SEED_AGGREGATE, n3)
def ser2(p): return b"".join(c.to_bytes(48,"big") for c in (p[0][0],p[0][1],p[1][0],p[1][1]))
for world in (2, 8):
    parts = []
    for rank in range(world):
        lo, hi = D.shard_range(n3, rank, world)
        sc = sc_all[lo:hi]
        pts = engine.scalar_mul(np.tile(g2, hi - lo), sc, True)
        part = engine.point_sum(pts, True).tobytes()
        tot = sum(int.from_bytes(bytes(r), "big") for r in sc) % O.N
        ok = part == ser2(O.aff_mul(tot, O.G2))
        parts.append(part)
        print(world, rank, hi - lo, "partial ok:", ok, flush=True)
    comb = engine.point_sum(b"".join(parts), True).tobytes()
    tot = sum(int.from_bytes(bytes(r), "big") for r in sc_all) % O.N
    print(world, "combined ok:", comb == ser2(O.aff_mul(tot, O.G2)), flush=True)
