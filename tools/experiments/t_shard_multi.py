import os, sys
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "python-bls_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bls_oracle as O
from bls_b200 import _lib, engine, synth, distributed as D
from bls_b200.programs.hashg2 import G2_GEN
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gloo = dist.new_group(backend="gloo")
_lib.init(local)
g2 = np.frombuffer(b"".join(c.to_bytes(48, "big") for c in (G2_GEN[0] + G2_GEN[1])), dtype=np.uint8)
def ser2(p): return b"".join(c.to_bytes(48,"big") for c in (p[0][0],p[0][1],p[1][0],p[1][1]))
n3 = 1_000_000
sc_all = synth.scalars(synth.SEED_AGGREGATE, n3)
lo, hi = D.shard_range(n3, rank, world)
sc = sc_all[lo:hi]
pts = engine.scalar_mul(np.tile(g2, hi - lo), sc, True)
tot = sum(int.from_bytes(bytes(r), "big") for r in sc) % O.N
want_part = ser2(O.aff_mul(tot, O.G2))
for rep in range(3):
    part = engine.point_sum(pts, True).tobytes()
    print("rank", rank, "rep", rep, "partial ok", part == want_part, flush=True)
class G:
    def __init__(s, g, b): s.group, s.backend = g, b
    def is_initialized(s): return True
    def get_world_size(s): return world
    def get_backend(s): return s.backend
    def all_gather(s, out, t): return dist.all_gather(out, t, group=s.group)
pg = D.gather_bytes(part, G(gloo, "gloo"))
pn = D.gather_bytes(part, G(None, "nccl"))
print("rank", rank, "gathers equal", pg == pn, "own slot ok", pg[rank] == part, flush=True)
comb = engine.point_sum(b"".join(pg), True).tobytes()
tot_all = sum(int.from_bytes(bytes(r), "big") for r in sc_all) % O.N
print("rank", rank, "combined ok", comb == ser2(O.aff_mul(tot_all, O.G2)), flush=True)
dist.destroy_process_group()
