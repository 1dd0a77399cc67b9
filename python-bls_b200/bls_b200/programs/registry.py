"""Name -> builder table of every VM program embedded into libb200bls.so."""
from . import fieldops, pairing

N_SLOTS = 18        # Fq2 slots per thread: 18 * 96 B * 128 threads = 216 KB of shared memory

PROGRAMS = {}
for _level in (1, 2, 6, 12):
    for _op in ("add", "sub", "mul", "sqr", "neg", "inv"):
        PROGRAMS["f%d_%s" % (_level, _op)] = fieldops.build_field_op(_level, _op)
PROGRAMS["fq2_mul_chain"] = fieldops.build_fq2_mul_chain(512)
PROGRAMS["pairing"] = pairing.build_pairing
PROGRAMS["miller_loop"] = pairing.build_miller_only
PROGRAMS["final_exp"] = pairing.build_final_exp
