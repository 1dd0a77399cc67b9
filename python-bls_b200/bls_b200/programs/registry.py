"""Name -> builder table of every VM program embedded into libb200bls.so."""
from . import fieldops, pairing

# Fq2 slots per thread.  One CTA of 128 threads per SM gets 18 slots (216 KB of shared memory);
# two co-resident CTAs per SM get 9 slots each and hide each other's latencies.  Every program
# is assembled for both shapes; the variant for n != 18 is named "<name>#<n>".
N_SLOTS = 18
SLOT_VARIANTS = (18, 9)

PROGRAMS = {}
for _level in (1, 2, 6, 12):
    for _op in ("add", "sub", "mul", "sqr", "neg", "inv"):
        PROGRAMS["f%d_%s" % (_level, _op)] = fieldops.build_field_op(_level, _op)
PROGRAMS["fq2_mul_chain"] = fieldops.build_fq2_mul_chain(512)
PROGRAMS["pairing"] = pairing.build_pairing
PROGRAMS["miller_loop"] = pairing.build_miller_only
PROGRAMS["final_exp"] = pairing.build_final_exp
