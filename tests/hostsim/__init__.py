"""Builds and drives the host simulation of the VM (test infrastructure, see hostsim_vm.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    """both executors in one library: hs_vm2_run = the product kernel's paired executor (vm_exec2.cuh, two
    threads per item), hs_vm_run = the one-thread-per-item executor (vm_exec.cuh), the plain statement of
    what every instruction means"""
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "_hostsim_vm.so")
        srcs = [os.path.join(HERE, "hostsim_vm.cpp"), os.path.join(HERE, "hostsim_vm2.cpp")]
        deps = srcs + [os.path.join(HERE, "..", "..", "python-bls_b200", "csrc", f)
                       for f in ("fp.cuh", "vm_exec.cuh", "vm_exec2.cuh", "gen/vm_isa.h", "gen/fp_consts.h")]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so] + srcs)
        _LIB = ctypes.CDLL(so)
    return _LIB


def run(asm, bufs, strides, n_items, n_blocks=1, nt=4, honor_skips=True, paired=None):
    """bufs: {id: np.uint8 array (modified in place)}; strides: {id: bytes per item, or item
    capacity for raw SoA buffers}; nt = items per block; paired selects the executor (see lib()).
    paired=None (default) runs BOTH executors -- the library ships both kernels -- and requires byte-identical
    buffers from them."""
    if paired is None:
        twin = {i: b.copy() for i, b in bufs.items()}
        run(asm, twin, strides, n_items, n_blocks, nt, honor_skips, paired=True)
        run(asm, bufs, strides, n_items, n_blocks, nt, honor_skips, paired=False)
        for i in bufs:
            assert np.array_equal(bufs[i], twin[i]), "paired and one-thread executors disagree on buffer %d" % i
        return bufs
    code = np.ascontiguousarray(asm.code).view(np.uint32).reshape(-1)
    consts = np.ascontiguousarray(asm.const_limbs())
    ptrs = (ctypes.c_void_p * 8)()
    st = (ctypes.c_long * 8)()
    for i in range(8):
        if i in bufs:
            assert bufs[i].dtype == np.uint8 and bufs[i].flags["C_CONTIGUOUS"]
            ptrs[i] = bufs[i].ctypes.data
            st[i] = strides.get(i, 0)
    fn = lib().hs_vm2_run if paired else lib().hs_vm_run
    rc = fn(code.ctypes.data_as(ctypes.c_void_p), len(asm.code), asm.body_start,
                         asm.epilogue_start, consts.ctypes.data_as(ctypes.c_void_p),
                         asm.n_slots, asm.n_tmem, max(asm.stats["max_cold"], 1), ptrs, st,
                         ctypes.c_long(n_items), n_blocks, nt, int(honor_skips))
    assert rc == 0
    return bufs
