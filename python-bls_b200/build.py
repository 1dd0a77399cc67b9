#!/usr/bin/env python3
"""Builds libb200bls.so for sm_100a (nvcc cross-compiles without a GPU).

  1. tools/gen_headers.py   -> csrc/gen/vm_isa.h, fp_consts.h
  2. tools/gen_programs.py  -> csrc/gen/programs.bin  (assembled VM programs)
  3. ld -r -b binary        -> csrc/gen/programs_blob.o (embedded into the library)
  4. nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> bls_b200/libb200bls.so
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
GEN = os.path.join(CSRC, "gen")
LIB = os.path.join(HERE, "bls_b200", "libb200bls.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-DB200BLS_MUL_CALL"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _walk(d, exts):
    out = []
    for base, _, files in os.walk(d):
        out += [os.path.join(base, f) for f in files if f.endswith(exts)]
    return out


def build(force=False, verbose=True):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    py_deps = _walk(os.path.join(HERE, "bls_b200", "programs"), (".py",)) + \
        _walk(os.path.join(HERE, "bls_b200", "vm"), (".py",)) + _walk(os.path.join(ROOT, "tools"), (".py",))
    blob = os.path.join(GEN, "programs.bin")
    if force or _newer(blob, py_deps) or _newer(os.path.join(GEN, "vm_isa.h"), py_deps):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_headers.py")])
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_programs.py")])
    blob_o = os.path.join(GEN, "programs_blob.o")
    if force or _newer(blob_o, [blob]):
        subprocess.check_call(["ld", "-r", "-b", "binary", "-z", "noexecstack", "-o", "programs_blob.o",
                               "programs.bin"], cwd=GEN)
    srcs = [os.path.join(CSRC, f) for f in ("b200bls.cu", "kernel1.cu", "kernel2.cu", "kernel3.cu", "microbench.cu")]
    deps = srcs + _walk(CSRC, (".cuh", ".h")) + [blob_o, os.path.join(ROOT, "include", "b200bls.h")]
    if force or _newer(LIB, deps):
        cmd = ["nvcc"] + NVCC_FLAGS + ["-shared", "-o", LIB] + srcs + [blob_o, "-ldl", "-lrt"]
        if verbose:
            print(" ".join(cmd))
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        log = os.path.join(GEN, "nvcc_build.log")
        with open(log, "w") as fh:
            fh.write(res.stdout)
        if verbose or res.returncode:
            for line in res.stdout.splitlines():
                if any(k in line for k in ("error", "warning", "registers", "spill", "vm_kernel")):
                    print(line)
        if res.returncode:
            raise RuntimeError("nvcc failed, see %s" % log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
