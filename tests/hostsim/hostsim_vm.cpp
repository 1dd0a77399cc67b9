// Host build of the VM instruction bodies (csrc/vm_exec.cuh) with emulated PTX.
// TEST INFRASTRUCTURE for the CPU-only development box: runs assembled VM programs with the
// same cell / cold / SoA layouts as the CUDA kernel, thread by thread in lock step.  Never
// part of the product library (which has no CPU path).
#define B200BLS_HOSTSIM 1
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../python-bls_b200/csrc/vm_exec.cuh"
using namespace b200bls;

namespace {
struct Shared {
  const uint32_t* consts;
  uint8_t* bufs[8];
  long strides[8];
  long n_items;
  int nt, n_blocks, total;
  int n_cells;
  int smem_cells;                // cells >= smem_cells model Tensor Memory (thread-private)
  std::vector<uint32_t>* smem;   // per block: [cell][chunk][tid][4]
  // cross-thread hazard detector (all threads of a block step together, so one flag per cell):
  // a cell written since the last barrier must not be read by another thread, and a cell read
  // by another thread must not be written before the next barrier
  std::vector<unsigned char> written, xread;
  std::vector<uint32_t> cold;    // [(g*6+k)*total + gtid][4]
};

struct HostEnv {
  Shared* sh;
  int tid, blk, gtid;
  long item_raw, item;
  uint32_t flags;

  uint32_t* cellp(int cell, int k, int t) {
    return sh->smem[blk].data() + (((size_t)cell * 3 + k) * sh->nt + t) * 4;
  }
  void ld1(int c, fp& x) {
    for (int k = 0; k < 3; k++) memcpy(&x.v[4 * k], cellp(c, k, tid), 16);
  }
  void st1(int c, const fp& x) {
    if (c < sh->smem_cells) {
      if (sh->xread[c]) abort();   // write-after-cross-read without a barrier
      sh->written[c] = 1;
    }
    for (int k = 0; k < 3; k++) memcpy(cellp(c, k, tid), &x.v[4 * k], 16);
  }
  void ld2(int c, fp2& x) { ld1(c, x.c0); ld1(c + 1, x.c1); }
  void st2(int c, const fp2& x) { st1(c, x.c0); st1(c + 1, x.c1); }
  void ld2_lane(int c, int off, fp2& x) {
    if (c + 1 >= sh->smem_cells) abort();  // TMEM lanes cannot be read by another thread
    if (sh->written[c] || sh->written[c + 1]) abort();  // cross-read of a cell written after the last barrier
    sh->xread[c] = sh->xread[c + 1] = 1;
    int t = (tid + off) % sh->nt;
    for (int k = 0; k < 3; k++) {
      memcpy(&x.c0.v[4 * k], cellp(c, k, t), 16);
      memcpy(&x.c1.v[4 * k], cellp(c + 1, k, t), 16);
    }
  }
  void ldc(int idx, fp& x) { memcpy(x.v, sh->consts + (size_t)idx * NL, 4 * NL); }
  void set_flag(int f, bool v) { flags = (flags & ~(1u << f)) | ((v ? 1u : 0u) << f); }
  bool get_flag(int f) { return (flags >> f) & 1; }
  bool any_flag(int f) { return true; }  // hostsim never skips: SKIPZ is an optimisation only
  bool active() { return item_raw < sh->n_items; }
  uint32_t ld_byte(int buf, int off) { return sh->bufs[buf][item * sh->strides[buf] + off]; }
  void st_byte(int buf, int off, uint8_t v, bool block_only) {
    if (block_only) {
      if (tid == 0) sh->bufs[buf][blk * sh->strides[buf] + off] = v;
    } else if (active()) {
      sh->bufs[buf][item * sh->strides[buf] + off] = v;
    }
  }
  void ld_be(int buf, int off, int nwords, fp& x) {
    const uint8_t* p = sh->bufs[buf] + item * sh->strides[buf] + off;
    fp_set_zero(x);
    for (int i = 0; i < nwords; i++) {
      uint32_t w;
      memcpy(&w, p + 4 * (nwords - 1 - i), 4);
      x.v[i] = __builtin_bswap32(w);
    }
  }
  void st_be48(int buf, int off, const fp& x, bool block_only) {
    long it = item;
    if (block_only) {
      if (tid != 0) return;
      it = blk;
    } else if (!active()) {
      return;
    }
    uint8_t* p = sh->bufs[buf] + it * sh->strides[buf] + off;
    for (int i = 0; i < NL; i++) {
      uint32_t w = __builtin_bswap32(x.v[i]);
      memcpy(p + 4 * (NL - 1 - i), &w, 4);
    }
  }
  uint32_t* rawp(int buf, int elem, int k, long it) {
    return (uint32_t*)sh->bufs[buf] + (((size_t)elem * 6 + k) * sh->strides[buf] + it) * 4;
  }
  void ld_raw2(int buf, int elem, fp2& x) {
    for (int k = 0; k < 3; k++) {
      memcpy(&x.c0.v[4 * k], rawp(buf, elem, k, item), 16);
      memcpy(&x.c1.v[4 * k], rawp(buf, elem, 3 + k, item), 16);
    }
  }
  void st_raw2(int buf, int elem, const fp2& x, bool block_only) {
    long it;
    if (block_only) {
      if (tid != 0) return;
      it = blk;
    } else {
      if (!active()) return;
      it = item_raw;
    }
    for (int k = 0; k < 3; k++) {
      memcpy(rawp(buf, elem, k, it), &x.c0.v[4 * k], 16);
      memcpy(rawp(buf, elem, 3 + k, it), &x.c1.v[4 * k], 16);
    }
  }
  uint32_t* coldp(int g, int k) { return sh->cold.data() + (((size_t)g * 6 + k) * sh->total + gtid) * 4; }
  void st_cold(int g, const fp2& x) {
    for (int k = 0; k < 3; k++) {
      memcpy(coldp(g, k), &x.c0.v[4 * k], 16);
      memcpy(coldp(g, 3 + k), &x.c1.v[4 * k], 16);
    }
  }
  void ld_cold(int g, fp2& x) {
    for (int k = 0; k < 3; k++) {
      memcpy(&x.c0.v[4 * k], coldp(g, k), 16);
      memcpy(&x.c1.v[4 * k], coldp(g, 3 + k), 16);
    }
  }
  void discard_cold(int g) {   // the device drops the cache lines: whatever is read from here afterwards is garbage
    for (int k = 0; k < 6; k++)
      for (int j = 0; j < 4; j++) coldp(g, k)[j] = 0xdeadbeefu;
  }
  void sync() {
    std::fill(sh->written.begin(), sh->written.end(), 0);
    std::fill(sh->xread.begin(), sh->xread.end(), 0);
  }
};
}  // namespace

extern "C" int hs_vm_run(const uint32_t* code, int n_ins, int body_start, int epi_start,
                         const uint32_t* consts, int n_slots, int n_tmem, int n_cold, uint8_t** bufs,
                         const long* strides, long n_items, int n_blocks, int nt,
                         int honor_skips) {
  Shared sh;
  sh.consts = consts;
  for (int i = 0; i < 8; i++) { sh.bufs[i] = bufs[i]; sh.strides[i] = strides[i]; }
  sh.n_items = n_items;
  sh.nt = nt;
  sh.n_blocks = n_blocks;
  sh.total = nt * n_blocks;
  sh.n_cells = 2 * (n_slots + n_tmem);
  sh.smem_cells = 2 * n_slots;
  sh.written.assign(sh.n_cells, 0);
  sh.xread.assign(sh.n_cells, 0);
  std::vector<std::vector<uint32_t>> smem(n_blocks);
  for (auto& s : smem) s.assign((size_t)sh.n_cells * 3 * nt * 4, 0xdeadbeefu);
  sh.smem = smem.data();
  sh.cold.assign((size_t)n_cold * 6 * sh.total * 4, 0xdeadbeefu);
  std::vector<HostEnv> env(sh.total);
  for (int g = 0; g < sh.total; g++) {
    env[g].sh = &sh;
    env[g].gtid = g;
    env[g].blk = g / nt;
    env[g].tid = g % nt;
    env[g].flags = 0;
  }
  long iters = (n_items + sh.total - 1) / sh.total;
  auto run = [&](int lo, int hi, long it) {
    for (int g = 0; g < sh.total; g++) {
      env[g].item_raw = it * sh.total + g;
      long last = n_items > 0 ? n_items - 1 : 0;
      env[g].item = env[g].item_raw < last ? env[g].item_raw : last;
    }
    // blocks are independent; inside a block all threads step together and SKIPZ is decided
    // block-wide (the block plays the role of the warp)
    for (int blk = 0; blk < n_blocks; blk++) {
      for (int pc = lo; pc < hi; pc++) {
        uint32_t w0 = code[2 * pc], w1 = code[2 * pc + 1];
        if ((w0 & 0xff) == OP_SKIPZ) {
          bool any = false;
          for (int t = 0; t < nt; t++) any = any || env[blk * nt + t].get_flag(w0 >> 16);
          if (!any && honor_skips) pc += (int)(w1 & 0xffff);
          continue;
        }
        for (int t = 0; t < nt; t++) vm_exec(env[blk * nt + t], w0, w1);
      }
    }
  };
  run(0, body_start, 0);
  for (long it = 0; it < iters; it++) run(body_start, epi_start, it);
  run(epi_start, n_ins, 0);  // the epilogue addresses the thread's own record: item_raw = global thread id
  return 0;
}

#include "../../python-bls_b200/csrc/sha256.cuh"
// SHA stage of hash-to-G2 on the host (same function the CUDA kernel runs per thread)
extern "C" void hs_sha_stage(const uint8_t* hashes, uint8_t* out, long n) {
  for (long item = 0; item < n; item++)
    for (int sub = 0; sub < 8; sub++) {
      int j = sub >> 2, k = (sub >> 1) & 1, suffix = sub & 1;
      sha256_h_label(hashes + item * 32, j, k, suffix, out + item * 256 + (j * 2 + k) * 64 + suffix * 32);
    }
}

// aggregation exponents T_i (same function hash_pks_kernel runs per thread)
extern "C" void hs_hash_pks(const uint8_t* pk_hash, uint32_t first, uint8_t* out, long n) {
  uint32_t h[8];
  for (int i = 0; i < 8; i++)
    h[i] = ((uint32_t)pk_hash[4 * i] << 24) | ((uint32_t)pk_hash[4 * i + 1] << 16) | ((uint32_t)pk_hash[4 * i + 2] << 8) |
           pk_hash[4 * i + 3];
  for (long t = 0; t < n; t++) b200bls::hash_pks_exponent(first + (uint32_t)t, h, out + t * 32);
}
