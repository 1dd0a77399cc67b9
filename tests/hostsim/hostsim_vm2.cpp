// Host build of the PAIRED VM instruction bodies (csrc/vm_exec2.cuh) with emulated PTX.
// TEST INFRASTRUCTURE for the CPU-only development box: runs assembled VM programs the way the
// product kernel (csrc/vm_kernel2.cuh) does -- two threads per item, thread `role` owning coefficient
// `role` of every Fq2 slot, the same slot / cold / raw SoA layouts -- item by item in lock step.
// Never part of the product library (which has no CPU path).
//
// The two threads of a pair are simulated in two passes per instruction: pass 0 runs both roles with
// stores suppressed and records what each would hand to its partner (Env::xchg, the device's
// shuffle); pass 1 runs both again with the partner's recorded value and commits the stores after
// both have finished -- so every load of an instruction sees the state before it, as on the device
// where both threads of a warp load before either stores.
#define B200BLS_HOSTSIM 1
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../python-bls_b200/csrc/vm_exec2.cuh"
using namespace b200bls;

namespace {
struct Store {
  uint32_t* dst;
  uint32_t v[4];
};
struct ByteStore {
  uint8_t* dst;
  uint8_t v;
};

struct Shared {
  const uint32_t* consts;
  uint8_t* bufs[8];
  long strides[8];
  long n_items;
  int items_per_block, n_blocks, total_items;
  int n_slots_all;               // shared + tensor-memory slots
  int smem_slots;
  std::vector<uint32_t>* ws;     // per block: [slot][chunk 0..2][thread = 2 * item + role][4]
  std::vector<unsigned char> written, xread;   // per slot: cross-ITEM hazard detector (barriers)
  std::vector<uint32_t> cold;    // [(g * 3 + k) * total_threads + gthread][4]
  // two-pass machinery
  int pass;
  std::vector<Store> stores;
  std::vector<ByteStore> byte_stores;
  fp xchg_val[2];
  int xchg_count[2];
};

struct PairHostEnv {
  Shared* sh;
  int tid, blk;                  // thread within the block (2 * item + role), block
  long item_raw, item;
  uint32_t flags;

  int role() const { return tid & 1; }
  int nt() const { return 2 * sh->items_per_block; }
  long gthread() const { return (long)blk * nt() + tid; }
  uint32_t* wsp(int slot, int k, int t) { return sh->ws[blk].data() + (((size_t)slot * 3 + k) * nt() + t) * 4; }
  void put(uint32_t* dst, const uint32_t* v) {
    if (sh->pass == 0) return;
    Store s;
    s.dst = dst;
    memcpy(s.v, v, 16);
    sh->stores.push_back(s);
  }
  void ld_thread(int slot, int t, fp& x) {
    for (int k = 0; k < 3; k++) memcpy(&x.v[4 * k], wsp(slot, k, t), 16);
  }
  void ld_own(int slot, fp& x) { ld_thread(slot, tid, x); }
  void ld_oth(int slot, fp& x) { ld_thread(slot, tid ^ 1, x); }
  void ld_cell(int c, fp& x) { ld_thread(c >> 1, (tid & ~1) | (c & 1), x); }
  void st_thread(int slot, int t, const fp& x) {
    if (slot < sh->smem_slots) {
      if (sh->xread[slot]) abort();   // write-after-cross-read without a barrier
      if (sh->pass == 1) sh->written[slot] = 1;
    }
    for (int k = 0; k < 3; k++) put(wsp(slot, k, t), &x.v[4 * k]);
  }
  void st_own(int slot, const fp& x) { st_thread(slot, tid, x); }
  void st_cell(int c, const fp& x) {
    if ((c & 1) == role()) st_thread(c >> 1, tid, x);
  }
  void xchg(fp& x) {
    const int r = role();
    if (sh->pass == 0) {
      if (sh->xchg_count[r]++ != 0) abort();   // one exchange per instruction is what the two passes support
      sh->xchg_val[r] = x;
    } else {
      x = sh->xchg_val[r ^ 1];
    }
  }
  void mul2(fp& r, const fp& a, const fp& b, int slot, bool swap) {
    fp own, oth;
    ld_own(slot, own);
    ld_oth(slot, oth);
    if (swap)
      fp_mul2_looped(r, a, oth, b, own);
    else
      fp_mul2_looped(r, a, own, b, oth);
  }
  void ld_lane_own(int slot, int off, fp& x) {
    if (slot >= sh->smem_slots) abort();      // TMEM lanes cannot be read by another thread
    if (sh->written[slot]) abort();           // cross-read of a slot written after the last barrier
    if (sh->pass == 1) sh->xread[slot] = 1;
    ld_thread(slot, (tid + 2 * off) % nt(), x);
  }
  void ldc(int idx, fp& x) { memcpy(x.v, sh->consts + (size_t)idx * NL, 4 * NL); }
  void set_flag(int f, bool v) { flags = (flags & ~(1u << f)) | ((v ? 1u : 0u) << f); }
  bool get_flag(int f) { return (flags >> f) & 1; }
  bool any_flag(int) { return true; }  // SKIPZ is handled by the driver loop below
  bool active() { return item_raw < sh->n_items; }
  uint32_t ld_byte(int buf, int off) { return sh->bufs[buf][item * sh->strides[buf] + off]; }
  void put_byte(uint8_t* dst, uint8_t v) {
    if (sh->pass == 0) return;
    ByteStore s = {dst, v};
    sh->byte_stores.push_back(s);
  }
  void st_byte(int buf, int off, uint8_t v, bool block_only) {
    if (role() != 0) return;
    if (block_only) {
      if (tid == 0) put_byte(&sh->bufs[buf][blk * sh->strides[buf] + off], v);
    } else if (active()) {
      put_byte(&sh->bufs[buf][item * sh->strides[buf] + off], v);
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
  void st_be48(int buf, int off, const fp& x, bool block_only, int parity) {
    if (role() != parity) return;
    long it = item;
    if (block_only) {
      if (tid >= 2) return;
      it = blk;
    } else if (!active()) {
      return;
    }
    uint8_t* p = sh->bufs[buf] + it * sh->strides[buf] + off;
    for (int i = 0; i < NL; i++) {
      uint32_t w = __builtin_bswap32(x.v[i]);
      uint8_t b[4];
      memcpy(b, &w, 4);
      for (int j = 0; j < 4; j++) put_byte(p + 4 * (NL - 1 - i) + j, b[j]);
    }
  }
  uint32_t* rawp(int buf, int elem, int k, long it) {
    return (uint32_t*)sh->bufs[buf] + (((size_t)elem * 6 + k) * sh->strides[buf] + it) * 4;
  }
  void ld_raw_own(int buf, int elem, fp& x) {
    for (int k = 0; k < 3; k++) memcpy(&x.v[4 * k], rawp(buf, elem, 3 * role() + k, item), 16);
  }
  void st_raw_own(int buf, int elem, const fp& x, bool block_only) {
    long it;
    if (block_only) {
      if (tid >= 2) return;
      it = blk;
    } else {
      if (!active()) return;
      it = item_raw;
    }
    for (int k = 0; k < 3; k++) put(rawp(buf, elem, 3 * role() + k, it), &x.v[4 * k]);
  }
  uint32_t* coldp(int g, int k) {
    const size_t total = (size_t)sh->n_blocks * nt();
    return sh->cold.data() + (((size_t)g * 3 + k) * total + gthread()) * 4;
  }
  void st_cold_own(int g, const fp& x) {
    for (int k = 0; k < 3; k++) put(coldp(g, k), &x.v[4 * k]);
  }
  void ld_cold_own(int g, fp& x) {
    for (int k = 0; k < 3; k++) memcpy(&x.v[4 * k], coldp(g, k), 16);
  }
  void discard_cold_own(int g) {   // the device drops the cache lines: a later read would see garbage
    static const uint32_t poison[4] = {0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu};
    for (int k = 0; k < 3; k++) put(coldp(g, k), poison);
  }
  void sync() {
    if (sh->pass == 0) return;
    std::fill(sh->written.begin(), sh->written.end(), 0);
    std::fill(sh->xread.begin(), sh->xread.end(), 0);
  }
};
}  // namespace

extern "C" int hs_vm2_run(const uint32_t* code, int n_ins, int body_start, int epi_start, const uint32_t* consts,
                          int n_slots, int n_tmem, int n_cold, uint8_t** bufs, const long* strides, long n_items,
                          int n_blocks, int items_per_block, int honor_skips) {
  (void)n_ins;
  Shared sh;
  sh.consts = consts;
  for (int i = 0; i < 8; i++) {
    sh.bufs[i] = bufs[i];
    sh.strides[i] = strides[i];
  }
  sh.n_items = n_items;
  sh.items_per_block = items_per_block;
  sh.n_blocks = n_blocks;
  sh.total_items = items_per_block * n_blocks;
  sh.n_slots_all = n_slots + n_tmem;
  sh.smem_slots = n_slots;
  sh.written.assign(sh.n_slots_all, 0);
  sh.xread.assign(sh.n_slots_all, 0);
  const int nt = 2 * items_per_block;
  std::vector<std::vector<uint32_t>> ws(n_blocks);
  for (auto& s : ws) s.assign((size_t)sh.n_slots_all * 3 * nt * 4, 0xdeadbeefu);
  sh.ws = ws.data();
  sh.cold.assign((size_t)n_cold * 3 * nt * n_blocks * 4, 0xdeadbeefu);
  std::vector<PairHostEnv> env((size_t)nt * n_blocks);
  for (int b = 0; b < n_blocks; b++)
    for (int t = 0; t < nt; t++) {
      PairHostEnv& e = env[(size_t)b * nt + t];
      e.sh = &sh;
      e.blk = b;
      e.tid = t;
      e.flags = 0;
    }
  const long iters = (n_items + sh.total_items - 1) / sh.total_items;
  auto run = [&](int lo, long it) {
    for (int b = 0; b < n_blocks; b++)
      for (int t = 0; t < nt; t++) {
        PairHostEnv& e = env[(size_t)b * nt + t];
        e.item_raw = it * sh.total_items + (long)b * items_per_block + (t >> 1);
        const long last = n_items > 0 ? n_items - 1 : 0;
        e.item = e.item_raw < last ? e.item_raw : last;
      }
    // blocks are independent; inside a block all pairs step together, SKIPZ is decided block-wide
    // (the block plays the role of the warp) and the section runs to its END instruction
    for (int b = 0; b < n_blocks; b++) {
      for (int pc = lo;; pc++) {
        const uint32_t w0 = code[2 * pc], w1 = code[2 * pc + 1];
        const int op = w0 & 0xff;
        if (op == OP_END) break;
        if (op == OP_SKIPZ) {
          bool any = false;
          for (int t = 0; t < nt; t++) any = any || env[(size_t)b * nt + t].get_flag(w0 >> 16);
          if (!any && honor_skips) pc += (int)(w1 & 0xffff);
          continue;
        }
        sh.stores.clear();
        sh.byte_stores.clear();
        for (int item = 0; item < items_per_block; item++) {
          PairHostEnv* pr = &env[(size_t)b * nt + 2 * item];
          const uint32_t f0 = pr[0].flags, f1 = pr[1].flags;
          sh.xchg_count[0] = sh.xchg_count[1] = 0;
          sh.pass = 0;
          vm_exec2(pr[0], w0, w1);
          vm_exec2(pr[1], w0, w1);
          pr[0].flags = f0;
          pr[1].flags = f1;
          sh.pass = 1;
          vm_exec2(pr[0], w0, w1);
          vm_exec2(pr[1], w0, w1);
          if (pr[0].flags != pr[1].flags) abort();   // flags are replicated in the two threads of a pair
        }
        // all loads of the instruction precede all of its stores (lock step inside the block)
        for (const Store& s : sh.stores) memcpy(s.dst, s.v, 16);
        for (const ByteStore& s : sh.byte_stores) *s.dst = s.v;
      }
    }
  };
  run(0, 0);
  for (long it = 0; it < iters; it++) run(body_start, it);
  run(epi_start, iters > 0 ? iters - 1 : 0);
  return 0;
}
