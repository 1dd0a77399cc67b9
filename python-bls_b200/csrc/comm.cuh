// comm.cuh -- the one exchange step of the sharded reductions, native and torch-free.
//
// One process per GPU (the launcher's RANK / WORLD_SIZE); every rank reduces its contiguous slice to ONE tiny
// partial -- a 576-byte Miller product or one affine point -- and the partials are all-gathered so that every
// rank finishes locally.  Two transports, both measured by bench.py:
//   * host gather: a POSIX shared-memory segment of the node (payloads <= 4 KB per rank, sequence-numbered
//     slots in two banks, acquire / release flags) -- no library at all;
//   * ncclAllGather over NVLink / NVSwitch on the library stream, NCCL loaded at run time with dlopen (the
//     unique id travels through the same shared-memory segment).
// all_reduce cannot express either reduction: the group operations are not built-in reductions.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <atomic>

namespace b200bls {

constexpr int COMM_MAX_RANKS = 64;
constexpr size_t COMM_SLOT_BYTES = 4096;

struct NcclUniqueId {
  char internal[128];
};

struct CommShm {
  std::atomic<uint32_t> attached;                       // ranks that have mapped the segment
  std::atomic<uint32_t> id_ready;                       // rank 0 has published the NCCL unique id
  NcclUniqueId nccl_id;
  std::atomic<uint32_t> flag[2][COMM_MAX_RANKS];        // flag[bank][rank] = sequence number + 1 of the payload there
  alignas(64) unsigned char slot[2][COMM_MAX_RANKS][COMM_SLOT_BYTES];
};

struct Comm {
  int rank = 0, world = 1;
  CommShm* shm = nullptr;
  char shm_name[96] = "";
  uint32_t seq = 0;
  // NCCL through dlopen
  void* nccl_lib = nullptr;
  void* nccl_comm = nullptr;
  int (*p_get_unique_id)(NcclUniqueId*) = nullptr;
  int (*p_comm_init_rank)(void**, int, NcclUniqueId, int) = nullptr;
  int (*p_all_gather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*p_comm_destroy)(void*) = nullptr;
  const char* (*p_error_string)(int) = nullptr;
  void* gather_dev = nullptr;   // world x COMM_SLOT_BYTES device buffer for the NCCL path
  void* send_dev = nullptr;
};

inline bool comm_wait(std::atomic<uint32_t>& a, uint32_t want, double timeout_s) {
  timespec t0;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (unsigned spins = 0;; spins++) {
    if (a.load(std::memory_order_acquire) >= want) return true;
    if ((spins & 1023) == 1023) {
      timespec t1;
      clock_gettime(CLOCK_MONOTONIC, &t1);
      if ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec) > timeout_s) return false;
      usleep(50);
    }
  }
}

// all-gather of `bytes` (<= COMM_SLOT_BYTES) host bytes per rank through the shared segment -> recv[world][bytes]
inline int comm_allgather_host(Comm& c, const void* send, void* recv, size_t bytes) {
  if (c.world == 1) {
    memcpy(recv, send, bytes);
    return 0;
  }
  const uint32_t s = c.seq++;
  const int bank = s & 1;
  // bank reuse is safe: a rank writes sequence s + 2 only after it has seen every rank's s + 1, and a rank
  // publishes s + 1 only after it has read all of s
  memcpy(c.shm->slot[bank][c.rank], send, bytes);
  c.shm->flag[bank][c.rank].store(s + 1, std::memory_order_release);
  for (int r = 0; r < c.world; r++) {
    if (!comm_wait(c.shm->flag[bank][r], s + 1, 120.0)) return -1;
    memcpy((unsigned char*)recv + (size_t)r * bytes, c.shm->slot[bank][r], bytes);
  }
  return 0;
}

}  // namespace b200bls
