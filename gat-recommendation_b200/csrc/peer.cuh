// Peer-memory communicator shared by peer.cu (barrier, small all-reduce, sums) and optim.cu (the fused
// reduce-scatter + AdamW + all-gather of the item table).  One process per GPU of one NVSwitch box; every rank
// owns one cudaMalloc'ed *region* that all peers map (CUDA IPC, or plain pointers when several ranks live in one
// process), and the kernels exchange data with ordinary loads / stores over NVLink plus system-scope flags.
#pragma once

#include "common.cuh"

namespace etpgt {

constexpr int kMaxRanks = ETPGT_MAX_RANKS;
constexpr int kBarChannels = 4;    // independent barrier sequences (one per stream that synchronises)
constexpr int kArSlots = 8;        // ring of exchange slots of the small all-reduce
constexpr int kArMaxCount = 520;   // doubles per contribution (BatchNorm: 2*256 + 1)

// Start of every region.  Flags are written by PEERS (system-scope release stores) and polled locally;
// the counters are only touched by this rank's own kernels, so that a captured CUDA graph replays correctly.
struct CommControl {
  unsigned long long bar_flag[kBarChannels][kMaxRanks];   // [c][r] = last barrier epoch of channel c rank r reached
  unsigned long long ar_flag[kArSlots][kMaxRanks];    // ar_flag[s][r] = sequence number of r's data in slot s
  unsigned long long bar_epoch[kBarChannels];         // barriers this rank has entered, per channel
  unsigned long long ar_seq;                          // small all-reduces this rank has entered
  unsigned int status;                                // != 0: a wait timed out (see etpgt_comm_status)
  unsigned int pad_[61];
  double ar_slot[kArSlots][kMaxRanks][kArMaxCount];   // ar_slot[s][r] = contribution of rank r
};

// What a kernel needs of the communicator (passed by value).
struct CommView {
  char* base[kMaxRanks];     // this process's mapping of every rank's region
  int rank, world;
  unsigned long long timeout_ns;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Spins until *flag >= want.  Bounded: after timeout_ns the rank records the failure in its status word and
// carries on (results are then wrong, but the GPU is never left hanging on a peer that died).
__device__ __forceinline__ void wait_flag(const unsigned long long* flag, unsigned long long want,
                                          unsigned long long timeout_ns, unsigned int* status, unsigned int code) {
  if (ld_acquire_sys(flag) >= want) return;
  if (*reinterpret_cast<volatile unsigned int*>(status) != 0) return;   // a peer is already known dead: one time-out only
  const unsigned long long t0 = global_timer_ns();
  while (ld_acquire_sys(flag) < want) {
    __nanosleep(64);
    if (global_timer_ns() - t0 > timeout_ns) {
      atomicExch(status, code);
      return;
    }
  }
}

CommView comm_view(const etpgt_comm_t* comm);   // peer.cu

}  // namespace etpgt
