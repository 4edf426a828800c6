// (e) multi-GPU: the peer-memory communicator of session-batch data parallelism.
//
// The reference is single-GPU (etpgt/train/trainer.py:69-131).  Sharded over one 8 x B200 NVSwitch box the
// step has exactly three cross-rank couplings: BatchNorm statistics (2*dim+1 doubles, four times per step for
// two layers), the dense parameter gradients (2.1 MB) and the item-table gradient / update (84 MB at 82k
// items, 1 GB at 1M).  None of them needs a library collective: every rank maps every peer's *region* and
//   * the BatchNorm exchange is ONE single-CTA kernel (push my 4 KB to every peer, flag, wait, sum in rank
//     order) between the statistics kernel and the apply kernel — no phase cut, no host call, ~10 us;
//   * the dense gradients are summed by one kernel that reads every peer's flat gradient buffer;
//   * the table is reduce-scattered, updated and all-gathered by ONE kernel (optim.cu: etpgt_dp_adam_table).
// Every sum runs in rank order 0..world-1 on every rank, so replicas stay bit-identical.
//
// Ordering protocol: flags live in the control block at the start of a region, are written by peers with
// st.release.sys after __threadfence_system(), and are polled locally with ld.acquire.sys.  Sequence numbers
// come from device-side counters (one kernel increments its own rank's counter), so the kernels can sit in a
// captured CUDA graph.  All ranks must issue the same sequence of communicator calls, on one stream each.
#include "peer.cuh"

struct etpgt_comm {
  int rank = 0, world = 1, device = 0;
  size_t region_bytes = 0;
  void* base[etpgt::kMaxRanks] = {};
  bool opened[etpgt::kMaxRanks] = {};
  unsigned long long timeout_ns = 30ull * 1000 * 1000 * 1000;
};

namespace etpgt {

CommView comm_view(const etpgt_comm_t* comm) {
  CommView v{};
  for (int r = 0; r < kMaxRanks; ++r) v.base[r] = static_cast<char*>(comm->base[r]);
  v.rank = comm->rank;
  v.world = comm->world;
  v.timeout_ns = comm->timeout_ns;
  return v;
}

namespace {

__device__ __forceinline__ CommControl* control(const CommView& c, int r) {
  return reinterpret_cast<CommControl*>(c.base[r]);
}

// One CTA, one thread per rank: "everything this rank issued before me is done" -> every peer; then wait for
// the same statement from every peer.  Kernel boundaries order it against the neighbouring kernels.
__global__ void comm_barrier_kernel(const __grid_constant__ CommView c, int channel) {
  CommControl* me = control(c, c.rank);
  __shared__ unsigned long long epoch_s;
  if (threadIdx.x == 0) epoch_s = ++me->bar_epoch[channel];
  __syncthreads();
  const unsigned long long epoch = epoch_s;
  const int p = threadIdx.x;
  if (p < c.world) {
    __threadfence_system();
    st_release_sys(&control(c, p)->bar_flag[channel][c.rank], epoch);
    wait_flag(&me->bar_flag[channel][p], epoch, c.timeout_ns, &me->status, 1u);
  }
}

// out[i] = sum over ranks r = 0..world-1 (in that order) of rank r's in[i].  One CTA.
__global__ void __launch_bounds__(512)
comm_allreduce_f64_kernel(const __grid_constant__ CommView c, const double* __restrict__ in, double* __restrict__ out, int count) {
  CommControl* me = control(c, c.rank);
  __shared__ unsigned long long seq_s;
  if (threadIdx.x == 0) seq_s = ++me->ar_seq;
  __syncthreads();
  const unsigned long long seq = seq_s;
  const int slot = (int)(seq % kArSlots);
  // A slot is reused every kArSlots calls.  A peer can be at most one call ahead of the slowest rank (it needs
  // everybody's contribution to finish a call), so the previous tenant of the slot was consumed long ago.
  for (int p = 0; p < c.world; ++p) {
    double* dst = control(c, p)->ar_slot[slot][c.rank];
    for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = in[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < c.world) {
    const int p = threadIdx.x;
    __threadfence_system();
    st_release_sys(&control(c, p)->ar_flag[slot][c.rank], seq);
    wait_flag(&me->ar_flag[slot][p], seq, c.timeout_ns, &me->status, 2u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < c.world; ++r) s += __ldcg(&me->ar_slot[slot][r][i]);
    out[i] = s;
  }
}

// out[i] = sum_r region_r[offset][i], rank order; float4 body, scalar tail.  Reads go over NVLink.
__global__ void __launch_bounds__(256)
comm_sum_f32_kernel(const __grid_constant__ CommView c, size_t offset, int64_t numel, float* __restrict__ out) {
  const int64_t n4 = numel / 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v[kMaxRanks];
#pragma unroll
    for (int r = 0; r < kMaxRanks; ++r)
      if (r < c.world) v[r] = __ldcg(reinterpret_cast<const float4*>(c.base[r] + offset) + i);
    float4 s = v[0];
#pragma unroll
    for (int r = 1; r < kMaxRanks; ++r)
      if (r < c.world) s = add4(s, v[r]);
    st4(out + 4 * i, s);
  }
  for (int64_t i = 4 * n4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < numel; i += stride) {
    float s = 0.f;
    for (int r = 0; r < c.world; ++r) {
      const float v = __ldcg(reinterpret_cast<const float*>(c.base[r] + offset) + i);
      s = r == 0 ? v : s + v;
    }
    out[i] = s;
  }
}

bool cuda_ok(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  set_error("%s: %s", what, cudaGetErrorString(e));
  (void)cudaGetLastError();
  return false;
}

bool connected(const etpgt_comm_t* comm) {
  for (int r = 0; r < comm->world; ++r)
    if (comm->base[r] == nullptr) return false;
  return true;
}

}  // namespace
}  // namespace etpgt

using namespace etpgt;

extern "C" size_t etpgt_comm_control_bytes(void) { return align_up(sizeof(CommControl), 65536); }

extern "C" int etpgt_comm_create(int rank, int world, size_t region_bytes, etpgt_comm_t** out) {
  ETPGT_REQUIRE(out != nullptr, "comm_create: null output");
  ETPGT_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "comm_create: rank %d of %d (max %d)",
                rank, world, kMaxRanks);
  ETPGT_REQUIRE(region_bytes >= etpgt_comm_control_bytes(), "comm_create: region %zu < control block %zu",
                region_bytes, etpgt_comm_control_bytes());
  etpgt_comm_t* comm = new etpgt_comm();
  comm->rank = rank;
  comm->world = world;
  comm->region_bytes = region_bytes;
  void* p = nullptr;
  if (!cuda_ok(cudaGetDevice(&comm->device), "comm_create") || !cuda_ok(cudaMalloc(&p, region_bytes), "comm_create: cudaMalloc") ||
      !cuda_ok(cudaMemset(p, 0, etpgt_comm_control_bytes()), "comm_create: cudaMemset")) {
    if (p) cudaFree(p);
    delete comm;
    return ETPGT_ECUDA;
  }
  comm->base[rank] = p;
  *out = comm;
  return ETPGT_OK;
}

extern "C" int etpgt_comm_ipc_handle(const etpgt_comm_t* comm, unsigned char* out) {
  ETPGT_REQUIRE(comm && out, "comm_ipc_handle: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == ETPGT_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  if (!cuda_ok(cudaIpcGetMemHandle(&h, comm->base[comm->rank]), "comm_ipc_handle")) return ETPGT_ECUDA;
  memcpy(out, &h, sizeof(h));
  return ETPGT_OK;
}

extern "C" int etpgt_comm_connect_ipc(etpgt_comm_t* comm, const unsigned char* handles) {
  ETPGT_REQUIRE(comm && handles, "comm_connect_ipc: null argument");
  for (int r = 0; r < comm->world; ++r) {
    if (r == comm->rank || comm->base[r] != nullptr) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * ETPGT_IPC_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    if (!cuda_ok(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "comm_connect_ipc: cudaIpcOpenMemHandle"))
      return ETPGT_ECUDA;
    comm->base[r] = p;
    comm->opened[r] = true;
  }
  return ETPGT_OK;
}

extern "C" int etpgt_comm_connect_ptrs(etpgt_comm_t* comm, void* const* bases) {
  ETPGT_REQUIRE(comm && bases, "comm_connect_ptrs: null argument");
  for (int r = 0; r < comm->world; ++r) {
    if (r == comm->rank) continue;
    ETPGT_REQUIRE(bases[r] != nullptr, "comm_connect_ptrs: rank %d has no region", r);
    comm->base[r] = bases[r];
  }
  return ETPGT_OK;
}

extern "C" void* etpgt_comm_region(const etpgt_comm_t* comm, int rank) {
  if (comm == nullptr || rank < 0 || rank >= comm->world) return nullptr;
  return comm->base[rank];
}

extern "C" int etpgt_comm_set_timeout(etpgt_comm_t* comm, double seconds) {
  ETPGT_REQUIRE(comm && seconds > 0.0 && seconds < 3600.0, "comm_set_timeout: bad arguments");
  comm->timeout_ns = (unsigned long long)(seconds * 1e9);
  return ETPGT_OK;
}

extern "C" int etpgt_comm_destroy(etpgt_comm_t* comm) {
  if (comm == nullptr) return ETPGT_OK;
  for (int r = 0; r < comm->world; ++r)
    if (comm->opened[r] && comm->base[r]) cudaIpcCloseMemHandle(comm->base[r]);
  if (comm->base[comm->rank]) cudaFree(comm->base[comm->rank]);
  (void)cudaGetLastError();
  delete comm;
  return ETPGT_OK;
}

extern "C" int etpgt_comm_barrier(const etpgt_comm_t* comm, int channel, etpgt_stream_t stream) {
  ETPGT_REQUIRE(comm && connected(comm), "comm_barrier: communicator not connected");
  ETPGT_REQUIRE(channel >= 0 && channel < kBarChannels, "comm_barrier: channel %d must be 0..%d", channel, kBarChannels - 1);
  comm_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(comm_view(comm), channel);
  ETPGT_CHECK_LAUNCH("comm_barrier");
  return ETPGT_OK;
}

extern "C" int etpgt_comm_allreduce_f64(const etpgt_comm_t* comm, const double* in, double* out, int count,
                                        etpgt_stream_t stream) {
  ETPGT_REQUIRE(comm && connected(comm), "comm_allreduce_f64: communicator not connected");
  ETPGT_REQUIRE(in && out && count >= 1 && count <= kArMaxCount, "comm_allreduce_f64: count %d must be 1..%d", count,
                kArMaxCount);
  comm_allreduce_f64_kernel<<<1, 512, 0, static_cast<cudaStream_t>(stream)>>>(comm_view(comm), in, out, count);
  ETPGT_CHECK_LAUNCH("comm_allreduce_f64");
  return ETPGT_OK;
}

extern "C" int etpgt_comm_sum_f32(const etpgt_comm_t* comm, size_t offset, int64_t numel, float* out,
                                  etpgt_stream_t stream) {
  ETPGT_REQUIRE(comm && connected(comm), "comm_sum_f32: communicator not connected");
  ETPGT_REQUIRE(numel >= 0 && (numel == 0 || out != nullptr) && offset % 16 == 0 && ((uintptr_t)out & 15) == 0 &&
                    offset + (size_t)numel * 4 <= comm->region_bytes,
                "comm_sum_f32: bad arguments");
  if (numel == 0) return ETPGT_OK;
  comm_sum_f32_kernel<<<grid_for(numel / 4 + 1, 256 * 2, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      comm_view(comm), offset, numel, out);
  ETPGT_CHECK_LAUNCH("comm_sum_f32");
  return ETPGT_OK;
}

extern "C" int etpgt_comm_status(const etpgt_comm_t* comm, int* status) {
  ETPGT_REQUIRE(comm && status, "comm_status: null argument");
  unsigned int s = 0;
  const char* word = static_cast<const char*>(comm->base[comm->rank]) + offsetof(CommControl, status);
  if (!cuda_ok(cudaMemcpy(&s, word, sizeof(s), cudaMemcpyDeviceToHost), "comm_status")) return ETPGT_ECUDA;
  *status = (int)s;
  return ETPGT_OK;
}
