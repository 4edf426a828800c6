// co-occurrence ("co-event") graph construction on the device — the step before the
// hot path, scripts/data/04_build_graph.py:25-127 (`build_co_event_graph`): for every session, every
// pair of events at most `window` steps apart contributes one co-occurrence to the undirected edge
// (min(item), max(item)); an edge carries its count and the largest timestamp seen (the timestamp of
// the event that holds the smaller item id, 04_build_graph.py:64-71); edges are written sorted by
// count, descending.  The reference does this with a Python dict over ~4 M pairs (minutes at 82k items,
// intractable at the 1M-item / 20M-edge configuration).
//
// Here: one thread per (event, offset) slot emits a packed (min,max) key -> stable radix sort -> run
// heads give (edge, count, first emission index, max timestamp) -> a second radix sort orders the edges
// by (count desc, first emission asc).  The tie order is the STABLE order of the reference's dict
// (insertion = first emission); pandas' default quicksort leaves ties unspecified, so the oracle
// (oracle/graph_ref.py co_event_graph) pins the stable order.  `event_pair_hist` (a per-edge dict of
// event-type strings) is not produced: nothing on the training path reads it (dataloader.py:43-48 uses
// item_i / item_j only).
//
// Integer work, HBM-bound: two 64-bit radix sorts over events*window slots plus streaming passes.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace etpgt {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ int64_t session_of(const int64_t* __restrict__ sess_ptr, int64_t num_sessions, int64_t ev) {
  int64_t lo = 0, hi = num_sessions;  // last s with sess_ptr[s] <= ev
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (sess_ptr[mid] <= ev) lo = mid; else hi = mid;
  }
  return lo;
}

// slot p = event * window + (d - 1): the pair (event, event + d) if both lie in the same session
__global__ void __launch_bounds__(kThreads)
cooc_emit_kernel(const int64_t* __restrict__ sess_ptr, const int64_t* __restrict__ items,
                 const int64_t* __restrict__ timestamps, int64_t num_sessions, int64_t num_events, int window,
                 int item_bits, uint64_t* __restrict__ keys, uint32_t* __restrict__ slot, int64_t* __restrict__ ts_out) {
  const int64_t total = num_events * window;
  const uint64_t sentinel = uint64_t(1) << (2 * item_bits);
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ev = p / window;
    const int d = (int)(p - ev * window) + 1;
    const int64_t s = session_of(sess_ptr, num_sessions, ev);
    uint64_t key = sentinel;
    int64_t ts = 0;
    if (ev + d < sess_ptr[s + 1]) {
      const int64_t a = items[ev], b = items[ev + d];
      const bool swap = a > b;                       // 04_build_graph.py:64-71
      key = ((uint64_t)(swap ? b : a) << item_bits) | (uint64_t)(swap ? a : b);
      if (timestamps != nullptr) ts = swap ? timestamps[ev + d] : timestamps[ev];
    }
    keys[p] = key;
    slot[p] = (uint32_t)p;
    if (ts_out != nullptr) ts_out[p] = ts;
  }
}

// head[p] = 1 when sorted position p starts a run of a real (non-sentinel) key
__global__ void __launch_bounds__(kThreads)
cooc_heads_kernel(const uint64_t* __restrict__ keys, int64_t total, uint64_t sentinel, int32_t* __restrict__ head) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x)
    head[p] = (keys[p] != sentinel && (p == 0 || keys[p] != keys[p - 1])) ? 1 : 0;
}

// padding of the ordering arrays (positions that hold no run record sort to the end) + the edge count
__global__ void __launch_bounds__(kThreads)
cooc_padding_kernel(const int32_t* __restrict__ head, const int32_t* __restrict__ run_of, int64_t total,
                    uint64_t* __restrict__ order_key, uint32_t* __restrict__ order_val,
                    int64_t* __restrict__ num_edges) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
    if (p == total - 1) *num_edges = (int64_t)run_of[p] + head[p];
    order_key[p] = ~uint64_t(0);
    order_val[p] = 0;
  }
}

// one thread per run head: walks its run (count, max timestamp) and writes the run record r plus its
// ordering key (count desc, first emission asc)
__global__ void __launch_bounds__(kThreads)
cooc_records_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ slot,
                    const int64_t* __restrict__ ts, const int32_t* __restrict__ head,
                    const int32_t* __restrict__ run_of, int64_t total, uint64_t* __restrict__ run_key,
                    int64_t* __restrict__ run_count, int64_t* __restrict__ run_ts, uint64_t* __restrict__ order_key,
                    uint32_t* __restrict__ order_val) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
    if (!head[p]) continue;
    const uint64_t key = keys[p];
    int64_t q = p, best_ts = INT64_MIN;
    while (q < total && keys[q] == key) {
      if (ts != nullptr) { const int64_t t = ts[slot[q]]; best_ts = t > best_ts ? t : best_ts; }
      ++q;
    }
    const int64_t count = q - p;
    const int32_t r = run_of[p];
    run_key[r] = key;
    run_count[r] = count;
    // the reference starts last_ts at 0 and takes max(): negative timestamps clamp to 0
    run_ts[r] = ts != nullptr ? (best_ts > 0 ? best_ts : 0) : 0;
    // stable sort kept emission order inside a run, so slot[p] is the first emission of this edge
    order_key[r] = ((uint64_t)(0xFFFFFFFFu - (uint32_t)count) << 32) | (uint64_t)slot[p];
    order_val[r] = (uint32_t)r;
  }
}

__global__ void __launch_bounds__(kThreads)
cooc_gather_kernel(const uint32_t* __restrict__ order, const uint64_t* __restrict__ run_key,
                   const int64_t* __restrict__ run_count, const int64_t* __restrict__ run_ts,
                   const int64_t* __restrict__ num_edges, int64_t capacity, int item_bits,
                   int64_t* __restrict__ item_i, int64_t* __restrict__ item_j, int64_t* __restrict__ count,
                   int64_t* __restrict__ last_ts) {
  const int64_t n = *num_edges < capacity ? *num_edges : capacity;
  const uint64_t mask = (uint64_t(1) << item_bits) - 1;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t r = order[e];
    const uint64_t key = run_key[r];
    item_i[e] = (int64_t)(key >> item_bits);
    item_j[e] = (int64_t)(key & mask);
    count[e] = run_count[r];
    if (last_ts != nullptr) last_ts[e] = run_ts[r];
  }
}

int bits_for(int64_t n) {
  int bits = 1;
  while ((int64_t(1) << bits) < n && bits < 31) ++bits;
  return bits;
}

size_t sort_temp_bytes(int64_t n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, n, 0, 64);
  return bytes;
}

size_t scan_temp_bytes(int64_t n) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr, n);
  return bytes;
}

}  // namespace
}  // namespace etpgt

using namespace etpgt;

extern "C" size_t etpgt_cooc_graph_workspace_bytes(int64_t num_events, int window) {
  const int64_t total = (num_events > 0 ? num_events : 1) * (int64_t)(window > 0 ? window : 1);
  const size_t temp = sort_temp_bytes(total) > scan_temp_bytes(total) ? sort_temp_bytes(total) : scan_temp_bytes(total);
  // keys x2, slots x2, ts, head, run_of, run_key, run_count, run_ts, order keys x2, order vals x2
  return 2 * align_up(total * 8) + 2 * align_up(total * 4) + align_up(total * 8) + 2 * align_up(total * 4) +
         3 * align_up(total * 8) + 2 * align_up(total * 8) + 2 * align_up(total * 4) + align_up(temp) + 256;
}

extern "C" int etpgt_cooc_graph_build(const int64_t* sess_ptr, const int64_t* sess_items, const int64_t* timestamps,
                                      int64_t num_sessions, int64_t num_events, int window, int64_t num_items,
                                      int64_t capacity, int64_t* item_i, int64_t* item_j, int64_t* count,
                                      int64_t* last_ts, int64_t* num_edges, void* ws, size_t ws_bytes,
                                      etpgt_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ETPGT_REQUIRE(num_sessions >= 0 && num_events >= 0 && window >= 1 && window <= 64 && num_items >= 1 &&
                    num_items < (int64_t(1) << 31),
                "cooc_graph_build: bad sizes (window in [1, 64], num_items < 2^31)");
  ETPGT_REQUIRE(num_events * (int64_t)window < (int64_t(1) << 31), "cooc_graph_build: events * window must fit int32");
  ETPGT_REQUIRE(num_edges != nullptr && capacity >= 0 &&
                    (num_events == 0 || (sess_ptr && sess_items && (capacity == 0 || (item_i && item_j && count)))),
                "cooc_graph_build: null pointer");
  ETPGT_REQUIRE(last_ts == nullptr || timestamps != nullptr, "cooc_graph_build: last_ts needs timestamps");
  if (ws_bytes < etpgt_cooc_graph_workspace_bytes(num_events, window)) {
    set_error("cooc_graph_build: workspace %zu < %zu", ws_bytes, etpgt_cooc_graph_workspace_bytes(num_events, window));
    return ETPGT_EWORKSPACE;
  }
  if (num_events == 0 || num_sessions == 0) {
    cudaMemsetAsync(num_edges, 0, sizeof(int64_t), stream);
    return ETPGT_OK;
  }
  const int64_t total = num_events * window;
  const int item_bits = bits_for(num_items);
  const uint64_t sentinel = uint64_t(1) << (2 * item_bits);
  Workspace w(ws, ws_bytes);
  uint64_t* keys_a = w.take<uint64_t>(total);
  uint64_t* keys_b = w.take<uint64_t>(total);
  uint32_t* slot_a = w.take<uint32_t>(total);
  uint32_t* slot_b = w.take<uint32_t>(total);
  int64_t* ts = w.take<int64_t>(total);
  int32_t* head = w.take<int32_t>(total);
  int32_t* run_of = w.take<int32_t>(total);
  uint64_t* run_key = w.take<uint64_t>(total);
  int64_t* run_count = w.take<int64_t>(total);
  int64_t* run_ts = w.take<int64_t>(total);
  uint64_t* okey_a = w.take<uint64_t>(total);
  uint64_t* okey_b = w.take<uint64_t>(total);
  uint32_t* oval_a = w.take<uint32_t>(total);
  uint32_t* oval_b = w.take<uint32_t>(total);
  size_t temp_bytes = sort_temp_bytes(total) > scan_temp_bytes(total) ? sort_temp_bytes(total) : scan_temp_bytes(total);
  void* temp = w.take<char>(temp_bytes);
  const int grid = grid_for(total, kThreads, 8);

  cooc_emit_kernel<<<grid, kThreads, 0, stream>>>(sess_ptr, sess_items, timestamps, num_sessions, num_events, window,
                                                  item_bits, keys_a, slot_a, timestamps ? ts : nullptr);
  ETPGT_CHECK_LAUNCH("cooc_emit");
  cudaError_t err = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_a, keys_b, slot_a, slot_b, total, 0,
                                                    2 * item_bits + 1, stream);
  if (err != cudaSuccess) { set_error("cooc sort 1: %s", cudaGetErrorString(err)); return ETPGT_ECUDA; }
  count_launch(4);
  cooc_heads_kernel<<<grid, kThreads, 0, stream>>>(keys_b, total, sentinel, head);
  ETPGT_CHECK_LAUNCH("cooc_heads");
  err = cub::DeviceScan::ExclusiveSum(temp, temp_bytes, head, run_of, total, stream);
  if (err != cudaSuccess) { set_error("cooc scan: %s", cudaGetErrorString(err)); return ETPGT_ECUDA; }
  count_launch(2);
  cooc_padding_kernel<<<grid, kThreads, 0, stream>>>(head, run_of, total, okey_a, oval_a, num_edges);
  ETPGT_CHECK_LAUNCH("cooc_padding");
  cooc_records_kernel<<<grid, kThreads, 0, stream>>>(keys_b, slot_b, timestamps ? ts : nullptr, head, run_of, total,
                                                     run_key, run_count, run_ts, okey_a, oval_a);
  ETPGT_CHECK_LAUNCH("cooc_records");
  err = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, okey_a, okey_b, oval_a, oval_b, total, 0, 64, stream);
  if (err != cudaSuccess) { set_error("cooc sort 2: %s", cudaGetErrorString(err)); return ETPGT_ECUDA; }
  count_launch(8);
  cooc_gather_kernel<<<grid, kThreads, 0, stream>>>(oval_b, run_key, run_count, run_ts, num_edges, capacity, item_bits,
                                                    item_i, item_j, count, last_ts);
  ETPGT_CHECK_LAUNCH("cooc_gather");
  return ETPGT_OK;
}
