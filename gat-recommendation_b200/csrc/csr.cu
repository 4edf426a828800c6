// Destination-sorted CSR + source-sorted CSC from an int64 COO edge list, and segment
// pointers from a sorted id vector.  Integer work, bit-exact against oracle/graph_ref.py.
//
// HBM-bound: 2 stable radix sorts of E (key,value) int32 pairs (cub::DeviceRadixSort, only
// the bits needed for num_nodes) plus three streaming passes.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace etpgt {
namespace {

constexpr int kThreads = 256;

__global__ void narrow_keys_kernel(const int64_t* __restrict__ key64, int32_t* __restrict__ key32,
                                   int32_t* __restrict__ iota, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    key32[i] = static_cast<int32_t>(key64[i]);
    iota[i] = static_cast<int32_t>(i);
  }
}

// ptr[r] = first position p with sorted_key[p] >= r, for r in [0, num_rows]; thread p fills
// the rows between sorted_key[p-1] and sorted_key[p] (sentinels -1 and num_rows).
template <typename KeyT>
__global__ void boundaries_kernel(const KeyT* __restrict__ sorted_key, int64_t n, int64_t num_rows,
                                  int32_t* __restrict__ ptr) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p <= n; p += (int64_t)gridDim.x * blockDim.x) {
    int64_t prev = p == 0 ? -1 : static_cast<int64_t>(sorted_key[p - 1]);
    int64_t cur = p == n ? num_rows : static_cast<int64_t>(sorted_key[p]);
    if (cur > num_rows) cur = num_rows;
    for (int64_t r = prev + 1; r <= cur; ++r) ptr[r] = static_cast<int32_t>(p);
  }
}

// col[p] = src[eperm[p]]  (also emits the iota for the second sort)
__global__ void gather_src_kernel(const int64_t* __restrict__ src, const int32_t* __restrict__ eperm,
                                  int32_t* __restrict__ col, int32_t* __restrict__ iota, int64_t n) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    col[p] = static_cast<int32_t>(src[eperm[p]]);
    iota[p] = static_cast<int32_t>(p);
  }
}

// row[p] = dst_sorted[cpos[p]]
__global__ void gather_i32_kernel(const int32_t* __restrict__ values, const int32_t* __restrict__ index,
                                  int32_t* __restrict__ out, int64_t n) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x)
    out[p] = values[index[p]];
}

// flag |= 1 when any id of the three lists lies outside [0, num_rows): the device-side counterpart of the
// IndexError nn.Embedding raises in the reference (item ids index the table and its gradient buffer raw)
__global__ void ids_check_kernel(const int64_t* __restrict__ a, int64_t na, const int64_t* __restrict__ b,
                                 int64_t nb, const int64_t* __restrict__ c, int64_t nc, int64_t num_rows,
                                 int32_t* __restrict__ flag) {
  bool bad = false;
  const int64_t total = na + nb + nc;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t id = i < na ? a[i] : (i < na + nb ? b[i - na] : c[i - na - nb]);
    bad = bad || id < 0 || id >= num_rows;
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

// Batch preparation, one sort for three: element i of the combined array is edge i keyed by its destination
// (segment 0), node i - E keyed by its item id (segment 1), or loss key i - E - N (segment 2); the segment
// number sits above the value bits, so ONE stable radix sort orders all three at once (small sorts are bound by
// the fixed cost of a pass, not by their size).
__global__ void combined_keys_kernel(const int64_t* __restrict__ dst, int64_t E, const int64_t* __restrict__ ids,
                                     int64_t N, const int64_t* __restrict__ targets,
                                     const int64_t* __restrict__ negatives, int64_t M, int num_neg, int bits,
                                     int32_t* __restrict__ key, int32_t* __restrict__ val) {
  const int64_t total = E + N + M;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < E) {
      key[i] = static_cast<int32_t>(dst[i]);
      val[i] = static_cast<int32_t>(i);
    } else if (i < E + N) {
      key[i] = (1 << bits) | static_cast<int32_t>(ids[i - E]);
      val[i] = static_cast<int32_t>(i - E);
    } else {
      const int64_t j = i - E - N, s = j / (num_neg + 1);
      const int c = static_cast<int>(j - s * (num_neg + 1));   // [s][0] = target, [s][1 + c] = negative c
      key[i] = (2 << bits) | static_cast<int32_t>(c == 0 ? targets[s] : negatives[s * num_neg + c - 1]);
      val[i] = static_cast<int32_t>(j);
    }
  }
}

// the sorted combined arrays back into the caller's separate outputs (segment bits stripped)
__global__ void split_sorted_kernel(const int32_t* __restrict__ key, const int32_t* __restrict__ val, int64_t E,
                                    int64_t N, int64_t M, int bits, int32_t* __restrict__ dst_sorted,
                                    int32_t* __restrict__ eperm, int32_t* __restrict__ nodes_key,
                                    int32_t* __restrict__ nodes_perm, int32_t* __restrict__ loss_key,
                                    int32_t* __restrict__ loss_perm) {
  const int64_t total = E + N + M;
  const int32_t mask = (1 << bits) - 1;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < E) {
      dst_sorted[i] = key[i];
      eperm[i] = val[i];
    } else if (i < E + N) {
      nodes_key[i - E] = key[i] & mask;
      nodes_perm[i - E] = val[i];
    } else {
      loss_key[i - E - N] = key[i] & mask;
      loss_perm[i - E - N] = val[i];
    }
  }
}

int key_bits(int64_t num_nodes) {
  int bits = 1;
  while ((int64_t(1) << bits) < num_nodes && bits < 31) ++bits;
  return bits;
}

size_t sort_temp_bytes(int64_t n, int bits) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, static_cast<int>(n), 0, bits);
  return bytes;
}

}  // namespace
}  // namespace etpgt

using namespace etpgt;

extern "C" size_t etpgt_csr_workspace_bytes(int64_t num_edges, int64_t num_nodes) {
  int64_t e = num_edges > 0 ? num_edges : 1;
  // cub's temp-size query does not touch the device; 5 int32 arrays of E plus the sort temp
  size_t temp = sort_temp_bytes(e, key_bits(num_nodes));
  return 5 * align_up(e * sizeof(int32_t)) + align_up(temp) + 256;
}

extern "C" int etpgt_csr_from_coo(const int64_t* src, const int64_t* dst, int64_t num_edges, int64_t num_nodes,
                                  int32_t* rowptr, int32_t* col, int32_t* eperm, int32_t* colptr,
                                  int32_t* row, int32_t* cpos, void* ws, size_t ws_bytes,
                                  etpgt_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ETPGT_REQUIRE(num_edges >= 0 && num_nodes >= 0, "csr_from_coo: negative size");
  ETPGT_REQUIRE(num_edges < (int64_t(1) << 31) && num_nodes < (int64_t(1) << 31) - 1,
                "csr_from_coo: sizes must fit int32");
  ETPGT_REQUIRE(rowptr && colptr, "csr_from_coo: null output");
  if (ws_bytes < etpgt_csr_workspace_bytes(num_edges, num_nodes)) {
    set_error("csr_from_coo: workspace %zu < %zu", ws_bytes, etpgt_csr_workspace_bytes(num_edges, num_nodes));
    return ETPGT_EWORKSPACE;
  }
  const int64_t E = num_edges;
  const int grid_e = grid_for(E + 1, kThreads, 8);
  if (E == 0) {
    boundaries_kernel<int32_t><<<1, kThreads, 0, stream>>>(nullptr, 0, num_nodes, rowptr);
    ETPGT_CHECK_LAUNCH("boundaries(rowptr, E=0)");
    boundaries_kernel<int32_t><<<1, kThreads, 0, stream>>>(nullptr, 0, num_nodes, colptr);
    ETPGT_CHECK_LAUNCH("boundaries(colptr, E=0)");
    return ETPGT_OK;
  }
  Workspace w(ws, ws_bytes);
  int32_t* key_a = w.take<int32_t>(E);
  int32_t* key_b = w.take<int32_t>(E);   // dst in CSR order
  int32_t* iota = w.take<int32_t>(E);
  int32_t* key_c = w.take<int32_t>(E);   // src in CSC order
  const int bits = key_bits(num_nodes);
  size_t temp_bytes = sort_temp_bytes(E, bits);
  void* temp = w.take<char>(temp_bytes);

  narrow_keys_kernel<<<grid_e, kThreads, 0, stream>>>(dst, key_a, iota, E);
  ETPGT_CHECK_LAUNCH("narrow_keys");
  cudaError_t err = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, key_a, key_b, iota, eperm,
                                                    static_cast<int>(E), 0, bits, stream);
  if (err != cudaSuccess) { set_error("csr sort 1: %s", cudaGetErrorString(err)); return ETPGT_ECUDA; }
  count_launch(4);
  boundaries_kernel<int32_t><<<grid_e, kThreads, 0, stream>>>(key_b, E, num_nodes, rowptr);
  ETPGT_CHECK_LAUNCH("boundaries(rowptr)");
  gather_src_kernel<<<grid_e, kThreads, 0, stream>>>(src, eperm, col, iota, E);
  ETPGT_CHECK_LAUNCH("gather_src");
  err = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, col, key_c, iota, cpos, static_cast<int>(E), 0, bits,
                                        stream);
  if (err != cudaSuccess) { set_error("csr sort 2: %s", cudaGetErrorString(err)); return ETPGT_ECUDA; }
  count_launch(4);
  boundaries_kernel<int32_t><<<grid_e, kThreads, 0, stream>>>(key_c, E, num_nodes, colptr);
  ETPGT_CHECK_LAUNCH("boundaries(colptr)");
  gather_i32_kernel<<<grid_e, kThreads, 0, stream>>>(key_b, cpos, row, E);
  ETPGT_CHECK_LAUNCH("gather_row");
  return ETPGT_OK;
}

extern "C" size_t etpgt_batch_prepare_workspace_bytes(int64_t num_edges, int64_t num_nodes, int64_t num_loss_keys) {
  const int64_t e = num_edges > 0 ? num_edges : 1;
  const int64_t total = e + (num_nodes > 0 ? num_nodes : 0) + (num_loss_keys > 0 ? num_loss_keys : 0);
  const size_t temp = sort_temp_bytes(total, 31);
  // combined key / value in and out (4 x total), dst in CSR order, iota, src in CSC order (3 x E), sort temp
  return 4 * align_up(total * sizeof(int32_t)) + 3 * align_up(e * sizeof(int32_t)) + align_up(temp) + 256;
}

extern "C" int etpgt_batch_prepare(const int64_t* src, const int64_t* dst, int64_t num_edges, int64_t num_nodes,
                                   const int64_t* ids, const int64_t* targets, const int64_t* negatives,
                                   int64_t num_sessions, int num_neg, int64_t num_items, int32_t* rowptr,
                                   int32_t* col, int32_t* eperm, int32_t* colptr, int32_t* row, int32_t* cpos,
                                   int32_t* nodes_key, int32_t* nodes_perm, int32_t* loss_key, int32_t* loss_perm,
                                   void* ws, size_t ws_bytes, etpgt_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool plan_loss = targets != nullptr && negatives != nullptr && num_sessions > 0;
  const int64_t E = num_edges, N = num_nodes, M = plan_loss ? num_sessions * (num_neg + 1) : 0;
  ETPGT_REQUIRE(E >= 0 && N >= 1 && num_items >= 1 && num_neg >= 0, "batch_prepare: bad sizes");
  ETPGT_REQUIRE(E + N + M < (int64_t(1) << 31) && num_items < (int64_t(1) << 28) && N < (int64_t(1) << 28),
                "batch_prepare: sizes must fit int32 with two segment bits");
  ETPGT_REQUIRE(rowptr && colptr && ids && nodes_key && nodes_perm && (!plan_loss || (loss_key && loss_perm)) &&
                    (E == 0 || (src && dst && col && eperm && row && cpos)),
                "batch_prepare: null pointer");
  if (ws_bytes < etpgt_batch_prepare_workspace_bytes(E, N, M)) {
    set_error("batch_prepare: workspace %zu < %zu", ws_bytes, etpgt_batch_prepare_workspace_bytes(E, N, M));
    return ETPGT_EWORKSPACE;
  }
  const int64_t total = E + N + M;
  const int node_bits = key_bits(N), item_bits = key_bits(num_items);
  const int bits = node_bits > item_bits ? node_bits : item_bits;
  Workspace w(ws, ws_bytes);
  int32_t* key_in = w.take<int32_t>(total);
  int32_t* val_in = w.take<int32_t>(total);
  int32_t* key_out = w.take<int32_t>(total);
  int32_t* val_out = w.take<int32_t>(total);
  int32_t* dst_sorted = w.take<int32_t>(E > 0 ? E : 1);
  int32_t* iota = w.take<int32_t>(E > 0 ? E : 1);
  int32_t* src_sorted = w.take<int32_t>(E > 0 ? E : 1);
  size_t temp_bytes = sort_temp_bytes(total, 31);
  void* temp = w.take<char>(temp_bytes);
  const int grid_t = grid_for(total, kThreads, 8);
  combined_keys_kernel<<<grid_t, kThreads, 0, stream>>>(dst, E, ids, N, targets, negatives, M, num_neg, bits, key_in,
                                                        val_in);
  ETPGT_CHECK_LAUNCH("batch_prepare keys");
  cudaError_t err = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, key_in, key_out, val_in, val_out,
                                                    static_cast<int>(total), 0, bits + 2, stream);
  if (err != cudaSuccess) { set_error("batch_prepare sort: %s", cudaGetErrorString(err)); return ETPGT_ECUDA; }
  count_launch(4);
  split_sorted_kernel<<<grid_t, kThreads, 0, stream>>>(key_out, val_out, E, N, M, bits, dst_sorted, eperm, nodes_key,
                                                       nodes_perm, loss_key, loss_perm);
  ETPGT_CHECK_LAUNCH("batch_prepare split");
  const int grid_e = grid_for(E + 1, kThreads, 8);
  boundaries_kernel<int32_t><<<grid_e, kThreads, 0, stream>>>(E > 0 ? dst_sorted : nullptr, E, N, rowptr);
  ETPGT_CHECK_LAUNCH("boundaries(rowptr)");
  if (E == 0) {
    boundaries_kernel<int32_t><<<1, kThreads, 0, stream>>>(nullptr, 0, N, colptr);
    ETPGT_CHECK_LAUNCH("boundaries(colptr, E=0)");
    return ETPGT_OK;
  }
  // the CSC order is the CSR order stably re-sorted by source, so it has to follow the first sort
  gather_src_kernel<<<grid_e, kThreads, 0, stream>>>(src, eperm, col, iota, E);
  ETPGT_CHECK_LAUNCH("gather_src");
  err = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, col, src_sorted, iota, cpos, static_cast<int>(E), 0,
                                        node_bits, stream);
  if (err != cudaSuccess) { set_error("batch_prepare sort 2: %s", cudaGetErrorString(err)); return ETPGT_ECUDA; }
  count_launch(4);
  boundaries_kernel<int32_t><<<grid_e, kThreads, 0, stream>>>(src_sorted, E, N, colptr);
  ETPGT_CHECK_LAUNCH("boundaries(colptr)");
  gather_i32_kernel<<<grid_e, kThreads, 0, stream>>>(dst_sorted, cpos, row, E);
  ETPGT_CHECK_LAUNCH("gather_row");
  return ETPGT_OK;
}

extern "C" int etpgt_segment_ptr(const int64_t* seg_ids, int64_t n, int64_t num_segments, int32_t* ptr,
                                 etpgt_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ETPGT_REQUIRE(n >= 0 && num_segments >= 0 && n < (int64_t(1) << 31), "segment_ptr: bad size");
  boundaries_kernel<int64_t><<<grid_for(n + 1, kThreads, 8), kThreads, 0, stream>>>(seg_ids, n, num_segments, ptr);
  ETPGT_CHECK_LAUNCH("boundaries(segment_ptr)");
  return ETPGT_OK;
}

extern "C" int etpgt_ids_check(const int64_t* a, int64_t na, const int64_t* b, int64_t nb, const int64_t* c,
                               int64_t nc, int64_t num_rows, int32_t* flag, etpgt_stream_t stream) {
  ETPGT_REQUIRE(na >= 0 && nb >= 0 && nc >= 0 && (na == 0 || a) && (nb == 0 || b) && (nc == 0 || c) && flag &&
                    num_rows >= 0,
                "ids_check: bad arguments");
  const int64_t total = na + nb + nc;
  if (total == 0) return ETPGT_OK;
  ids_check_kernel<<<grid_for(total, kThreads * 4, 4), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      a, na, b, nb, c, nc, num_rows, flag);
  ETPGT_CHECK_LAUNCH("ids_check");
  return ETPGT_OK;
}
