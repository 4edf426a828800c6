// a4: item-embedding gather fused with the Laplacian-PE projection.
//   out[n] = table[ids[n]] + pe[row(n)] @ w_pe^T + b_pe
// etpgt/model/graph_transformer.py:140-152, etpgt/encodings/laplacian_pe.py:170-199.
//
// Gather-bound: per node one table row (DIM*4 B, 128-bit loads, a lane group per row), one PE
// row (k_pe*4 B, broadcast) and one output row.  w_pe^T (k_pe x DIM) and b_pe are staged once
// per CTA in shared memory so the projection costs no global traffic.
#include <cuda_bf16.h>

#include "common.cuh"

namespace etpgt {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxKpe = 64;

template <int DIM>
__global__ void __launch_bounds__(kThreads)
embed_pe_fwd_kernel(const int64_t* __restrict__ ids, int64_t n, const float* __restrict__ table,
                    const float* __restrict__ pe, int pe_per_node, const float* __restrict__ w_pe,
                    const float* __restrict__ b_pe, int k_pe, float* __restrict__ out,
                    __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  using G = RowGeom<DIM>;
  extern __shared__ float smem[];  // w^T [k_pe][DIM] then bias [DIM]
  float* wt = smem;
  float* bias = smem + (size_t)k_pe * DIM;
  if (pe != nullptr) {
    for (int i = threadIdx.x; i < k_pe * DIM; i += kThreads) {
      const int d = i / k_pe, k = i % k_pe;  // w_pe is [DIM][k_pe] (nn.Linear weight)
      wt[k * DIM + d] = w_pe[i];
    }
    for (int d = threadIdx.x; d < DIM; d += kThreads) bias[d] = b_pe[d];
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int lig = lane % G::LPN;
  const int64_t groups_per_cta = (kThreads / 32) * G::GROUPS;
  const int64_t group0 = blockIdx.x * groups_per_cta + (threadIdx.x >> 5) * G::GROUPS + lane / G::LPN;
  const int64_t stride = (int64_t)gridDim.x * groups_per_cta;
  const unsigned gmask = G::LPN == 32 ? 0xffffffffu : (((1u << (G::LPN % 32)) - 1u) << ((lane / G::LPN) * G::LPN));
  const int src0 = (lane / G::LPN) * G::LPN;
  const bool quad_pe = pe != nullptr && (k_pe & 3) == 0 && (k_pe >> 2) <= G::LPN;
  const int quads = k_pe >> 2;
  // NB nodes per group and iteration: all id loads, then all table rows and PE rows are in flight together
  // (the chain id -> row is two dependent HBM round trips), and every W_pe^T float4 read from shared
  // memory is applied to NB nodes (the projection is shared-memory-bandwidth bound otherwise: 32 LDS.128 per
  // node and lane).
  constexpr int NB = 4;
  for (int64_t node0 = group0; node0 < n; node0 += NB * stride) {
    int64_t node[NB], id[NB];
    bool on[NB];
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      node[u] = node0 + u * stride;
      on[u] = node[u] < n;
      id[u] = on[u] ? ids[node[u]] : 0;
    }
    float4 acc[NB][G::V], mine[NB];
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      const float* trow = table + id[u] * DIM;
#pragma unroll
      for (int v = 0; v < G::V; ++v) acc[u][v] = ldg4(trow + 4 * (v * G::LPN + lig));
      mine[u] = zero4();
      if (quad_pe && lig < quads && on[u]) mine[u] = ldg4(pe + (pe_per_node ? node[u] : id[u]) * (int64_t)k_pe + 4 * lig);
    }
    if (pe != nullptr) {
#pragma unroll
      for (int v = 0; v < G::V; ++v) {
        const float4 b = ld4(bias + 4 * (v * G::LPN + lig));
#pragma unroll
        for (int u = 0; u < NB; ++u) acc[u][v] = add4(acc[u][v], b);
      }
      if (quad_pe) {
        for (int j = 0; j < quads; ++j) {
          float p[NB][4];
#pragma unroll
          for (int u = 0; u < NB; ++u) {   // the PE row was loaded once (lane j holds floats 4j..4j+3): broadcast
            p[u][0] = __shfl_sync(gmask, mine[u].x, src0 + j);
            p[u][1] = __shfl_sync(gmask, mine[u].y, src0 + j);
            p[u][2] = __shfl_sync(gmask, mine[u].z, src0 + j);
            p[u][3] = __shfl_sync(gmask, mine[u].w, src0 + j);
          }
          const float* w = wt + (size_t)(4 * j) * DIM;
#pragma unroll
          for (int v = 0; v < G::V; ++v) {
            const int c = 4 * (v * G::LPN + lig);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 wq = ld4(w + q * DIM + c);
#pragma unroll
              for (int u = 0; u < NB; ++u) acc[u][v] = fma4(p[u][q], wq, acc[u][v]);
            }
          }
        }
      } else {
#pragma unroll
        for (int u = 0; u < NB; ++u) {
          const float* prow = pe + (pe_per_node ? node[u] : id[u]) * (int64_t)k_pe;
          for (int k = 0; k < k_pe && on[u]; ++k) {
            const float pk = __ldg(prow + k);
#pragma unroll
            for (int v = 0; v < G::V; ++v) acc[u][v] = fma4(pk, ld4(wt + k * DIM + 4 * (v * G::LPN + lig)), acc[u][v]);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      if (on[u]) {
        float* orow = out + node[u] * DIM;
#pragma unroll
        for (int v = 0; v < G::V; ++v) st4(orow + 4 * (v * G::LPN + lig), acc[u][v]);
        if (out_hi != nullptr) {   // the first layer's projection operand, split for the tensor cores (x = hi + lo)
#pragma unroll
          for (int v = 0; v < G::V; ++v) {
            const float4 o = acc[u][v];
            const __nv_bfloat162 h0 = __floats2bfloat162_rn(o.x, o.y), h1 = __floats2bfloat162_rn(o.z, o.w);
            const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
            const __nv_bfloat162 l0 = __floats2bfloat162_rn(o.x - f0.x, o.y - f0.y),
                                 l1 = __floats2bfloat162_rn(o.z - f1.x, o.w - f1.y);
            uint2 ph, pl;
            ph.x = *reinterpret_cast<const uint32_t*>(&h0); ph.y = *reinterpret_cast<const uint32_t*>(&h1);
            pl.x = *reinterpret_cast<const uint32_t*>(&l0); pl.y = *reinterpret_cast<const uint32_t*>(&l1);
            const int64_t at = node[u] * DIM + 4 * (v * G::LPN + lig);
            *reinterpret_cast<uint2*>(out_hi + at) = ph;
            *reinterpret_cast<uint2*>(out_lo + at) = pl;
          }
        }
      }
    }
  }
}

// d_w_pe[d][k] = sum_n d_out[n][d] * pe[row(n)][k];  d_b_pe[d] = sum_n d_out[n][d].
// Stage 1: CTA c owns a contiguous chunk of nodes; thread d accumulates k_pe+1 sums in
// registers (PE rows staged through shared memory in tiles).  Stage 2 adds the per-CTA partials
// in CTA order (fixed order -> deterministic).
template <int DIM, int KPE>
__global__ void __launch_bounds__(DIM)
pe_wgrad_partial_kernel(const int64_t* __restrict__ ids, int64_t n, const float* __restrict__ d_out,
                        const float* __restrict__ pe, int pe_per_node, int64_t chunk,
                        float* __restrict__ partial /* [grid][DIM][KPE+1] */) {
  constexpr int TILE = 64;
  __shared__ float pe_tile[TILE][KPE];
  const int d = threadIdx.x;
  const int64_t begin = blockIdx.x * chunk;
  const int64_t end = begin + chunk < n ? begin + chunk : n;
  float acc[KPE + 1];
#pragma unroll
  for (int k = 0; k <= KPE; ++k) acc[k] = 0.f;
  for (int64_t base = begin; base < end; base += TILE) {
    const int rows = (int)(end - base < TILE ? end - base : TILE);
    __syncthreads();
    for (int i = threadIdx.x; i < rows * KPE; i += DIM) {
      const int r = i / KPE, k = i % KPE;
      const int64_t prow = pe_per_node ? base + r : ids[base + r];
      pe_tile[r][k] = pe[prow * KPE + k];
    }
    __syncthreads();
    for (int r0 = 0; r0 < rows; r0 += 8) {   // eight independent row loads in flight per thread
      float g[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) g[u] = r0 + u < rows ? d_out[(base + r0 + u) * DIM + d] : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (r0 + u < rows) {
#pragma unroll
          for (int k = 0; k < KPE; ++k) acc[k] = fmaf(g[u], pe_tile[r0 + u][k], acc[k]);
          acc[KPE] += g[u];
        }
      }
    }
  }
  float* dst = partial + ((int64_t)blockIdx.x * DIM + d) * (KPE + 1);
#pragma unroll
  for (int k = 0; k <= KPE; ++k) dst[k] = acc[k];
}

constexpr int kWgradReduceWarps = 32;
__global__ void __launch_bounds__(kWgradReduceWarps * 32)
pe_wgrad_reduce_kernel(const float* __restrict__ partial, int parts, int dim, int kpe,
                       float* __restrict__ d_w, float* __restrict__ d_b) {
  // A CTA owns 32 consecutive outputs (of dim*(kpe+1)): lane = output, so every load is a coalesced
  // 128-byte segment; warp w adds parts w, w+32, ... with eight loads in flight; the 32 warp sums are added
  // in warp order (fixed order -> deterministic).
  __shared__ float warp_sum[kWgradReduceWarps][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int total = dim * (kpe + 1);
  const int i = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (i < total) {
    int p = w;
    for (; p + 7 * kWgradReduceWarps < parts; p += 8 * kWgradReduceWarps) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = partial[(int64_t)(p + kWgradReduceWarps * u) * total + i];
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; p < parts; p += kWgradReduceWarps) s += partial[(int64_t)p * total + i];
  }
  warp_sum[w][lane] = s;
  __syncthreads();
  if (w != 0 || i >= total) return;
  float t = 0.f;
#pragma unroll
  for (int q = 0; q < kWgradReduceWarps; ++q) t += warp_sum[q][lane];
  const int d = i / (kpe + 1), k = i % (kpe + 1);
  if (k == kpe) d_b[d] = t; else d_w[d * kpe + k] = t;
}

int wgrad_parts(int64_t n) {
  int64_t parts = (n + 63) / 64;
  if (parts > 8 * kNumSMs) parts = 8 * kNumSMs;
  return parts < 1 ? 1 : (int)parts;
}

}  // namespace
}  // namespace etpgt

using namespace etpgt;

extern "C" int etpgt_embed_pe_fwd(const int64_t* ids, int64_t n, const float* table, int64_t num_items,
                                  const float* pe, int pe_per_node, const float* w_pe, const float* b_pe,
                                  int k_pe, int dim, float* out, etpgt_stream_t stream) {
  return etpgt_embed_pe_fwd_split(ids, n, table, num_items, pe, pe_per_node, w_pe, b_pe, k_pe, dim, out, nullptr,
                                  nullptr, stream);
}

extern "C" int etpgt_embed_pe_fwd_split(const int64_t* ids, int64_t n, const float* table, int64_t num_items,
                                        const float* pe, int pe_per_node, const float* w_pe, const float* b_pe,
                                        int k_pe, int dim, float* out, void* out_hi, void* out_lo,
                                        etpgt_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ETPGT_REQUIRE((out_hi == nullptr) == (out_lo == nullptr), "embed_pe_fwd: out_hi and out_lo come together");
  ETPGT_REQUIRE(supported_dim(dim), "embed_pe_fwd: unsupported dim %d", dim);
  ETPGT_REQUIRE(n >= 0 && num_items > 0, "embed_pe_fwd: bad size");
  ETPGT_REQUIRE(pe == nullptr || (w_pe && b_pe && k_pe >= 1 && k_pe <= kMaxKpe), "embed_pe_fwd: bad PE arguments");
  if (n == 0) return ETPGT_OK;
  const size_t smem = pe ? ((size_t)k_pe * dim + dim) * sizeof(float) : 0;
#define CALL(D)                                                                                     \
  {                                                                                                 \
    if (smem > 48 * 1024)                                                                           \
      cudaFuncSetAttribute(embed_pe_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    const int64_t gpc = (kThreads / 32) * RowGeom<D>::GROUPS;                                       \
    embed_pe_fwd_kernel<D><<<grid_for(n, (int)gpc * 4, 8), kThreads, smem, stream>>>(                \
        ids, n, table, pe, pe_per_node, w_pe, b_pe, k_pe, out, static_cast<__nv_bfloat16*>(out_hi), \
        static_cast<__nv_bfloat16*>(out_lo));                                                       \
  }
  ETPGT_DISPATCH_DIM(dim, CALL)
#undef CALL
  ETPGT_CHECK_LAUNCH("embed_pe_fwd");
  return ETPGT_OK;
}

extern "C" size_t etpgt_embed_pe_bwd_workspace_bytes(int64_t n, int dim, int k_pe) {
  return etpgt_scatter_rows_workspace_bytes(n) +
         align_up((size_t)wgrad_parts(n) * dim * (k_pe + 1) * sizeof(float)) + 256;
}

extern "C" int etpgt_embed_pe_bwd(const int64_t* ids, int64_t n, const float* d_out, int64_t num_items,
                                  const float* pe, int pe_per_node, int k_pe, int dim, int64_t padding_idx,
                                  float* d_table, float* d_w_pe, float* d_b_pe, void* ws, size_t ws_bytes,
                                  etpgt_stream_t stream) {
  return etpgt_embed_pe_bwd_planned(ids, n, d_out, num_items, pe, pe_per_node, k_pe, dim, padding_idx, nullptr,
                                    nullptr, d_table, d_w_pe, d_b_pe, ws, ws_bytes, stream);
}

extern "C" int etpgt_embed_pe_bwd_planned(const int64_t* ids, int64_t n, const float* d_out, int64_t num_items,
                                          const float* pe, int pe_per_node, int k_pe, int dim,
                                          int64_t padding_idx, const int32_t* plan_sorted_key,
                                          const int32_t* plan_perm, float* d_table, float* d_w_pe, float* d_b_pe,
                                          void* ws, size_t ws_bytes, etpgt_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ETPGT_REQUIRE(supported_dim(dim), "embed_pe_bwd: unsupported dim %d", dim);
  ETPGT_REQUIRE(pe == nullptr || k_pe == 8 || k_pe == 16 || k_pe == 32,
                "embed_pe_bwd: k_pe must be 8, 16 or 32 (got %d)", k_pe);
  if (ws_bytes < etpgt_embed_pe_bwd_workspace_bytes(n, dim, k_pe)) {
    set_error("embed_pe_bwd: workspace too small");
    return ETPGT_EWORKSPACE;
  }
  if (d_table != nullptr && n > 0) {
    // the sort of the node ids may come from a per-batch scatter plan (etpgt_scatter_plan over `ids`)
    int rc = plan_sorted_key != nullptr
                 ? etpgt_scatter_rows_planned(plan_sorted_key, plan_perm, nullptr, d_out, n, 1, dim, padding_idx,
                                              d_table, stream_)
                 : etpgt_scatter_rows(ids, nullptr, d_out, n, 1, dim, num_items, padding_idx, d_table, ws,
                                      etpgt_scatter_rows_workspace_bytes(n), stream_);
    if (rc != ETPGT_OK) return rc;
  }
  if (pe == nullptr) return ETPGT_OK;
  ETPGT_REQUIRE(d_w_pe && d_b_pe, "embed_pe_bwd: null PE gradient outputs");
  float* partial = reinterpret_cast<float*>(static_cast<char*>(ws) + etpgt_scatter_rows_workspace_bytes(n));
  const int parts = wgrad_parts(n);
  const int64_t chunk = n > 0 ? (n + parts - 1) / parts : 0;
#define CALL_K(D, K) \
  pe_wgrad_partial_kernel<D, K><<<parts, D, 0, stream>>>(ids, n, d_out, pe, pe_per_node, chunk, partial);
#define CALL(D)                                   \
  {                                               \
    if (k_pe == 8) { CALL_K(D, 8) }               \
    else if (k_pe == 16) { CALL_K(D, 16) }        \
    else { CALL_K(D, 32) }                        \
  }
  ETPGT_DISPATCH_DIM(dim, CALL)
#undef CALL
#undef CALL_K
  ETPGT_CHECK_LAUNCH("pe_wgrad_partial");
  const int total = dim * (k_pe + 1);
  pe_wgrad_reduce_kernel<<<(total + 31) / 32, kWgradReduceWarps * 32, 0, stream>>>(partial, parts, dim, k_pe, d_w_pe,
                                                                                 d_b_pe);
  ETPGT_CHECK_LAUNCH("pe_wgrad_reduce");
  return ETPGT_OK;
}
