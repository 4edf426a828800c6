// a7: session readout as a segmented reduction over contiguous node ranges.
// etpgt/model/base.py:136-193 (the reference loops over sessions in Python with a boolean
// mask per session).  One lane group per session; mean / max / last / softmax-weighted sum.
// Streaming, HBM-bound: reads N*DIM*4 B, writes S*DIM*4 B.
#include <math.h>

#include "common.cuh"

namespace etpgt {
namespace {

constexpr int kThreads = 256;

template <int DIM>
__global__ void __launch_bounds__(kThreads)
readout_fwd_kernel(const float* __restrict__ x, const int32_t* __restrict__ ptr, int64_t num_sessions, int mode,
                   const float* __restrict__ scores, float* __restrict__ out, int32_t* __restrict__ argmax,
                   float* __restrict__ weights) {
  using G = RowGeom<DIM>;
  constexpr int V = G::V, LPN = G::LPN;
  const int lane = threadIdx.x & 31;
  const int lig = lane % LPN;
  const int64_t s = ((blockIdx.x * (int64_t)kThreads + threadIdx.x) >> 5) * G::GROUPS + lane / LPN;
  if (s >= num_sessions) return;  // no warp-wide collectives below
  const int begin = ptr[s], end = ptr[s + 1];
  float4 acc[V];
  if (mode == ETPGT_READOUT_MEAN) {
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = zero4();
    for (int r = begin; r < end; ++r)
#pragma unroll
      for (int v = 0; v < V; ++v) acc[v] = add4(acc[v], ldg4(x + (int64_t)r * DIM + 4 * (v * LPN + lig)));
    // mean over an empty range is 0/0 = NaN in the reference as well (base.py:155)
    const float inv = 1.f / (float)(end - begin);
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = scale4(inv, acc[v]);
  } else if (mode == ETPGT_READOUT_MAX) {
    int4 arg[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { acc[v] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY); arg[v] = make_int4(begin, begin, begin, begin); }
    for (int r = begin; r < end; ++r)
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float4 a = ldg4(x + (int64_t)r * DIM + 4 * (v * LPN + lig));
        if (a.x > acc[v].x) { acc[v].x = a.x; arg[v].x = r; }
        if (a.y > acc[v].y) { acc[v].y = a.y; arg[v].y = r; }
        if (a.z > acc[v].z) { acc[v].z = a.z; arg[v].z = r; }
        if (a.w > acc[v].w) { acc[v].w = a.w; arg[v].w = r; }
      }
#pragma unroll
    for (int v = 0; v < V; ++v) *reinterpret_cast<int4*>(argmax + s * DIM + 4 * (v * LPN + lig)) = arg[v];
  } else if (mode == ETPGT_READOUT_LAST) {
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = ldg4(x + (int64_t)(end - 1) * DIM + 4 * (v * LPN + lig));
  } else {  // attention: softmax(scores[begin:end]) @ x[begin:end]
    float mx = -INFINITY;
    for (int r = begin; r < end; ++r) mx = fmaxf(mx, scores[r]);
    float denom = 0.f;
    for (int r = begin; r < end; ++r) denom += expf(scores[r] - mx);
    const float inv = 1.f / denom;
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = zero4();
    for (int r = begin; r < end; ++r) {
      const float wgt = expf(scores[r] - mx) * inv;
      if (lig == 0) weights[r] = wgt;
#pragma unroll
      for (int v = 0; v < V; ++v) acc[v] = fma4(wgt, ldg4(x + (int64_t)r * DIM + 4 * (v * LPN + lig)), acc[v]);
    }
  }
#pragma unroll
  for (int v = 0; v < V; ++v) st4(out + s * DIM + 4 * (v * LPN + lig), acc[v]);
}

// One lane group per session again: every node row of the session is written exactly once.
template <int DIM>
__global__ void __launch_bounds__(kThreads)
readout_bwd_kernel(const float* __restrict__ x, const float* __restrict__ out, const float* __restrict__ d_out,
                   const int32_t* __restrict__ ptr, int64_t num_sessions, int mode,
                   const int32_t* __restrict__ argmax, const float* __restrict__ weights,
                   float* __restrict__ d_x, float* __restrict__ d_scores) {
  using G = RowGeom<DIM>;
  constexpr int V = G::V, LPN = G::LPN;
  const int lane = threadIdx.x & 31;
  const int lig = lane % LPN;
  const int64_t s0 = ((blockIdx.x * (int64_t)kThreads + threadIdx.x) >> 5) * G::GROUPS;
  if (s0 >= num_sessions) return;  // warp-uniform
  const int64_t s = s0 + lane / LPN;
  const bool valid = s < num_sessions;
  const int64_t srow = valid ? s : 0;
  const int begin = valid ? ptr[srow] : 0, end = valid ? ptr[srow + 1] : 0;
  float4 g[V];
#pragma unroll
  for (int v = 0; v < V; ++v) g[v] = ldg4(d_out + srow * DIM + 4 * (v * LPN + lig));
  if (mode == ETPGT_READOUT_MEAN) {
    const float inv = 1.f / (float)(end - begin);
    for (int r = begin; r < end; ++r)
#pragma unroll
      for (int v = 0; v < V; ++v) st4(d_x + (int64_t)r * DIM + 4 * (v * LPN + lig), scale4(inv, g[v]));
  } else if (mode == ETPGT_READOUT_MAX) {
    int4 arg[V];
#pragma unroll
    for (int v = 0; v < V; ++v) arg[v] = *reinterpret_cast<const int4*>(argmax + srow * DIM + 4 * (v * LPN + lig));
    for (int r = begin; r < end; ++r)
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float4 o;
        o.x = arg[v].x == r ? g[v].x : 0.f; o.y = arg[v].y == r ? g[v].y : 0.f;
        o.z = arg[v].z == r ? g[v].z : 0.f; o.w = arg[v].w == r ? g[v].w : 0.f;
        st4(d_x + (int64_t)r * DIM + 4 * (v * LPN + lig), o);
      }
  } else if (mode == ETPGT_READOUT_LAST) {
    for (int r = begin; r < end; ++r)
#pragma unroll
      for (int v = 0; v < V; ++v) st4(d_x + (int64_t)r * DIM + 4 * (v * LPN + lig), r == end - 1 ? g[v] : zero4());
  } else {
    // out = sum_n w_n x_n;  d_w_n = <g, x_n>;  d_score_n = w_n (d_w_n - <g, out>);  d_x_n = w_n g
    float part = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) part += dot4(g[v], ldg4(out + srow * DIM + 4 * (v * LPN + lig)));
    const float g_out = group_sum<LPN>(part);
    int len = end - begin, len_max = len;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) len_max = max(len_max, __shfl_xor_sync(0xffffffffu, len_max, off));
    for (int t = 0; t < len_max; ++t) {
      const bool on = t < len;
      const int r = on ? begin + t : 0;
      float p = 0.f;
      float4 xr[V];
#pragma unroll
      for (int v = 0; v < V; ++v) { xr[v] = ldg4(x + (int64_t)r * DIM + 4 * (v * LPN + lig)); p += dot4(g[v], xr[v]); }
      const float dw = group_sum<LPN>(p);
      if (on) {
        const float wgt = weights[r];
        if (lig == 0) d_scores[r] = wgt * (dw - g_out);
#pragma unroll
        for (int v = 0; v < V; ++v) st4(d_x + (int64_t)r * DIM + 4 * (v * LPN + lig), scale4(wgt, g[v]));
      }
    }
  }
}

}  // namespace
}  // namespace etpgt

using namespace etpgt;

extern "C" int etpgt_readout_fwd(const float* x, const int32_t* ptr, int64_t num_sessions, int dim, int mode,
                                 const float* scores, float* out, void* aux, etpgt_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ETPGT_REQUIRE(supported_dim(dim), "readout_fwd: unsupported dim %d", dim);
  ETPGT_REQUIRE(mode >= 0 && mode <= 3, "Unknown readout type: %d", mode);
  ETPGT_REQUIRE(mode != ETPGT_READOUT_ATTENTION || (scores && aux), "readout_fwd: attention needs scores and aux");
  ETPGT_REQUIRE(mode != ETPGT_READOUT_MAX || aux, "readout_fwd: max needs aux");
  if (num_sessions == 0) return ETPGT_OK;
#define CALL(D)                                                                                      \
  {                                                                                                  \
    const int64_t spc = (kThreads / 32) * RowGeom<D>::GROUPS;                                        \
    readout_fwd_kernel<D><<<(unsigned)((num_sessions + spc - 1) / spc), kThreads, 0, stream>>>(       \
        x, ptr, num_sessions, mode, scores, out, static_cast<int32_t*>(aux), static_cast<float*>(aux)); \
  }
  ETPGT_DISPATCH_DIM(dim, CALL)
#undef CALL
  ETPGT_CHECK_LAUNCH("readout_fwd");
  return ETPGT_OK;
}

extern "C" int etpgt_readout_bwd(const float* x, const float* out, const float* d_out, const int32_t* ptr,
                                 int64_t num_nodes, int64_t num_sessions, int dim, int mode, const void* aux,
                                 float* d_x, float* d_scores, etpgt_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ETPGT_REQUIRE(supported_dim(dim), "readout_bwd: unsupported dim %d", dim);
  ETPGT_REQUIRE(mode >= 0 && mode <= 3, "Unknown readout type: %d", mode);
  ETPGT_REQUIRE(mode != ETPGT_READOUT_ATTENTION || (x && out && aux && d_scores), "readout_bwd: attention needs x/out/aux/d_scores");
  (void)num_nodes;
  if (num_sessions == 0) return ETPGT_OK;
#define CALL(D)                                                                                      \
  {                                                                                                  \
    const int64_t spc = (kThreads / 32) * RowGeom<D>::GROUPS;                                        \
    readout_bwd_kernel<D><<<(unsigned)((num_sessions + spc - 1) / spc), kThreads, 0, stream>>>(       \
        x, out, d_out, ptr, num_sessions, mode, static_cast<const int32_t*>(aux),                    \
        static_cast<const float*>(aux), d_x, d_scores);                                              \
  }
  ETPGT_DISPATCH_DIM(dim, CALL)
#undef CALL
  ETPGT_CHECK_LAUNCH("readout_bwd");
  return ETPGT_OK;
}
