// a8: sampled BPR / listwise / dual loss, forward and backward.
// etpgt/train/losses.py:20-164 and etpgt/model/base.py:80-113: gather the target row and the
// num_neg negative rows of the item table, 1+num_neg dot products per session, then
//   bpr      = mean_{b,n} -log(sigmoid(pos_b - neg_bn) + 1e-8)
//   listwise = mean_b     logsumexp([pos_b, neg_b*]/T) - pos_b/T
//   dual     = alpha*listwise + (1-alpha)*bpr
// Gather-bound: (1+num_neg)*DIM*4 B per session forward, the same again backward.  The loss
// scalars stay on the device (the reference syncs three .item() per step, losses.py:158-162).
#include <math.h>

#include "common.cuh"

namespace etpgt {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxNeg = 64;

template <int DIM>
__global__ void __launch_bounds__(kThreads)
loss_fwd_kernel(const float* __restrict__ sess, const float* __restrict__ table, const int64_t* __restrict__ targets,
                const int64_t* __restrict__ negatives, int64_t batch, int num_neg, float inv_temp,
                float* __restrict__ scores, float2* __restrict__ terms) {
  using G = RowGeom<DIM>;
  constexpr int V = G::V, LPN = G::LPN;
  const int lane = threadIdx.x & 31;
  const int lig = lane % LPN;
  const int64_t b0 = ((blockIdx.x * (int64_t)kThreads + threadIdx.x) >> 5) * G::GROUPS;
  if (b0 >= batch) return;  // warp-uniform
  const int64_t b = b0 + lane / LPN;
  const bool valid = b < batch;
  const int64_t brow = valid ? b : 0;
  float4 s[V];
#pragma unroll
  for (int v = 0; v < V; ++v) s[v] = ldg4(sess + brow * DIM + 4 * (v * LPN + lig));
  float pos = 0.f, bpr = 0.f, mx = -INFINITY, denom = 0.f;
  for (int c = 0; c <= num_neg; ++c) {
    const int64_t id = c == 0 ? targets[brow] : negatives[brow * num_neg + c - 1];
    float part = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) part += dot4(s[v], ldg4(table + id * DIM + 4 * (v * LPN + lig)));
    const float sc = group_sum<LPN>(part);
    if (valid && lig == 0) scores[brow * (num_neg + 1) + c] = sc;
    if (c == 0) pos = sc;
    else {
      const float sg = 1.f / (1.f + expf(-(pos - sc)));
      bpr += -logf(sg + 1e-8f);
    }
    const float z = sc * inv_temp;  // online logsumexp
    const float m_new = fmaxf(mx, z);
    denom = denom * expf(mx - m_new) + expf(z - m_new);
    mx = m_new;
  }
  if (valid && lig == 0) terms[brow] = make_float2(mx + logf(denom) - pos * inv_temp, bpr);
}

// Single CTA, fixed assignment and fixed tree: deterministic.
__global__ void __launch_bounds__(1024)
loss_reduce_kernel(const float2* __restrict__ terms, int64_t batch, int num_neg, int mode, float alpha,
                   double total_sessions, float* __restrict__ losses) {
  __shared__ double red[2][1024];
  double lw = 0, bp = 0;
  for (int64_t i = threadIdx.x; i < batch; i += 1024) { lw += terms[i].x; bp += terms[i].y; }
  red[0][threadIdx.x] = lw;
  red[1][threadIdx.x] = bp;
  __syncthreads();
  for (int off = 512; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) {
      red[0][threadIdx.x] += red[0][threadIdx.x + off];
      red[1][threadIdx.x] += red[1][threadIdx.x + off];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double listwise = red[0][0] / total_sessions;
    const double bpr = red[1][0] / (total_sessions * num_neg);
    double total = mode == ETPGT_LOSS_BPR ? bpr : mode == ETPGT_LOSS_LISTWISE ? listwise
                                                                             : alpha * listwise + (1.0 - alpha) * bpr;
    losses[0] = (float)total;
    losses[1] = (float)listwise;
    losses[2] = (float)bpr;
  }
}

template <int DIM>
__global__ void __launch_bounds__(kThreads)
loss_bwd_kernel(const float* __restrict__ table, const int64_t* __restrict__ targets,
                const int64_t* __restrict__ negatives, int64_t batch, int num_neg, float w_listwise, float w_bpr,
                float inv_temp, double total_sessions, const float* __restrict__ scores,
                const float* __restrict__ d_loss, float* __restrict__ d_sess, int64_t* __restrict__ keys,
                float* __restrict__ coef) {
  using G = RowGeom<DIM>;
  constexpr int V = G::V, LPN = G::LPN;
  const int lane = threadIdx.x & 31;
  const int lig = lane % LPN;
  const int64_t b = ((blockIdx.x * (int64_t)kThreads + threadIdx.x) >> 5) * G::GROUPS + lane / LPN;
  if (b >= batch) return;  // no warp collectives in this kernel
  const float g = *d_loss;
  const float c_lw = g * w_listwise * (float)(1.0 / total_sessions) * inv_temp;
  const float c_bpr = g * w_bpr * (float)(1.0 / (total_sessions * num_neg));
  const float* sc = scores + b * (num_neg + 1);
  const float pos = sc[0];
  float mx = -INFINITY;
  for (int c = 0; c <= num_neg; ++c) mx = fmaxf(mx, sc[c] * inv_temp);
  float denom = 0.f;
  for (int c = 0; c <= num_neg; ++c) denom += expf(sc[c] * inv_temp - mx);
  const float inv_denom = 1.f / denom;
  float4 ds[V];
#pragma unroll
  for (int v = 0; v < V; ++v) ds[v] = zero4();
  float d_pos = (expf(pos * inv_temp - mx) * inv_denom - 1.f) * c_lw;
  // negatives first (their BPR terms also feed d_pos), then the target row
  for (int c = 1; c <= num_neg; ++c) {
    const int64_t id = negatives[b * num_neg + c - 1];
    const float sg = 1.f / (1.f + expf(-(pos - sc[c])));
    const float dfdx = -sg * (1.f - sg) / (sg + 1e-8f) * c_bpr;
    d_pos += dfdx;
    const float d_neg = expf(sc[c] * inv_temp - mx) * inv_denom * c_lw - dfdx;
#pragma unroll
    for (int v = 0; v < V; ++v) ds[v] = fma4(d_neg, ldg4(table + id * DIM + 4 * (v * LPN + lig)), ds[v]);
    if (lig == 0) { keys[b * (num_neg + 1) + c] = id; coef[b * (num_neg + 1) + c] = d_neg; }
  }
  const int64_t tid = targets[b];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    ds[v] = fma4(d_pos, ldg4(table + tid * DIM + 4 * (v * LPN + lig)), ds[v]);
    st4(d_sess + b * DIM + 4 * (v * LPN + lig), ds[v]);
  }
  if (lig == 0) { keys[b * (num_neg + 1)] = tid; coef[b * (num_neg + 1)] = d_pos; }
}

size_t bwd_fixed_bytes(int64_t batch, int num_neg) {
  const size_t m = (size_t)batch * (num_neg + 1);
  return align_up(m * sizeof(int64_t)) + align_up(m * sizeof(float));
}

}  // namespace
}  // namespace etpgt

using namespace etpgt;

extern "C" size_t etpgt_sampled_loss_workspace_bytes(int64_t batch, int num_neg, int dim) {
  (void)dim;
  const size_t fwd = align_up((size_t)(batch > 0 ? batch : 1) * sizeof(float2));
  const size_t bwd = bwd_fixed_bytes(batch, num_neg) + etpgt_scatter_rows_workspace_bytes(batch * (num_neg + 1));
  return (fwd > bwd ? fwd : bwd) + 256;
}

static int check_loss_args(const char* who, int64_t batch, int num_neg, int dim, int mode, float temperature,
                           double total_sessions) {
  ETPGT_REQUIRE(supported_dim(dim), "%s: unsupported dim %d", who, dim);
  ETPGT_REQUIRE(batch >= 0 && num_neg >= 1 && num_neg <= kMaxNeg, "%s: bad batch/num_neg", who);
  ETPGT_REQUIRE(mode >= 0 && mode <= 2, "Unknown loss type: %d", mode);
  ETPGT_REQUIRE(temperature > 0.f && total_sessions >= 1.0, "%s: bad temperature / total_sessions", who);
  return ETPGT_OK;
}

extern "C" int etpgt_sampled_loss_fwd(const float* sess, const float* table, const int64_t* targets,
                                      const int64_t* negatives, int64_t batch, int num_neg, int dim, int mode,
                                      float alpha, float temperature, double total_sessions, float* scores,
                                      float* losses, void* ws, size_t ws_bytes, etpgt_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_loss_args("sampled_loss_fwd", batch, num_neg, dim, mode, temperature, total_sessions);
  if (rc != ETPGT_OK) return rc;
  if (ws_bytes < align_up((size_t)(batch > 0 ? batch : 1) * sizeof(float2))) {
    set_error("sampled_loss_fwd: workspace too small");
    return ETPGT_EWORKSPACE;
  }
  float2* terms = static_cast<float2*>(ws);
  if (batch > 0) {
#define CALL(D)                                                                                      \
  {                                                                                                  \
    const int64_t spc = (kThreads / 32) * RowGeom<D>::GROUPS;                                        \
    loss_fwd_kernel<D><<<(unsigned)((batch + spc - 1) / spc), kThreads, 0, stream>>>(                 \
        sess, table, targets, negatives, batch, num_neg, 1.f / temperature, scores, terms);          \
  }
    ETPGT_DISPATCH_DIM(dim, CALL)
#undef CALL
    ETPGT_CHECK_LAUNCH("loss_fwd");
  }
  loss_reduce_kernel<<<1, 1024, 0, stream>>>(terms, batch, num_neg, mode, alpha, total_sessions, losses);
  ETPGT_CHECK_LAUNCH("loss_reduce");
  return ETPGT_OK;
}

extern "C" int etpgt_sampled_loss_bwd_planned(const float* sess, const float* table, const int64_t* targets,
                                              const int64_t* negatives, int64_t batch, int num_neg, int dim,
                                              int mode, float alpha, float temperature, double total_sessions,
                                              const float* scores, const float* d_loss, int64_t num_items,
                                              int64_t padding_idx, const int32_t* plan_sorted_key,
                                              const int32_t* plan_perm, float* d_sess, float* d_table, void* ws,
                                              size_t ws_bytes, etpgt_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_loss_args("sampled_loss_bwd", batch, num_neg, dim, mode, temperature, total_sessions);
  if (rc != ETPGT_OK) return rc;
  if (ws_bytes < etpgt_sampled_loss_workspace_bytes(batch, num_neg, dim)) {
    set_error("sampled_loss_bwd: workspace too small");
    return ETPGT_EWORKSPACE;
  }
  if (batch == 0) return ETPGT_OK;
  Workspace w(ws, ws_bytes);
  const int64_t m = batch * (num_neg + 1);
  int64_t* keys = w.take<int64_t>(m);
  float* coef = w.take<float>(m);
  const float w_lw = mode == ETPGT_LOSS_LISTWISE ? 1.f : mode == ETPGT_LOSS_DUAL ? alpha : 0.f;
  const float w_bpr = mode == ETPGT_LOSS_BPR ? 1.f : mode == ETPGT_LOSS_DUAL ? 1.f - alpha : 0.f;
#define CALL(D)                                                                                      \
  {                                                                                                  \
    const int64_t spc = (kThreads / 32) * RowGeom<D>::GROUPS;                                        \
    loss_bwd_kernel<D><<<(unsigned)((batch + spc - 1) / spc), kThreads, 0, stream>>>(                 \
        table, targets, negatives, batch, num_neg, w_lw, w_bpr, 1.f / temperature, total_sessions, scores, d_loss, \
        d_sess, keys, coef);                                                                         \
  }
  ETPGT_DISPATCH_DIM(dim, CALL)
#undef CALL
  ETPGT_CHECK_LAUNCH("loss_bwd");
  if (d_table != nullptr) {
    // the padding row keeps a zero gradient (nn.Embedding(padding_idx=0), base.py:36)
    if (plan_sorted_key != nullptr)   // keys [b][0] = target, [b][1 + c] = negative c: sorted once per batch
      return etpgt_scatter_rows_planned(plan_sorted_key, plan_perm, coef, sess, m, num_neg + 1, dim, padding_idx,
                                        d_table, stream_);
    return etpgt_scatter_rows(keys, coef, sess, m, num_neg + 1, dim, num_items, padding_idx, d_table,
                              static_cast<char*>(ws) + w.used, ws_bytes - w.used, stream_);
  }
  return ETPGT_OK;
}

extern "C" int etpgt_sampled_loss_bwd(const float* sess, const float* table, const int64_t* targets,
                                      const int64_t* negatives, int64_t batch, int num_neg, int dim, int mode,
                                      float alpha, float temperature, double total_sessions, const float* scores,
                                      const float* d_loss, int64_t num_items, int64_t padding_idx,
                                      float* d_sess, float* d_table,
                                      void* ws, size_t ws_bytes, etpgt_stream_t stream) {
  return etpgt_sampled_loss_bwd_planned(sess, table, targets, negatives, batch, num_neg, dim, mode, alpha, temperature,
                                        total_sessions, scores, d_loss, num_items, padding_idx, nullptr, nullptr,
                                        d_sess, d_table, ws, ws_bytes, stream);
}
