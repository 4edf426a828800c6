// a11 / SURVEY.md §8(f1): the optimizer step of Trainer.train_epoch (etpgt/train/trainer.py:125-127) —
// torch.optim.AdamW(lr, weight_decay) built at scripts/train/train_baseline.py:252-256 and
// torch.optim.Adam(lr=1e-3) at scripts/pipeline/run_full_pipeline.py:210 — as ONE launch over every
// parameter of the model (the [num_items, 256] item table is 97 % of the bytes: 7 x 84 MB of traffic
// per step at 82k items, the largest byte mover of the step after the edge kernels).
//
// Dense decoupled-decay semantics of torch's single-tensor implementation, in its operation order:
//   AdamW:  p *= 1 - lr*wd                      Adam (L2):  g += wd * p
//   m  = m + (1-b1) * (g - m)                   (Tensor.lerp_)
//   v  = b2*v + (1-b2)*g*g                      (mul_ + addcmul_)
//   p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
// bc1 = 1 - b1^step, bc2 = 1 - b2^step are computed by the caller in double (they are host scalars in
// torch too).  zero_grad != 0 also clears the gradient (the next backward ACCUMULATES rows into the same
// persistent buffer, so no separate 84 MB memset / add passes are needed).
//
// HBM-bound: 4 reads + 3 (4 with zero_grad) writes of 4 bytes per element; float4 accesses, chunked
// multi-tensor launch (one CTA = one 4,096-element chunk of one tensor).
#include <stdlib.h>

#include "peer.cuh"

namespace etpgt {
namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 4096;       // elements per CTA: 4 float4 per thread in flight
constexpr int kMaxTensors = 48;    // per launch (kernel-parameter space)

struct AdamArgs {
  float* param[kMaxTensors];
  float* grad[kMaxTensors];
  float* exp_avg[kMaxTensors];
  float* exp_avg_sq[kMaxTensors];
  int64_t numel[kMaxTensors];
  int block_start[kMaxTensors + 1];  // first CTA of tensor t
  int count;
};

struct AdamScalars {
  float decay;        // AdamW: 1 - lr*wd;  Adam: wd
  float one_minus_b1, b1, b2, one_minus_b2, step_size, bc2_sqrt, eps;
  int decoupled, zero_grad, lerp_low;  // lerp_low: weight (1-b1) < 0.5, torch's first lerp formula
};

__device__ __forceinline__ void adam_one(float& p, float& g, float& m, float& v, const AdamScalars& s) {
  if (s.decoupled) p = p * s.decay;
  else g = fmaf(s.decay, p, g);
  m = s.lerp_low ? m + s.one_minus_b1 * (g - m) : g - (g - m) * s.b1;
  v = v * s.b2 + (s.one_minus_b2 * g) * g;
  const float denom = sqrtf(v) / s.bc2_sqrt + s.eps;
  p = p - (s.step_size * m) / denom;
}

__global__ void __launch_bounds__(kThreads)
adam_step_kernel(const __grid_constant__ AdamArgs a, const AdamScalars s) {
  // which tensor does this CTA belong to?  (<= 48 entries: a short scan, uniform over the CTA)
  int t = 0;
  while (t + 1 < a.count && (int)blockIdx.x >= a.block_start[t + 1]) ++t;
  const int64_t n = a.numel[t];
  const int64_t base = (int64_t)(blockIdx.x - a.block_start[t]) * kChunk;
  float* __restrict__ p = a.param[t];
  float* __restrict__ g = a.grad[t];
  float* __restrict__ m = a.exp_avg[t];
  float* __restrict__ v = a.exp_avg_sq[t];
  const bool vec = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0) && base + kChunk <= n;
  if (vec) {
    float4 P[4], G[4], M[4], V[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = base + 4 * (u * kThreads + threadIdx.x);
      P[u] = ld4(p + i);
      G[u] = ld4(g + i);
      M[u] = ld4(m + i);
      V[u] = ld4(v + i);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      adam_one(P[u].x, G[u].x, M[u].x, V[u].x, s);
      adam_one(P[u].y, G[u].y, M[u].y, V[u].y, s);
      adam_one(P[u].z, G[u].z, M[u].z, V[u].z, s);
      adam_one(P[u].w, G[u].w, M[u].w, V[u].w, s);
      const int64_t i = base + 4 * (u * kThreads + threadIdx.x);
      st4(p + i, P[u]);
      st4(m + i, M[u]);
      st4(v + i, V[u]);
      if (s.zero_grad) st4(g + i, zero4());
    }
  } else {
    const int64_t end = base + kChunk < n ? base + kChunk : n;
    for (int64_t i = base + threadIdx.x; i < end; i += kThreads) {
      float P = p[i], G = g[i], M = m[i], V = v[i];
      adam_one(P, G, M, V, s);
      p[i] = P;
      m[i] = M;
      v[i] = V;
      if (s.zero_grad) g[i] = 0.f;
    }
  }
}

// Data parallelism over peer memory: reduce-scatter + AdamW + all-gather of the item table as ONE kernel.
// Rank `c.rank` owns the float4 range [begin4, end4) of the [rows, dim] table.  Per element it
//   * pulls that element of every rank's gradient buffer over NVLink and adds them in rank order (the
//     reduce-scatter; deterministic, so any owner would compute the same sum),
//   * applies the dense AdamW / Adam update with its own moments (only the owner keeps them current),
//   * pushes the new parameter value into every rank's copy of the table (the all-gather).
// Inbound (gradients) and outbound (parameters) traffic use opposite NVLink directions and overlap.
// The caller brackets the kernel with two communicator barriers: gradients complete before, every copy of
// the table complete (and every gradient buffer free to be cleared) after.
struct DpTableArgs {
  size_t param_offset, grad_offset;   // byte offsets of the table / its gradient inside every region
  float *exp_avg, *exp_avg_sq;        // this rank's moments, [rows, dim] (only the owned rows are touched)
  int64_t begin4, end4;
};

// UN elements per thread, ranks r < WMAX = 16 / UN: all UN * world gradient loads of a thread (up to 16 x 16 B,
// most of them over NVLink, ~2 us away) are in flight before the first one is used, so ONE CTA per SM (a quarter of
// its registers, no shared memory) keeps the links busy and leaves the SM to the kernels of the other stream (the
// exchange runs underneath the last phase of the backward pass).
template <int UN, int WMAX>
__global__ void __launch_bounds__(kThreads)
dp_adam_table_kernel(const __grid_constant__ CommView c, const DpTableArgs a, const AdamScalars s) {
  float4* __restrict__ m4 = reinterpret_cast<float4*>(a.exp_avg);
  float4* __restrict__ v4 = reinterpret_cast<float4*>(a.exp_avg_sq);
  const float4* p_own = reinterpret_cast<const float4*>(c.base[c.rank] + a.param_offset);
  const int64_t tile_elems = (int64_t)kThreads * UN;
  const int64_t tiles = (a.end4 - a.begin4 + tile_elems - 1) / tile_elems;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t first = a.begin4 + tile * tile_elems + threadIdx.x;
    float4 g[UN][WMAX];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t i = first + (int64_t)u * kThreads;
#pragma unroll
      for (int r = 0; r < WMAX; ++r)
        if (r < c.world && i < a.end4) g[u][r] = __ldcg(reinterpret_cast<const float4*>(c.base[r] + a.grad_offset) + i);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t i = first + (int64_t)u * kThreads;
      if (i >= a.end4) break;
      float4 P = p_own[i], M = m4[i], V = v4[i];
      float4 G = g[u][0];
#pragma unroll
      for (int r = 1; r < WMAX; ++r)
        if (r < c.world) G = add4(G, g[u][r]);      // rank order: every owner would form the same sum
      adam_one(P.x, G.x, M.x, V.x, s);
      adam_one(P.y, G.y, M.y, V.y, s);
      adam_one(P.z, G.z, M.z, V.z, s);
      adam_one(P.w, G.w, M.w, V.w, s);
#pragma unroll
      for (int r = 0; r < WMAX; ++r)
        if (r < c.world) reinterpret_cast<float4*>(c.base[r] + a.param_offset)[i] = P;
      m4[i] = M;
      v4[i] = V;
    }
  }
  __threadfence_system();   // the pushed rows are visible to the peers before this rank enters the barrier
}

bool fill_scalars(AdamScalars& s, double lr, double beta1, double beta2, double eps, double weight_decay,
                  int decoupled, int64_t step, int zero_grad) {
  // hyper-parameters arrive as doubles (Python floats in torch) and are rounded to fp32 only where
  // torch rounds them: 1 - beta is formed in double first
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  s.decay = (float)(decoupled ? 1.0 - lr * weight_decay : weight_decay);
  s.one_minus_b1 = (float)(1.0 - beta1);
  s.b2 = (float)beta2;
  s.one_minus_b2 = (float)(1.0 - beta2);
  s.step_size = (float)(lr / bc1);
  s.bc2_sqrt = (float)sqrt(bc2);
  s.b1 = (float)beta1;
  s.lerp_low = (1.0 - beta1) < 0.5;
  s.eps = (float)eps;
  s.decoupled = decoupled != 0;
  s.zero_grad = zero_grad != 0;
  return true;
}

}  // namespace
}  // namespace etpgt

using namespace etpgt;

extern "C" int etpgt_adam_step(const etpgt_adam_tensor* tensors, int count, double lr, double beta1, double beta2,
                               double eps, double weight_decay, int decoupled, int64_t step, int zero_grad,
                               etpgt_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ETPGT_REQUIRE(count >= 0 && (count == 0 || tensors != nullptr), "adam_step: bad tensor list");
  ETPGT_REQUIRE(step >= 1, "adam_step: step must be >= 1 (1-based, as torch counts it)");
  ETPGT_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0 && lr >= 0.0 &&
                    weight_decay >= 0.0,
                "adam_step: hyper-parameters out of range");
  for (int i = 0; i < count; ++i)
    ETPGT_REQUIRE(tensors[i].numel >= 0 && (tensors[i].numel == 0 || (tensors[i].param && tensors[i].grad &&
                                                                        tensors[i].exp_avg && tensors[i].exp_avg_sq)),
                  "adam_step: tensor %d has a NULL pointer", i);
  for (int i = 0; i < count; ++i)
    ETPGT_REQUIRE(tensors[i].numel < (int64_t(1) << 40), "adam_step: tensor %d too large", i);
  AdamScalars s;
  fill_scalars(s, lr, beta1, beta2, eps, weight_decay, decoupled, step, zero_grad);
  int done = 0;
  while (done < count) {
    AdamArgs a;
    a.count = 0;
    int blocks = 0;
    while (done < count && a.count < kMaxTensors) {
      const etpgt_adam_tensor& t = tensors[done++];
      if (t.numel == 0) continue;
      const int64_t nb = (t.numel + kChunk - 1) / kChunk;
      if (blocks + nb >= (int64_t(1) << 30)) { --done; break; }  // next launch
      a.param[a.count] = t.param;
      a.grad[a.count] = t.grad;
      a.exp_avg[a.count] = t.exp_avg;
      a.exp_avg_sq[a.count] = t.exp_avg_sq;
      a.numel[a.count] = t.numel;
      a.block_start[a.count] = blocks;
      blocks += (int)nb;
      ++a.count;
    }
    if (a.count == 0) break;
    a.block_start[a.count] = blocks;
    adam_step_kernel<<<blocks, kThreads, 0, stream>>>(a, s);
    ETPGT_CHECK_LAUNCH("adam_step");
  }
  return ETPGT_OK;
}

extern "C" int etpgt_dp_adam_table(const etpgt_comm_t* comm, size_t param_offset, size_t grad_offset, float* exp_avg,
                                   float* exp_avg_sq, int64_t rows, int dim, int64_t row_begin, int64_t row_end,
                                   double lr, double beta1, double beta2, double eps, double weight_decay,
                                   int decoupled, int64_t step, etpgt_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ETPGT_REQUIRE(comm != nullptr, "dp_adam_table: null communicator");
  const CommView c = comm_view(comm);
  for (int r = 0; r < c.world; ++r) ETPGT_REQUIRE(c.base[r] != nullptr, "dp_adam_table: communicator not connected");
  ETPGT_REQUIRE(rows >= 0 && dim >= 4 && dim % 4 == 0 && row_begin >= 0 && row_begin <= row_end && row_end <= rows,
                "dp_adam_table: bad shape (rows %lld, dim %d, shard [%lld, %lld))", (long long)rows, dim,
                (long long)row_begin, (long long)row_end);
  ETPGT_REQUIRE(param_offset % 16 == 0 && grad_offset % 16 == 0 && exp_avg && exp_avg_sq &&
                    (((uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
                "dp_adam_table: buffers must be 16-byte aligned");
  ETPGT_REQUIRE(step >= 1 && beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0 && lr >= 0.0 &&
                    weight_decay >= 0.0,
                "dp_adam_table: hyper-parameters out of range");
  if (row_begin == row_end) return ETPGT_OK;
  AdamScalars s;
  fill_scalars(s, lr, beta1, beta2, eps, weight_decay, decoupled, step, 0);
  DpTableArgs a;
  a.param_offset = param_offset;
  a.grad_offset = grad_offset;
  a.exp_avg = exp_avg;
  a.exp_avg_sq = exp_avg_sq;
  a.begin4 = row_begin * (dim / 4);
  a.end4 = row_end * (dim / 4);
  // one CTA per SM: see the kernel's comment (ETPGT_DP_TABLE_CTAS: tuning knob for the grid)
  int cap = kNumSMs;
  if (const char* forced = getenv("ETPGT_DP_TABLE_CTAS")) {
    const int f = atoi(forced);
    if (f >= 1 && f <= 8 * kNumSMs) cap = f;
  }
  auto grid_of = [&](int un) {
    const int64_t tiles = (a.end4 - a.begin4 + (int64_t)kThreads * un - 1) / ((int64_t)kThreads * un);
    return (int)(tiles < cap ? tiles : cap);
  };
  if (c.world <= 2) {
    dp_adam_table_kernel<8, 2><<<grid_of(8), kThreads, 0, stream>>>(c, a, s);
  } else if (c.world <= 4) {
    dp_adam_table_kernel<4, 4><<<grid_of(4), kThreads, 0, stream>>>(c, a, s);
  } else {
    dp_adam_table_kernel<2, 8><<<grid_of(2), kThreads, 0, stream>>>(c, a, s);
  }
  ETPGT_CHECK_LAUNCH("dp_adam_table");
  return ETPGT_OK;
}
