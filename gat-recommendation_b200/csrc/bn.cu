// a6: BatchNorm1d over the node rows of the whole batch, fused with the residual add (and the
// ReLU of the GAT / GraphSAGE wrappers).  etpgt/model/graph_transformer.py:175-176,
// gat.py:138-140, graphsage.py:76-77.
//
// The batch statistics couple every node of the (global) batch, so the op is split at the
// reduction: stats (per-CTA column sums in double, fixed-order second stage) -> [caller may
// all-reduce 2*dim doubles across ranks] -> finalize -> apply.  Streaming, HBM-bound:
// stats reads N*dim*4 B; apply reads 2 rows and writes 1 per node.
#include <cuda_bf16.h>

#include "common.cuh"

namespace etpgt {
namespace {

constexpr int kRowsPerCta = 64;   // small chunks: enough CTAs (<= 8 per SM) to cover the HBM latency
constexpr int kRowUnroll = 4;     // independent row loads in flight per thread

int stat_parts(int64_t n) {
  int64_t parts = (n + kRowsPerCta - 1) / kRowsPerCta;
  if (parts > 8 * kNumSMs) parts = 8 * kNumSMs;
  return parts < 1 ? 1 : (int)parts;
}

// blockDim = (dim/4 lanes-x, rows-y): thread (tx, ty) strides over rows ty, ty+RY, ... of the
// CTA's chunk and owns columns 4*tx..4*tx+3.  MODE 0: sum x, sum x^2.  MODE 1 (backward):
// sum g, sum g*xhat with g = d_y * (relu ? y > 0 : 1).
// RELU is a template parameter so that the plain-BatchNorm instance carries no registers for the y rows
// (three resident CTAs per SM instead of two).
template <int MODE, bool RELU>
__global__ void __launch_bounds__(256, RELU ? 2 : 3)
bn_partial_kernel(const float* __restrict__ x, const float* __restrict__ y,
                  const float* __restrict__ d_y, int64_t n, int dim,
                  const float* __restrict__ mean, const float* __restrict__ invstd,
                  uint32_t drop_threshold, float keep_scale, uint64_t drop_seed,
                  int64_t chunk, double* __restrict__ partial /* [grid][2][dim] */) {
  extern __shared__ double sm[];  // [blockDim.y][2][dim]
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t begin = blockIdx.x * chunk;
  const int64_t end = begin + chunk < n ? begin + chunk : n;
  double s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
  float4 mu = zero4(), is = zero4();
  if (MODE == 1) { mu = ld4(mean + 4 * tx); is = ld4(invstd + 4 * tx); }
  for (int64_t r0 = begin + ty; r0 < end; r0 += (int64_t)kRowUnroll * blockDim.y) {
    float4 a[kRowUnroll], g[kRowUnroll], o[RELU ? kRowUnroll : 1];
    bool on[kRowUnroll];
#pragma unroll
    for (int u = 0; u < kRowUnroll; ++u) {   // all loads of the unrolled rows are issued before any use
      const int64_t r = r0 + (int64_t)u * blockDim.y;
      on[u] = r < end;
      const int64_t rr = on[u] ? r : begin;
      a[u] = ldg4(x + rr * dim + 4 * tx);
      if (MODE == 1) {
        g[u] = ldg4(d_y + rr * dim + 4 * tx);
        if (RELU) o[u] = ldg4(y + rr * dim + 4 * tx);
      }
    }
#pragma unroll
    for (int u = 0; u < kRowUnroll; ++u) {
      if (!on[u]) continue;
      const float4 av = a[u];
      if (MODE == 0) {
        s0[0] += av.x; s0[1] += av.y; s0[2] += av.z; s0[3] += av.w;
        s1[0] += (double)av.x * av.x; s1[1] += (double)av.y * av.y; s1[2] += (double)av.z * av.z;
        s1[3] += (double)av.w * av.w;
      } else {
        float4 gv = g[u];
        if (drop_threshold != 0) {   // the layer output was dropout(...): its gradient passes the same mask
          const int64_t r = r0 + (int64_t)u * blockDim.y;
          gv = mul4(gv, dropout_factors4(drop_seed, (uint64_t)(r * (dim / 4) + tx), drop_threshold, keep_scale));
        }
        if (RELU) {
          gv.x = o[u].x > 0.f ? gv.x : 0.f; gv.y = o[u].y > 0.f ? gv.y : 0.f;
          gv.z = o[u].z > 0.f ? gv.z : 0.f; gv.w = o[u].w > 0.f ? gv.w : 0.f;
        }
        s0[0] += gv.x; s0[1] += gv.y; s0[2] += gv.z; s0[3] += gv.w;
        s1[0] += (double)gv.x * ((av.x - mu.x) * is.x); s1[1] += (double)gv.y * ((av.y - mu.y) * is.y);
        s1[2] += (double)gv.z * ((av.z - mu.z) * is.z); s1[3] += (double)gv.w * ((av.w - mu.w) * is.w);
      }
    }
  }
  double* mine = sm + (size_t)ty * 2 * dim;
#pragma unroll
  for (int c = 0; c < 4; ++c) { mine[4 * tx + c] = s0[c]; mine[dim + 4 * tx + c] = s1[c]; }
  __syncthreads();
  const int tid = ty * blockDim.x + tx;
  for (int i = tid; i < 2 * dim; i += blockDim.x * blockDim.y) {
    double s = 0;
    for (int yy = 0; yy < (int)blockDim.y; ++yy) s += sm[(size_t)yy * 2 * dim + i];
    partial[(int64_t)blockIdx.x * 2 * dim + i] = s;
  }
}

// one warp per output column: lanes stride over the per-CTA partials, fixed butterfly -> deterministic
__global__ void bn_reduce_kernel(const double* __restrict__ partial, int parts, int width, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= width) return;
  double s = 0;
  for (int p = lane; p < parts; p += 32) s += partial[(int64_t)p * width + i];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) out[i] = s;
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, int dim, float eps,
                                   float momentum, float* __restrict__ mean, float* __restrict__ invstd,
                                   float* __restrict__ running_mean, float* __restrict__ running_var) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= dim) return;
  if (count <= 0) count = sums[2 * dim];  // global row count carried (and all-reduced) behind the sums
  const double mu = sums[d] / count;
  double var = sums[dim + d] / count - mu * mu;
  if (var < 0) var = 0;
  mean[d] = (float)mu;
  invstd[d] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean != nullptr) {
    const double unbiased = count > 1 ? var * count / (count - 1) : var;
    running_mean[d] = (float)((1.0 - momentum) * running_mean[d] + momentum * mu);
    running_var[d] = (float)((1.0 - momentum) * running_var[d] + momentum * unbiased);
  }
}

__global__ void bn_from_running_kernel(const float* __restrict__ rm, const float* __restrict__ rv, int dim,
                                       float eps, float* __restrict__ mean, float* __restrict__ invstd) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= dim) return;
  mean[d] = rm[d];
  invstd[d] = 1.f / sqrtf(rv[d] + eps);
}

__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ x, int64_t total4, int dim4, const float* __restrict__ mean,
                const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ bias,
                const float* __restrict__ residual, int relu, uint32_t drop_threshold, float keep_scale,
                uint64_t drop_seed, float* __restrict__ y, __nv_bfloat16* __restrict__ y_hi,
                __nv_bfloat16* __restrict__ y_lo) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % dim4) * 4;
    const float4 a = ldg4(x + 4 * i);
    const float4 mu = ld4(mean + c), is = ld4(invstd + c), ga = ld4(gamma + c), be = ld4(bias + c);
    float4 o;
    o.x = (a.x - mu.x) * is.x * ga.x + be.x;
    o.y = (a.y - mu.y) * is.y * ga.y + be.y;
    o.z = (a.z - mu.z) * is.z * ga.z + be.z;
    o.w = (a.w - mu.w) * is.w * ga.w + be.w;
    if (residual != nullptr) o = add4(o, ldg4(residual + 4 * i));
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    if (drop_threshold != 0) o = mul4(o, dropout_factors4(drop_seed, (uint64_t)i, drop_threshold, keep_scale));
    st4(y + 4 * i, o);
    if (y_hi != nullptr) {   // the next layer's projection operand, split for the tensor cores (x = hi + lo)
      const __nv_bfloat162 h0 = __floats2bfloat162_rn(o.x, o.y), h1 = __floats2bfloat162_rn(o.z, o.w);
      const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
      const __nv_bfloat162 l0 = __floats2bfloat162_rn(o.x - f0.x, o.y - f0.y), l1 = __floats2bfloat162_rn(o.z - f1.x, o.w - f1.y);
      uint2 ph, pl;
      ph.x = *reinterpret_cast<const uint32_t*>(&h0); ph.y = *reinterpret_cast<const uint32_t*>(&h1);
      pl.x = *reinterpret_cast<const uint32_t*>(&l0); pl.y = *reinterpret_cast<const uint32_t*>(&l1);
      *reinterpret_cast<uint2*>(y_hi + 4 * i) = ph;
      *reinterpret_cast<uint2*>(y_lo + 4 * i) = pl;
    }
  }
}

// d_x = gamma*invstd*(g - sum_g/count - xhat*sum_gxhat/count)   (training)
//     = gamma*invstd*g                                         (eval)
// FIXED_COLS: the grid stride is a multiple of the row length, so a thread keeps the same four columns for its
// whole loop and the per-column coefficients (incl. the double -> float conversions of the sums) are formed once,
// in registers; UNROLL independent elements are loaded before any is used.
template <bool FIXED_COLS, bool RELU, int UNROLL>
__global__ void __launch_bounds__(256, 3)
bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ d_y,
                    int64_t total4, int dim, const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ gamma, int training, const double* __restrict__ sums,
                    double count, uint32_t drop_threshold, float keep_scale, uint64_t drop_seed,
                    float* __restrict__ d_x, float* __restrict__ d_res) {
  const int dim4 = dim / 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t first = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  float4 is, gi, mu, gs, gx;   // invstd, gamma * invstd, mean, sum_g / count, sum_gxhat / count of this thread's columns
  auto coefficients = [&](int c) {
    is = ld4(invstd + c);
    const float4 ga = ld4(gamma + c);
    gi = make_float4(ga.x * is.x, ga.y * is.y, ga.z * is.z, ga.w * is.w);
    mu = gs = gx = zero4();
    if (training) {
      mu = ld4(mean + c);
      const float inv_n = (float)(1.0 / (count > 0 ? count : sums[2 * dim]));
      gs = make_float4((float)sums[c] * inv_n, (float)sums[c + 1] * inv_n, (float)sums[c + 2] * inv_n,
                       (float)sums[c + 3] * inv_n);
      gx = make_float4((float)sums[dim + c] * inv_n, (float)sums[dim + c + 1] * inv_n,
                       (float)sums[dim + c + 2] * inv_n, (float)sums[dim + c + 3] * inv_n);
    }
  };
  if (FIXED_COLS) coefficients((int)(first % dim4) * 4);
  for (int64_t i0 = first; i0 < total4; i0 += UNROLL * stride) {
    float4 g[UNROLL], a[UNROLL], o[RELU ? UNROLL : 1];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < total4) {
        g[u] = ldg4(d_y + 4 * i);
        if (training) a[u] = ldg4(x + 4 * i);
        if (RELU) o[u] = ldg4(y + 4 * i);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= total4) break;
      if (!FIXED_COLS) coefficients((int)(i % dim4) * 4);
      float4 gv = g[u];
      if (drop_threshold != 0) gv = mul4(gv, dropout_factors4(drop_seed, (uint64_t)i, drop_threshold, keep_scale));
      if (RELU) {
        gv.x = o[u].x > 0.f ? gv.x : 0.f; gv.y = o[u].y > 0.f ? gv.y : 0.f;
        gv.z = o[u].z > 0.f ? gv.z : 0.f; gv.w = o[u].w > 0.f ? gv.w : 0.f;
      }
      if (d_res != nullptr) st4(d_res + 4 * i, gv);   // gradient of the residual branch (same mask, same ReLU gate)
      float4 r;
      if (training) {
        const float4 av = a[u];
        r.x = gi.x * (gv.x - gs.x - (av.x - mu.x) * is.x * gx.x);
        r.y = gi.y * (gv.y - gs.y - (av.y - mu.y) * is.y * gx.y);
        r.z = gi.z * (gv.z - gs.z - (av.z - mu.z) * is.z * gx.z);
        r.w = gi.w * (gv.w - gs.w - (av.w - mu.w) * is.w * gx.w);
      } else {
        r.x = gi.x * gv.x; r.y = gi.y * gv.y; r.z = gi.z * gv.z; r.w = gi.w * gv.w;
      }
      st4(d_x + 4 * i, r);
    }
  }
}

__global__ void bn_param_grad_kernel(const double* __restrict__ local_sums, int dim, float* __restrict__ d_gamma,
                                     float* __restrict__ d_bias) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= dim) return;
  d_bias[d] = (float)local_sums[d];
  d_gamma[d] = (float)local_sums[dim + d];
}

// out[i] = 0 with probability p, else 1/(1-p): the inverted-dropout mask over attention coefficients
// (PyG TransformerConv / GATConv `F.dropout(alpha, p)`), for every layer of a forward pass at once.
__global__ void __launch_bounds__(256)
dropout_mask_kernel(uint64_t seed, int64_t n4, int64_t n, uint32_t threshold, float keep_scale, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 f = dropout_factors4(seed, (uint64_t)i, threshold, keep_scale);
    if (4 * i + 3 < n) st4(out + 4 * i, f);
    else {
      const float e[4] = {f.x, f.y, f.z, f.w};
      for (int c = 0; c < 4 && 4 * i + c < n; ++c) out[4 * i + c] = e[c];
    }
  }
}

template <int MODE>
int launch_partial(const float* x, const float* y, const float* d_y, int64_t n, int dim, const float* mean,
                   const float* invstd, int relu, uint32_t drop_threshold, float keep_scale, uint64_t drop_seed,
                   double* sums, void* ws, size_t ws_bytes, cudaStream_t stream) {
  const int parts = stat_parts(n);
  const size_t need = align_up((size_t)parts * 2 * dim * sizeof(double));
  if (ws_bytes < need) { set_error("bn: workspace %zu < %zu", ws_bytes, need); return ETPGT_EWORKSPACE; }
  double* partial = static_cast<double*>(ws);
  const int64_t chunk = n > 0 ? (n + parts - 1) / parts : 0;
  const int tx = dim / 4;
  const int ty = 256 / tx > 0 ? 256 / tx : 1;
  const size_t smem = (size_t)ty * 2 * dim * sizeof(double);
  auto kern = (MODE == 1 && relu) ? bn_partial_kernel<MODE, true> : bn_partial_kernel<MODE, false>;
  kern<<<parts, dim3(tx, ty), smem, stream>>>(x, y, d_y, n, dim, mean, invstd, drop_threshold, keep_scale, drop_seed,
                                              chunk, partial);
  ETPGT_CHECK_LAUNCH("bn_partial");
  bn_reduce_kernel<<<(2 * dim * 32 + 255) / 256, 256, 0, stream>>>(partial, parts, 2 * dim, sums);
  ETPGT_CHECK_LAUNCH("bn_reduce");
  return ETPGT_OK;
}

}  // namespace
}  // namespace etpgt

using namespace etpgt;

extern "C" size_t etpgt_bn_workspace_bytes(int64_t n, int dim) {
  return align_up((size_t)stat_parts(n) * 2 * dim * sizeof(double)) + 256;
}

extern "C" int etpgt_bn_stats(const float* x, int64_t n, int dim, double* sums, void* ws, size_t ws_bytes,
                              etpgt_stream_t stream) {
  ETPGT_REQUIRE(dim % 4 == 0 && dim >= 4 && dim <= 1024, "bn_stats: dim %d must be a multiple of 4 <= 1024", dim);
  ETPGT_REQUIRE(n >= 0 && x && sums, "bn_stats: bad arguments");
  return launch_partial<0>(x, nullptr, nullptr, n, dim, nullptr, nullptr, 0, 0u, 1.f, 0ull, sums, ws, ws_bytes,
                           static_cast<cudaStream_t>(stream));
}

extern "C" int etpgt_bn_finalize(const double* sums, double count, int dim, float eps, float momentum, float* mean,
                                 float* invstd, float* running_mean, float* running_var, etpgt_stream_t stream) {
  ETPGT_REQUIRE((count >= 1 || count == 0) && dim > 0 && sums && mean && invstd, "bn_finalize: bad arguments");
  bn_finalize_kernel<<<(dim + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      sums, count, dim, eps, momentum, mean, invstd, running_mean, running_var);
  ETPGT_CHECK_LAUNCH("bn_finalize");
  return ETPGT_OK;
}

extern "C" int etpgt_bn_from_running(const float* running_mean, const float* running_var, int dim, float eps,
                                     float* mean, float* invstd, etpgt_stream_t stream) {
  ETPGT_REQUIRE(dim > 0 && running_mean && running_var && mean && invstd, "bn_from_running: bad arguments");
  bn_from_running_kernel<<<(dim + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(running_mean, running_var,
                                                                                        dim, eps, mean, invstd);
  ETPGT_CHECK_LAUNCH("bn_from_running");
  return ETPGT_OK;
}

// p -> (threshold = p * 2^32, scale = 1 / (1 - p)); p == 0 disables the mask
static bool dropout_params(double p, uint32_t* threshold, float* keep_scale) {
  if (!(p >= 0.0 && p < 1.0)) return false;
  const double t = p * 4294967296.0;
  *threshold = p > 0.0 ? (uint32_t)(t < 1.0 ? 1.0 : (t > 4294967295.0 ? 4294967295.0 : t)) : 0u;
  *keep_scale = (float)(1.0 / (1.0 - p));
  return true;
}

extern "C" int etpgt_dropout_mask(uint64_t seed, double p, int64_t n, float* out, etpgt_stream_t stream) {
  uint32_t threshold;
  float keep_scale;
  ETPGT_REQUIRE(dropout_params(p, &threshold, &keep_scale), "dropout_mask: p must be in [0, 1)");
  ETPGT_REQUIRE(n >= 0 && (n == 0 || out != nullptr) && ((uintptr_t)out & 15) == 0, "dropout_mask: bad arguments");
  if (n == 0) return ETPGT_OK;
  const int64_t n4 = (n + 3) / 4;
  dropout_mask_kernel<<<grid_for(n4, 256 * 4, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(seed, n4, n, threshold,
                                                                                           keep_scale, out);
  ETPGT_CHECK_LAUNCH("dropout_mask");
  return ETPGT_OK;
}

extern "C" int etpgt_bn_apply_ex(const float* x, int64_t n, int dim, const float* mean, const float* invstd,
                                 const float* gamma, const float* bias, const float* residual, int relu,
                                 double drop_p, uint64_t drop_seed, float* y, void* y_hi, void* y_lo,
                                 etpgt_stream_t stream) {
  ETPGT_REQUIRE(dim % 4 == 0 && dim >= 4, "bn_apply: dim %d must be a multiple of 4", dim);
  ETPGT_REQUIRE(n >= 0 && x && mean && invstd && gamma && bias && y, "bn_apply: bad arguments");
  ETPGT_REQUIRE((y_hi == nullptr) == (y_lo == nullptr), "bn_apply: split outputs come in pairs");
  uint32_t threshold;
  float keep_scale;
  ETPGT_REQUIRE(dropout_params(drop_p, &threshold, &keep_scale), "bn_apply: dropout p must be in [0, 1)");
  if (n == 0) return ETPGT_OK;
  const int64_t total4 = n * (dim / 4);
  bn_apply_kernel<<<grid_for(total4, 256 * 4, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, total4, dim / 4, mean, invstd, gamma, bias, residual, relu, threshold, keep_scale, drop_seed, y,
      static_cast<__nv_bfloat16*>(y_hi), static_cast<__nv_bfloat16*>(y_lo));
  ETPGT_CHECK_LAUNCH("bn_apply");
  return ETPGT_OK;
}

extern "C" int etpgt_bn_apply(const float* x, int64_t n, int dim, const float* mean, const float* invstd,
                              const float* gamma, const float* bias, const float* residual, int relu, float* y,
                              etpgt_stream_t stream) {
  return etpgt_bn_apply_ex(x, n, dim, mean, invstd, gamma, bias, residual, relu, 0.0, 0, y, nullptr, nullptr, stream);
}

extern "C" int etpgt_bn_bwd_stats_ex(const float* x, const float* y, const float* d_y, int64_t n, int dim,
                                     const float* mean, const float* invstd, int relu, double drop_p,
                                     uint64_t drop_seed, double* sums, void* ws, size_t ws_bytes,
                                     etpgt_stream_t stream) {
  ETPGT_REQUIRE(dim % 4 == 0 && dim >= 4 && dim <= 1024, "bn_bwd_stats: dim %d must be a multiple of 4 <= 1024", dim);
  ETPGT_REQUIRE(n >= 0 && x && d_y && mean && invstd && sums && (!relu || y), "bn_bwd_stats: bad arguments");
  uint32_t threshold;
  float keep_scale;
  ETPGT_REQUIRE(dropout_params(drop_p, &threshold, &keep_scale), "bn_bwd_stats: dropout p must be in [0, 1)");
  return launch_partial<1>(x, y, d_y, n, dim, mean, invstd, relu, threshold, keep_scale, drop_seed, sums, ws, ws_bytes,
                           static_cast<cudaStream_t>(stream));
}

extern "C" int etpgt_bn_bwd_stats(const float* x, const float* y, const float* d_y, int64_t n, int dim,
                                  const float* mean, const float* invstd, int relu, double* sums, void* ws,
                                  size_t ws_bytes, etpgt_stream_t stream) {
  return etpgt_bn_bwd_stats_ex(x, y, d_y, n, dim, mean, invstd, relu, 0.0, 0, sums, ws, ws_bytes, stream);
}

extern "C" int etpgt_bn_bwd_apply_ex(const float* x, const float* y, const float* d_y, int64_t n, int dim,
                                     const float* mean, const float* invstd, const float* gamma, int relu,
                                     int training, const double* sums, double count, const double* local_sums,
                                     double drop_p, uint64_t drop_seed, float* d_x, float* d_res, float* d_gamma,
                                     float* d_bias, etpgt_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ETPGT_REQUIRE(dim % 4 == 0 && dim >= 4, "bn_bwd_apply: dim %d must be a multiple of 4", dim);
  ETPGT_REQUIRE(n >= 0 && d_y && mean && invstd && gamma && d_x && local_sums && (!training || (x && sums && (count >= 1 || count == 0))),
                "bn_bwd_apply: bad arguments");
  uint32_t threshold;
  float keep_scale;
  ETPGT_REQUIRE(dropout_params(drop_p, &threshold, &keep_scale), "bn_bwd_apply: dropout p must be in [0, 1)");
  if (n > 0) {
    const int64_t total4 = n * (dim / 4);
    const int grid = grid_for(total4, 256 * 4, 6);
    const bool fixed = ((int64_t)grid * 256) % (dim / 4) == 0;
    // (two elements, i.e. four to six 16-byte loads, in flight per thread)
    auto kern = relu ? (fixed ? bn_bwd_apply_kernel<true, true, 2> : bn_bwd_apply_kernel<false, true, 2>)
                     : (fixed ? bn_bwd_apply_kernel<true, false, 2> : bn_bwd_apply_kernel<false, false, 2>);
    kern<<<grid, 256, 0, stream>>>(x, y, d_y, total4, dim, mean, invstd, gamma, training, sums, count, threshold,
                                   keep_scale, drop_seed, d_x, d_res);
    ETPGT_CHECK_LAUNCH("bn_bwd_apply");
  }
  if (d_gamma != nullptr && d_bias != nullptr) {
    bn_param_grad_kernel<<<(dim + 127) / 128, 128, 0, stream>>>(local_sums, dim, d_gamma, d_bias);
    ETPGT_CHECK_LAUNCH("bn_param_grad");
  }
  return ETPGT_OK;
}

extern "C" int etpgt_bn_bwd_apply(const float* x, const float* y, const float* d_y, int64_t n, int dim,
                                  const float* mean, const float* invstd, const float* gamma, int relu,
                                  int training, const double* sums, double count, const double* local_sums,
                                  float* d_x, float* d_gamma, float* d_bias, etpgt_stream_t stream) {
  return etpgt_bn_bwd_apply_ex(x, y, d_y, n, dim, mean, invstd, gamma, relu, training, sums, count, local_sums, 0.0, 0,
                               d_x, nullptr, d_gamma, d_bias, stream);
}
