// a12, the node-wise parts of PyG GATConv around the edge kernels of gat.cu (etpgt/model/gat.py:49-109,137):
//   a_src[n,h] = <h[n,h,:], att_src[h,:]>,  a_dst[n,h] = <h[n,h,:], att_dst[h,:]>         (attention scalars)
//   out[n,:]   = mean_h agg[n,h,:] + bias                                                   (concat=False)
// and their backward passes.  In PyTorch these are a dozen element-wise / reduction launches over
// [N, heads*C] tensors (two thirds of a GAT step on B200); here each is ONE streaming pass.
// HBM-bound: scores fwd reads N*W*4 B; scores bwd reads and rewrites d_h (2*N*W*4 B) and reads h;
// head mean fwd reads N*W*4, writes N*C*4; bwd the reverse.  Parameter gradients (d_att_src, d_att_dst,
// d_bias) are column sums: per-CTA partials + a fixed-order coalesced reduce (deterministic).
#include "common.cuh"

namespace etpgt {
namespace {

constexpr int kThreads = 256;

// One warp per node, lanes stride over the float4 columns of the row.  heads*C = W, C % 4 == 0.
__global__ void __launch_bounds__(kThreads)
gat_scores_fwd_kernel(const float* __restrict__ h, const float* __restrict__ att_src, const float* __restrict__ att_dst,
                      int64_t n, int width, int channels, float* __restrict__ a_src, float* __restrict__ a_dst) {
  const int lane = threadIdx.x & 31;
  const int heads = width / channels;
  const int f4_per_head = channels / 4;
  for (int64_t node = (blockIdx.x * (int64_t)kThreads + threadIdx.x) >> 5; node < n;
       node += ((int64_t)gridDim.x * kThreads) >> 5) {
    const float* row = h + node * width;
    for (int hd = 0; hd < heads; ++hd) {
      float ps = 0.f, pd = 0.f;
      for (int f = lane; f < f4_per_head; f += 32) {
        const int c = hd * channels + 4 * f;
        const float4 x = ldg4(row + c);
        ps += dot4(x, ldg4(att_src + c));
        pd += dot4(x, ldg4(att_dst + c));
      }
      ps = group_sum<32>(ps);
      pd = group_sum<32>(pd);
      if (lane == 0) {
        a_src[node * heads + hd] = ps;
        a_dst[node * heads + hd] = pd;
      }
    }
  }
}

// d_h[n,h,c] += d_a_src[n,h]*att_src[h,c] + d_a_dst[n,h]*att_dst[h,c];
// partial[cta][0:W] = sum_n d_a_src[n,h]*h[n,h,c], partial[cta][W:2W] = the same with d_a_dst.
// blockDim = (W/4, rows): thread (tx, ty) owns float4 column tx for rows ty, ty+RY, ... of the CTA's chunk.
__global__ void gat_scores_bwd_kernel(const float* __restrict__ h, const float* __restrict__ att_src,
                                      const float* __restrict__ att_dst, const float* __restrict__ d_a_src,
                                      const float* __restrict__ d_a_dst, int64_t n, int width, int channels,
                                      int64_t chunk, float* __restrict__ d_h, float* __restrict__ partial) {
  extern __shared__ float sm[];  // [blockDim.y][2][W]
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int heads = width / channels;
  const int hd = (4 * tx) / channels;
  const float4 ws = ldg4(att_src + 4 * tx), wd = ldg4(att_dst + 4 * tx);
  float4 gs = zero4(), gd = zero4();
  const int64_t begin = blockIdx.x * chunk;
  const int64_t end = begin + chunk < n ? begin + chunk : n;
  for (int64_t r = begin + ty; r < end; r += blockDim.y) {
    const float ds = d_a_src[r * heads + hd], dd = d_a_dst[r * heads + hd];
    const float4 x = ldg4(h + r * width + 4 * tx);
    float* out = d_h + r * width + 4 * tx;
    st4(out, fma4(ds, ws, fma4(dd, wd, ld4(out))));
    gs = fma4(ds, x, gs);
    gd = fma4(dd, x, gd);
  }
  float* mine = sm + (size_t)ty * 2 * width;
  st4(mine + 4 * tx, gs);
  st4(mine + width + 4 * tx, gd);
  __syncthreads();
  const int tid = ty * blockDim.x + tx;
  for (int i = tid; i < 2 * width; i += blockDim.x * blockDim.y) {
    float s = 0.f;
    for (int yy = 0; yy < (int)blockDim.y; ++yy) s += sm[(size_t)yy * 2 * width + i];
    partial[(int64_t)blockIdx.x * 2 * width + i] = s;
  }
}

// Attention scalars straight from the layer INPUT: a_src[n,h] = <lin(x)[n,h,:], att_src[h,:]> = <x[n,:], u[h,:]> with
// u[h,:] = W_h^T att_src[h,:] (a [heads, in] fold of two parameters, made by the caller).  Reads the [N, in] input
// instead of the [N, heads*C] projection (a quarter of the bytes at heads = 4), and — more importantly — its
// backward no longer touches d(lin(x)): d_x += d_a u, d_u = d_a^T x.  One warp per node; u[r,:] for r < R = 2*heads
// rows (sources then destinations) comes from L1.
template <int R>
__global__ void __launch_bounds__(kThreads)
rows_dot_fwd_kernel(const float* __restrict__ x, const float* __restrict__ u, int64_t n, int dim,
                    float* __restrict__ a_src, float* __restrict__ a_dst) {
  const int lane = threadIdx.x & 31;
  const int f4 = dim / 4;
  for (int64_t node = (blockIdx.x * (int64_t)kThreads + threadIdx.x) >> 5; node < n;
       node += ((int64_t)gridDim.x * kThreads) >> 5) {
    float p[R];
#pragma unroll
    for (int r = 0; r < R; ++r) p[r] = 0.f;
    for (int f = lane; f < f4; f += 32) {
      const float4 xv = ldg4(x + node * dim + 4 * f);
#pragma unroll
      for (int r = 0; r < R; ++r) p[r] += dot4(xv, ldg4(u + (int64_t)r * dim + 4 * f));
    }
#pragma unroll
    for (int r = 0; r < R; ++r) p[r] = group_sum<32>(p[r]);
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < R / 2; ++r) {
        a_src[node * (R / 2) + r] = p[r];
        a_dst[node * (R / 2) + r] = p[R / 2 + r];
      }
    }
  }
}

// d_x[n,:] = sum_r d_a[n,r] u[r,:]  (written, the caller's GEMM adds dY W onto it);
// partial[cta][r][:] = sum over the CTA's rows of d_a[n,r] x[n,:]  (d_u, reduced in a fixed order)
// blockDim = (dim/4, rows): thread (tx, ty) owns float4 column tx for rows ty, ty+RY, ... of the CTA's chunk.
template <int R>
__global__ void rows_dot_bwd_kernel(const float* __restrict__ x, const float* __restrict__ u,
                                    const float* __restrict__ d_a_src, const float* __restrict__ d_a_dst, int64_t n,
                                    int dim, int64_t chunk, float* __restrict__ d_x, float* __restrict__ partial) {
  extern __shared__ float sm[];  // [blockDim.y][R][dim]
  const int tx = threadIdx.x, ty = threadIdx.y;
  float4 uv[R], g[R];
#pragma unroll
  for (int r = 0; r < R; ++r) { uv[r] = ldg4(u + (int64_t)r * dim + 4 * tx); g[r] = zero4(); }
  const int64_t begin = blockIdx.x * chunk;
  const int64_t end = begin + chunk < n ? begin + chunk : n;
  for (int64_t row = begin + ty; row < end; row += blockDim.y) {
    const float4 xv = ldg4(x + row * dim + 4 * tx);
    float4 dx = zero4();
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float d = r < R / 2 ? d_a_src[row * (R / 2) + r] : d_a_dst[row * (R / 2) + r - R / 2];
      dx = fma4(d, uv[r], dx);
      g[r] = fma4(d, xv, g[r]);
    }
    st4(d_x + row * dim + 4 * tx, dx);
  }
  float* mine = sm + (size_t)ty * R * dim;
#pragma unroll
  for (int r = 0; r < R; ++r) st4(mine + (size_t)r * dim + 4 * tx, g[r]);
  __syncthreads();
  const int tid = ty * blockDim.x + tx;
  for (int i = tid; i < R * dim; i += blockDim.x * blockDim.y) {
    float t = 0.f;
    for (int yy = 0; yy < (int)blockDim.y; ++yy) t += sm[(size_t)yy * R * dim + i];
    partial[(int64_t)blockIdx.x * R * dim + i] = t;
  }
}

// out[n,c] = (1/H) sum_h agg[n,h,c] + bias[c]
__global__ void __launch_bounds__(kThreads)
head_mean_fwd_kernel(const float* __restrict__ agg, const float* __restrict__ bias, int64_t total4, int c4, int heads,
                     float* __restrict__ out) {
  const float inv = 1.f / (float)heads;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < total4; i += (int64_t)gridDim.x * kThreads) {
    const int64_t node = i / c4;
    const int f = (int)(i % c4);
    const float* row = agg + (node * heads) * (int64_t)(4 * c4) + 4 * f;
    float4 s = zero4();
    for (int hd = 0; hd < heads; ++hd) s = add4(s, ldg4(row + (int64_t)hd * 4 * c4));
    s = scale4(inv, s);
    if (bias != nullptr) s = add4(s, ldg4(bias + 4 * f));
    st4(out + 4 * i, s);
  }
}

// d_agg[n,h,c] = d_out[n,c] / H;  partial[cta][c] = sum over the CTA's rows of d_out[n,c]  (d_bias)
__global__ void head_mean_bwd_kernel(const float* __restrict__ d_out, int64_t n, int c4, int heads, int64_t chunk,
                                     float* __restrict__ d_agg, float* __restrict__ partial) {
  extern __shared__ float sm[];  // [blockDim.y][C]
  const int tx = threadIdx.x, ty = threadIdx.y;
  const float inv = 1.f / (float)heads;
  float4 acc = zero4();
  const int64_t begin = blockIdx.x * chunk;
  const int64_t end = begin + chunk < n ? begin + chunk : n;
  for (int64_t r = begin + ty; r < end; r += blockDim.y) {
    const float4 g = ldg4(d_out + (r * c4 + tx) * 4);
    acc = add4(acc, g);
    if (d_agg != nullptr) {   // (NULL: only the bias gradient — the fused GAT backward expands d_out in registers)
      const float4 gi = scale4(inv, g);
      float* row = d_agg + (r * heads) * (int64_t)(4 * c4) + 4 * tx;
      for (int hd = 0; hd < heads; ++hd) st4(row + (int64_t)hd * 4 * c4, gi);
    }
  }
  st4(sm + (size_t)ty * 4 * c4 + 4 * tx, acc);
  __syncthreads();
  const int tid = ty * blockDim.x + tx;
  for (int i = tid; i < 4 * c4; i += blockDim.x * blockDim.y) {
    float s = 0.f;
    for (int yy = 0; yy < (int)blockDim.y; ++yy) s += sm[(size_t)yy * 4 * c4 + i];
    partial[(int64_t)blockIdx.x * 4 * c4 + i] = s;
  }
}

// lane = column (coalesced), warp w adds parts w, w+8, ...; the eight warp sums are added in warp order
__global__ void __launch_bounds__(256)
column_reduce_kernel(const float* __restrict__ partial, int parts, int width, float* __restrict__ out_a,
                     float* __restrict__ out_b, int width_a) {
  __shared__ float warp_sum[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (i < width)
    for (int p = w; p < parts; p += 8) s += partial[(int64_t)p * width + i];
  warp_sum[w][lane] = s;
  __syncthreads();
  if (w != 0 || i >= width) return;
  float t = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) t += warp_sum[q][lane];
  if (i < width_a) out_a[i] = t; else out_b[i - width_a] = t;
}

int row_parts(int64_t n) {
  int64_t parts = (n + 63) / 64;
  if (parts > 4 * kNumSMs) parts = 4 * kNumSMs;
  return parts < 1 ? 1 : (int)parts;
}

bool gat_shape_ok(int width, int heads) {
  return heads >= 1 && width >= 32 && width <= 1024 && width % heads == 0 && (width / heads) % 4 == 0 && width % 4 == 0;
}

}  // namespace
}  // namespace etpgt

using namespace etpgt;

extern "C" size_t etpgt_gat_aux_workspace_bytes(int64_t num_nodes, int width) {
  return align_up((size_t)row_parts(num_nodes) * 2 * width * sizeof(float)) + 256;
}

extern "C" int etpgt_gat_scores_fwd(const float* h, const float* att_src, const float* att_dst, int64_t num_nodes,
                                    int width, int heads, float* a_src, float* a_dst, etpgt_stream_t stream) {
  ETPGT_REQUIRE(gat_shape_ok(width, heads), "gat_scores: unsupported (heads*channels=%d, heads=%d)", width, heads);
  ETPGT_REQUIRE(num_nodes >= 0 && h && att_src && att_dst && a_src && a_dst, "gat_scores_fwd: bad arguments");
  if (num_nodes == 0) return ETPGT_OK;
  gat_scores_fwd_kernel<<<grid_for(num_nodes, kThreads / 32, 8), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      h, att_src, att_dst, num_nodes, width, width / heads, a_src, a_dst);
  ETPGT_CHECK_LAUNCH("gat_scores_fwd");
  return ETPGT_OK;
}

extern "C" int etpgt_gat_scores_bwd(const float* h, const float* att_src, const float* att_dst, const float* d_a_src,
                                    const float* d_a_dst, int64_t num_nodes, int width, int heads, float* d_h,
                                    float* d_att_src, float* d_att_dst, void* ws, size_t ws_bytes,
                                    etpgt_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ETPGT_REQUIRE(gat_shape_ok(width, heads), "gat_scores: unsupported (heads*channels=%d, heads=%d)", width, heads);
  ETPGT_REQUIRE(num_nodes >= 0 && h && att_src && att_dst && d_a_src && d_a_dst && d_h && d_att_src && d_att_dst,
                "gat_scores_bwd: bad arguments");
  if (ws_bytes < etpgt_gat_aux_workspace_bytes(num_nodes, width)) {
    set_error("gat_scores_bwd: workspace too small");
    return ETPGT_EWORKSPACE;
  }
  float* partial = static_cast<float*>(ws);
  const int parts = num_nodes > 0 ? row_parts(num_nodes) : 0;
  if (parts > 0) {
    const int64_t chunk = (num_nodes + parts - 1) / parts;
    const int tx = width / 4;
    const int ty = 256 / tx > 0 ? 256 / tx : 1;
    const size_t smem = (size_t)ty * 2 * width * sizeof(float);
    gat_scores_bwd_kernel<<<parts, dim3(tx, ty), smem, stream>>>(h, att_src, att_dst, d_a_src, d_a_dst, num_nodes, width,
                                                                 width / heads, chunk, d_h, partial);
    ETPGT_CHECK_LAUNCH("gat_scores_bwd");
  }
  column_reduce_kernel<<<(2 * width + 31) / 32, 256, 0, stream>>>(partial, parts, 2 * width, d_att_src, d_att_dst, width);
  ETPGT_CHECK_LAUNCH("gat_scores_bwd reduce");
  return ETPGT_OK;
}

extern "C" size_t etpgt_gat_input_scores_workspace_bytes(int64_t num_nodes, int dim, int heads) {
  return align_up((size_t)row_parts(num_nodes) * 2 * heads * dim * sizeof(float)) + 256;
}

extern "C" int etpgt_gat_input_scores_fwd(const float* x, const float* u, int64_t num_nodes, int dim, int heads,
                                          float* a_src, float* a_dst, etpgt_stream_t stream) {
  ETPGT_REQUIRE(dim >= 4 && dim % 4 == 0 && (heads == 1 || heads == 2 || heads == 4 || heads == 8),
                "gat_input_scores: dim must be a multiple of 4, heads in {1,2,4,8}");
  ETPGT_REQUIRE(num_nodes >= 0 && x && u && a_src && a_dst, "gat_input_scores_fwd: bad arguments");
  if (num_nodes == 0) return ETPGT_OK;
  const int grid = grid_for(num_nodes, kThreads / 32, 16);
#define CALL(R) rows_dot_fwd_kernel<R><<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(x, u, num_nodes, dim, a_src, a_dst)
  switch (heads) {
    case 1: CALL(2); break;
    case 2: CALL(4); break;
    case 4: CALL(8); break;
    default: CALL(16); break;
  }
#undef CALL
  ETPGT_CHECK_LAUNCH("gat_input_scores_fwd");
  return ETPGT_OK;
}

extern "C" int etpgt_gat_input_scores_bwd(const float* x, const float* u, const float* d_a_src, const float* d_a_dst,
                                          int64_t num_nodes, int dim, int heads, float* d_x, float* d_u, void* ws,
                                          size_t ws_bytes, etpgt_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ETPGT_REQUIRE(dim >= 4 && dim % 4 == 0 && dim <= 1024 && (heads == 1 || heads == 2 || heads == 4 || heads == 8),
                "gat_input_scores: dim must be a multiple of 4 up to 1024, heads in {1,2,4,8}");
  ETPGT_REQUIRE(num_nodes >= 0 && x && u && d_a_src && d_a_dst && d_x && d_u, "gat_input_scores_bwd: bad arguments");
  if (ws_bytes < etpgt_gat_input_scores_workspace_bytes(num_nodes, dim, heads)) {
    set_error("gat_input_scores_bwd: workspace too small");
    return ETPGT_EWORKSPACE;
  }
  float* partial = static_cast<float*>(ws);
  const int parts = num_nodes > 0 ? row_parts(num_nodes) : 0;
  const int r = 2 * heads;
  if (parts > 0) {
    const int64_t chunk = (num_nodes + parts - 1) / parts;
    const int tx = dim / 4;
    int ty = 256 / tx > 0 ? 256 / tx : 1;
    while (ty > 1 && (size_t)ty * r * dim * sizeof(float) > 96 * 1024) ty /= 2;
    const size_t smem = (size_t)ty * r * dim * sizeof(float);
#define CALL(R)                                                                                                  \
  {                                                                                                              \
    if (smem > 48 * 1024)                                                                                        \
      cudaFuncSetAttribute(rows_dot_bwd_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
    rows_dot_bwd_kernel<R><<<parts, dim3(tx, ty), smem, stream>>>(x, u, d_a_src, d_a_dst, num_nodes, dim, chunk, d_x, \
                                                                 partial);                                       \
  }
    switch (heads) {
      case 1: CALL(2) break;
      case 2: CALL(4) break;
      case 4: CALL(8) break;
      default: CALL(16) break;
    }
#undef CALL
    ETPGT_CHECK_LAUNCH("gat_input_scores_bwd");
  }
  column_reduce_kernel<<<(r * dim + 31) / 32, 256, 0, stream>>>(partial, parts, r * dim, d_u, nullptr, r * dim);
  ETPGT_CHECK_LAUNCH("gat_input_scores_bwd reduce");
  return ETPGT_OK;
}

extern "C" int etpgt_head_mean_fwd(const float* agg, const float* bias, int64_t num_nodes, int heads, int channels,
                                   float* out, etpgt_stream_t stream) {
  ETPGT_REQUIRE(heads >= 1 && channels >= 4 && channels % 4 == 0, "head_mean: channels must be a multiple of 4");
  ETPGT_REQUIRE(num_nodes >= 0 && agg && out, "head_mean_fwd: bad arguments");
  if (num_nodes == 0) return ETPGT_OK;
  const int64_t total4 = num_nodes * (channels / 4);
  head_mean_fwd_kernel<<<grid_for(total4, kThreads * 2, 8), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      agg, bias, total4, channels / 4, heads, out);
  ETPGT_CHECK_LAUNCH("head_mean_fwd");
  return ETPGT_OK;
}

extern "C" int etpgt_head_mean_bwd(const float* d_out, int64_t num_nodes, int heads, int channels, float* d_agg,
                                   float* d_bias, void* ws, size_t ws_bytes, etpgt_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ETPGT_REQUIRE(heads >= 1 && channels >= 4 && channels % 4 == 0 && channels <= 1024,
                "head_mean: channels must be a multiple of 4 up to 1024");
  ETPGT_REQUIRE(num_nodes >= 0 && d_out && (d_agg || d_bias), "head_mean_bwd: bad arguments");
  if (ws_bytes < etpgt_gat_aux_workspace_bytes(num_nodes, channels)) {
    set_error("head_mean_bwd: workspace too small");
    return ETPGT_EWORKSPACE;
  }
  float* partial = static_cast<float*>(ws);
  const int parts = num_nodes > 0 ? row_parts(num_nodes) : 0;
  if (parts > 0) {
    const int64_t chunk = (num_nodes + parts - 1) / parts;
    const int tx = channels / 4;
    const int ty = 256 / tx > 0 ? 256 / tx : 1;
    const size_t smem = (size_t)ty * channels * sizeof(float);
    head_mean_bwd_kernel<<<parts, dim3(tx, ty), smem, stream>>>(d_out, num_nodes, channels / 4, heads, chunk, d_agg,
                                                                partial);
    ETPGT_CHECK_LAUNCH("head_mean_bwd");
  }
  if (d_bias != nullptr) {
    column_reduce_kernel<<<(channels + 31) / 32, 256, 0, stream>>>(partial, parts, channels, d_bias, nullptr, channels);
    ETPGT_CHECK_LAUNCH("head_mean_bwd reduce");
  }
  return ETPGT_OK;
}
