// f4: the sparse operator of the Laplacian positional-encoding eigen solver
// (etpgt/encodings/laplacian_pe.py:19-66: PyG get_laplacian(normalization="sym") -> scipy eigsh(k+1, "SM")).
//
// The solver (etpgt_b200/encodings/laplacian_pe.py, Chebyshev-filtered subspace iteration) spends its time in
//   Y = alpha * (L X) + beta * X + gamma * Z,      L = I - D^-1/2 A D^-1/2,
// the three-term Chebyshev recurrence on a block of b <= 32 vectors, fp64.  L is never formed: the kernel
// walks the graph's CSR rows, (L x)_i = x_i - s_i * sum_{j in N(i)} s_j x_j with s = deg^-1/2 (0 for isolated
// nodes, as PyG's masked_fill of inf).  One warp per row, one lane per vector of the block: the b doubles of
// a gathered row are one contiguous <= 256-byte segment, and the whole block (n x b x 8 bytes: 21 MB at 82k
// nodes, 256 MB at 1M) lives in the L2, so the kernel is bound by the HBM read of the index.
#include "common.cuh"

namespace etpgt {
namespace {

constexpr int kRowsPerCta = 8;

__global__ void __launch_bounds__(kRowsPerCta * 32)
lap_sym_block_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                     const double* __restrict__ scale, int64_t n, int b, const double* __restrict__ X,
                     const double* __restrict__ Z, double alpha, double beta, double gamma, double* __restrict__ Y) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * kRowsPerCta;
  for (int64_t i = (int64_t)blockIdx.x * kRowsPerCta + (threadIdx.x >> 5); i < n; i += warps) {
    const int64_t e0 = rowptr[i], e1 = rowptr[i + 1];
    double acc0 = 0.0, acc1 = 0.0;
    int64_t e = e0;
    // two gathered rows in flight per lane; the column index and its scale are broadcast loads
    for (; e + 1 < e1; e += 2) {
      const int32_t j0 = col[e], j1 = col[e + 1];
      const double s0 = scale[j0], s1 = scale[j1];
      if (lane < b) {
        acc0 = fma(s0, X[(int64_t)j0 * b + lane], acc0);
        acc1 = fma(s1, X[(int64_t)j1 * b + lane], acc1);
      }
    }
    if (e < e1) {
      const int32_t j0 = col[e];
      if (lane < b) acc0 = fma(scale[j0], X[(int64_t)j0 * b + lane], acc0);
    }
    if (lane < b) {
      const double xi = X[i * b + lane];
      const double lx = xi - scale[i] * (acc0 + acc1);
      double out = alpha * lx + beta * xi;
      if (Z != nullptr) out = fma(gamma, Z[i * b + lane], out);
      Y[i * b + lane] = out;
    }
  }
}

}  // namespace
}  // namespace etpgt

using namespace etpgt;

extern "C" int etpgt_lap_sym_block(const int64_t* rowptr, const int32_t* col, const double* scale, int64_t n, int b,
                                   const double* x, const double* z, double alpha, double beta, double gamma,
                                   double* y, etpgt_stream_t stream) {
  ETPGT_REQUIRE(n >= 0 && b >= 1 && b <= 32, "lap_sym_block: block width must be in [1, 32]");
  ETPGT_REQUIRE(rowptr != nullptr && scale != nullptr && x != nullptr && y != nullptr && x != y && z != y,
                "lap_sym_block: null or aliased argument (y must not alias x or z)");
  if (n == 0) return ETPGT_OK;
  lap_sym_block_kernel<<<grid_for(n, kRowsPerCta, 8), kRowsPerCta * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      rowptr, col, scale, n, b, x, z, alpha, beta, gamma, y);
  ETPGT_CHECK_LAUNCH("lap_sym_block");
  return ETPGT_OK;
}
