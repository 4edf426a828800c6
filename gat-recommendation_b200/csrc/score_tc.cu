// a9: full-catalogue scoring on the 5th-generation tensor cores with a fused top-k epilogue.
//   scores[b, i] = <S[b, :], E[i, :]>  (bf16 operands, fp32 accumulation in TMEM), top-k per row
// etpgt/model/base.py:59-78 (`torch.matmul(S, E.t())` + `torch.topk`).  The [B, I] score matrix is
// never written: the epilogue warps read each 128 x 256 accumulator tile straight out of TMEM and
// keep a per-row k-list; only [B, splits, k] candidates leave the SM, merged by etpgt_topk_merge.
//
// Structure (one CTA = one unit of 128 sessions x a contiguous range of 256-item tiles):
//   warp 0      TMA producer: session tile once (4 k-blocks of 128x64 bf16, SWIZZLE_128B), then a
//               ring of item k-blocks (256x64 bf16) with mbarrier full/empty hand-shakes
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128x256x16, kind::f16),
//               accumulators double-buffered in the 512 TMEM columns
//   warps 2-5   epilogue: tcgen05.ld 32x32b.x32 (lane = session row), threshold-pruned insertion
//               into the row's k-list (strict '>' while ids ascend => ties keep the lower id)
// Tensor-bound: 2*128*256*DIM flop per tile against DIM/16 MMA instructions of 128 cycles each.
#include <cuda.h>
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"

namespace etpgt {
namespace {

constexpr int BLOCK_M = 128;   // sessions per CTA (TMEM lanes)
constexpr int BLOCK_N = 256;   // items per accumulator tile (TMEM columns)
constexpr int BLOCK_K = 64;    // bf16 elements per 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kStages = 3;     // ring of item k-blocks
constexpr int kAccStages = 2;  // TMEM accumulator double buffer
constexpr int kTmemCols = 512;
constexpr int kThreads = 192;  // warp0 TMA, warp1 MMA, warps 2..5 epilogue
constexpr int kMaxKTc = 32;
constexpr int kMaxDim = 256;
constexpr uint32_t A_KBLOCK_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr uint32_t B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;   // 32 KB

// ------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch error reported to the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long start = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - start > 4000000000LL) {
      printf("etpgt score_tc: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* smem, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols));
}

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start address >> 4 in bits [0,14), LBO (unused for swizzled K-major) = 1 in [16,30),
// SBO = 1024 B (eight 128-byte rows) >> 4 in [32,46), version 1 in [46,48), layout type 2 in [61,64).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// cute::UMMA::InstrDescriptor: c_format F32 (1) at [4,6), a/b format BF16 (1) at [7,10)/[10,13),
// K-major A and B, N >> 3 at [17,23), M >> 4 at [24,29).
constexpr uint32_t kInstrDesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) |
                                ((uint32_t)(BLOCK_M >> 4) << 24);

struct __align__(8) Barriers {
  uint64_t a_full;
  uint64_t b_full[kStages];
  uint64_t b_empty[kStages];
  uint64_t acc_full[kAccStages];
  uint64_t acc_empty[kAccStages];
  uint32_t tmem_base;
};

// Per-row candidate list: vals[k][128] / idxs[k][128] (row fastest => conflict-free), the row's
// current minimum kept in registers.
//
// Inserting straight from the scan would serialise the warp: the 32 rows of a warp accept
// candidates at different columns, and every accept drags the whole warp through a k-step rescan
// (measured: 21 ms for 23,861 x 82,174, 3 % of the tensor peak).  Instead a candidate that beats
// the row's (possibly stale) threshold is only APPENDED to a small pending buffer, and the
// pending buffers of all 32 rows are merged into the lists together, in lockstep, when one of
// them is about to fill up or the range ends.  The threshold only ever rises, so a stale one
// admits extra candidates but never loses one; order inside the pending buffer is arrival (id)
// order, so the strict '>' tie rule is preserved.
constexpr int kPendingCap = 16;
constexpr int kPendingBlock = 8;  // columns scanned between two overflow checks

struct RowList {
  float* vals;
  int32_t* idxs;
  float* pend_vals;
  int32_t* pend_idxs;
  int k;
  int pending;
  float thr;
  int min_pos;
  __device__ __forceinline__ void init(float* v, int32_t* i, float* pv, int32_t* pi, int row, int kk) {
    vals = v + row; idxs = i + row; pend_vals = pv + row; pend_idxs = pi + row;
    k = kk; thr = -INFINITY; min_pos = 0; pending = 0;
    for (int t = 0; t < k; ++t) { vals[t * BLOCK_M] = -INFINITY; idxs[t * BLOCK_M] = INT32_MAX; }
  }
  __device__ __forceinline__ void append(float v, int32_t id) {
    pend_vals[pending * BLOCK_M] = v;
    pend_idxs[pending * BLOCK_M] = id;
    ++pending;
  }
  // all 32 lanes of the warp call this together
  __device__ __forceinline__ void flush() {
    int most = pending;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) most = max(most, __shfl_xor_sync(0xffffffffu, most, off));
    for (int p = 0; p < most; ++p)
      if (p < pending) offer(pend_vals[p * BLOCK_M], pend_idxs[p * BLOCK_M]);
    pending = 0;
  }
  // the entry to evict next: lowest score, and among equal scores the highest id
  __device__ __forceinline__ void rescan() {
    float mv = vals[0];
    int mi = idxs[0], mp = 0;
    for (int t = 1; t < k; ++t) {
      const float v = vals[t * BLOCK_M];
      const int i = idxs[t * BLOCK_M];
      if (v < mv || (v == mv && i > mi)) { mv = v; mi = i; mp = t; }
    }
    thr = mv; min_pos = mp;
  }
  __device__ __forceinline__ void offer(float v, int32_t id) {
    if (v > thr) {  // ids arrive in ascending order: an equal score never displaces an earlier id
      vals[min_pos * BLOCK_M] = v;
      idxs[min_pos * BLOCK_M] = id;
      rescan();
    }
  }
};

template <int NUM_KB>  // DIM / 64
__global__ void __launch_bounds__(kThreads, 1)
score_topk_tc_kernel(const __grid_constant__ CUtensorMap map_sess, const __grid_constant__ CUtensorMap map_items,
                     int64_t batch, int64_t num_items, int k, int tiles_per_split, int splits, int64_t id_base,
                     float* __restrict__ cand_val, int64_t* __restrict__ cand_idx) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;                                   // NUM_KB x 16 KB
  uint8_t* smem_b = smem_a + NUM_KB * A_KBLOCK_BYTES;       // kStages x 32 KB
  float* list_vals = reinterpret_cast<float*>(smem_b + kStages * B_STAGE_BYTES);
  int32_t* list_idxs = reinterpret_cast<int32_t*>(list_vals + kMaxKTc * BLOCK_M);
  float* pend_vals = reinterpret_cast<float*>(list_idxs + kMaxKTc * BLOCK_M);
  int32_t* pend_idxs = reinterpret_cast<int32_t*>(pend_vals + kPendingCap * BLOCK_M);
  Barriers* bars = reinterpret_cast<Barriers*>(pend_idxs + kPendingCap * BLOCK_M);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int split = blockIdx.x;
  const int m_tile = blockIdx.y;
  const int64_t total_tiles = (num_items + BLOCK_N - 1) / BLOCK_N;
  const int64_t tile_begin = (int64_t)split * tiles_per_split;
  const int64_t tile_end = tile_begin + tiles_per_split < total_tiles ? tile_begin + tiles_per_split : total_tiles;
  const int num_tiles = tile_end > tile_begin ? (int)(tile_end - tile_begin) : 0;

  if (threadIdx.x == 0) {
    mbar_init(&bars->a_full, 1);
    for (int s = 0; s < kStages; ++s) { mbar_init(&bars->b_full[s], 1); mbar_init(&bars->b_empty[s], 1); }
    for (int s = 0; s < kAccStages; ++s) { mbar_init(&bars->acc_full[s], 1); mbar_init(&bars->acc_empty[s], 4); }
    fence_barrier_init();
    tma_prefetch_desc(&map_sess);
    tma_prefetch_desc(&map_items);
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer (one elected lane) =====================
    if (lane == 0 && num_tiles > 0) {
      mbar_expect_tx(&bars->a_full, NUM_KB * A_KBLOCK_BYTES);
      for (int kb = 0; kb < NUM_KB; ++kb)
        tma_load_2d(&map_sess, &bars->a_full, smem_a + kb * A_KBLOCK_BYTES, kb * BLOCK_K, m_tile * BLOCK_M);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < num_tiles; ++t) {
        const int row0 = (int)((tile_begin + t) * BLOCK_N);
        for (int kb = 0; kb < NUM_KB; ++kb) {
          mbar_wait(&bars->b_empty[stage], phase ^ 1);
          mbar_expect_tx(&bars->b_full[stage], B_STAGE_BYTES);
          tma_load_2d(&map_items, &bars->b_full[stage], smem_b + stage * B_STAGE_BYTES, kb * BLOCK_K, row0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected lane) =====================
    if (lane == 0 && num_tiles > 0) {
      mbar_wait(&bars->a_full, 0);
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < num_tiles; ++t) {
        const int acc = t & 1;
        const uint32_t acc_phase = (uint32_t)(t >> 1) & 1;
        mbar_wait(&bars->acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = 0; kb < NUM_KB; ++kb) {
          mbar_wait(&bars->b_full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + kb * A_KBLOCK_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
            const uint64_t da = make_desc_sw128(a_addr + kk * UMMA_K * 2);
            const uint64_t db = make_desc_sw128(b_addr + kk * UMMA_K * 2);
            umma_bf16(tmem_d, da, db, kInstrDesc, (kb | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&bars->b_empty[stage]);  // frees the smem stage once these MMAs have read it
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&bars->acc_full[acc]);     // accumulator tile complete
      }
    }
  } else {
    // ===================== epilogue: 4 warps, TMEM lane quarter = warp % 4 =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;       // session row inside the tile == TMEM lane
    RowList list;
    list.init(list_vals, list_idxs, pend_vals, pend_idxs, row, k);
    for (int t = 0; t < num_tiles; ++t) {
      const int acc = t & 1;
      const uint32_t acc_phase = (uint32_t)(t >> 1) & 1;
      mbar_wait(&bars->acc_full[acc], acc_phase);
      tc_fence_after();
      const int64_t item0 = (tile_begin + t) * BLOCK_N;
      const int base = t * BLOCK_N;  // id relative to the first item of this CTA's range
      const int limit = num_items - item0 < BLOCK_N ? (int)(num_items - item0) : BLOCK_N;  // real columns
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BLOCK_N);
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + (uint32_t)c0, v);
        float mx = __uint_as_float(v[0]);
#pragma unroll
        for (int j = 1; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
        if (__any_sync(0xffffffffu, mx > list.thr)) {  // warp-uniform: flush() below is a warp collective
#pragma unroll
          for (int jb = 0; jb < 32; jb += kPendingBlock) {
            if (__any_sync(0xffffffffu, list.pending > kPendingCap - kPendingBlock)) list.flush();
#pragma unroll
            for (int j = jb; j < jb + kPendingBlock; ++j) {
              const float s = __uint_as_float(v[j]);
              if (s > list.thr && c0 + j < limit) list.append(s, base + c0 + j);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc_empty[acc]);
    }
    list.flush();
    const int64_t grow = (int64_t)m_tile * BLOCK_M + row;
    if (grow < batch) {
      for (int t = 0; t < k; ++t) {
        const float v = list.vals[t * BLOCK_M];
        const int32_t li = list.idxs[t * BLOCK_M];
        const int64_t o = (grow * splits + split) * k + t;
        cand_val[o] = li == INT32_MAX ? -INFINITY : v;
        cand_idx[o] = li == INT32_MAX ? INT64_MAX : id_base + tile_begin * BLOCK_N + li;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = ldg4(src + 4 * i);
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 packed;
    packed.x = *reinterpret_cast<uint32_t*>(&lo);
    packed.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(dst + 4 * i) = packed;
  }
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [rows, dim] bf16 row-major -> boxes of box_rows x 64 elements, 128-byte swizzle, zero OOB fill
bool make_map(CUtensorMap* map, const void* base, int64_t rows, int dim, int box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)dim * 2};
  const cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
  const cuuint32_t estride[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estride,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct TcPlan {
  int m_tiles, splits, tiles_per_split;
};

TcPlan tc_plan(int64_t batch, int64_t num_items) {
  TcPlan p;
  p.m_tiles = (int)((batch + BLOCK_M - 1) / BLOCK_M);
  const int64_t total_tiles = (num_items + BLOCK_N - 1) / BLOCK_N;
  int64_t want = (2 * kNumSMs + p.m_tiles - 1) / p.m_tiles;  // about two waves of CTAs in total
  if (want > total_tiles) want = total_tiles;
  if (want < 1) want = 1;
  p.tiles_per_split = (int)((total_tiles + want - 1) / want);
  p.splits = (int)((total_tiles + p.tiles_per_split - 1) / p.tiles_per_split);
  return p;
}

size_t tc_smem_bytes(int num_kb) {
  return 1024 + (size_t)num_kb * A_KBLOCK_BYTES + (size_t)kStages * B_STAGE_BYTES +
         (size_t)(kMaxKTc + kPendingCap) * BLOCK_M * 8 + sizeof(Barriers) + 64;
}

}  // namespace
}  // namespace etpgt

using namespace etpgt;

extern "C" int etpgt_f32_to_bf16(const float* src, void* dst, int64_t n, etpgt_stream_t stream) {
  ETPGT_REQUIRE(n >= 0 && n % 4 == 0, "f32_to_bf16: element count must be a multiple of 4");
  if (n == 0) return ETPGT_OK;
  f32_to_bf16_kernel<<<grid_for(n / 4, 256 * 4, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), n / 4);
  ETPGT_CHECK_LAUNCH("f32_to_bf16");
  return ETPGT_OK;
}

extern "C" size_t etpgt_score_topk_bf16_workspace_bytes(int64_t batch, int64_t num_items, int k) {
  if (batch <= 0 || num_items <= 0 || k <= 0) return 256;
  const TcPlan p = tc_plan(batch, num_items);
  const size_t m = (size_t)batch * p.splits * k;
  return align_up(m * sizeof(float)) + align_up(m * sizeof(int64_t)) + 256;
}

extern "C" int etpgt_score_topk_bf16(const void* sess_bf16, const void* table_bf16, int64_t batch,
                                     int64_t num_items, int dim, int k, int64_t id_base, float* top_val,
                                     int64_t* top_idx, void* ws, size_t ws_bytes, etpgt_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ETPGT_REQUIRE(dim == 64 || dim == 128 || dim == 192 || dim == 256,
                "score_topk_bf16: dim %d must be a multiple of 64 up to %d", dim, kMaxDim);
  ETPGT_REQUIRE(batch >= 0 && num_items >= 1 && num_items < (int64_t(1) << 31), "score_topk_bf16: bad sizes");
  ETPGT_REQUIRE(k >= 1 && k <= kMaxKTc && k <= num_items, "score_topk_bf16: k=%d must be in [1, min(%d, num_items)]",
                k, kMaxKTc);
  ETPGT_REQUIRE(((uintptr_t)sess_bf16 & 15) == 0 && ((uintptr_t)table_bf16 & 15) == 0,
                "score_topk_bf16: operands must be 16-byte aligned");
  if (ws_bytes < etpgt_score_topk_bf16_workspace_bytes(batch, num_items, k)) {
    set_error("score_topk_bf16: workspace too small");
    return ETPGT_EWORKSPACE;
  }
  if (batch == 0) return ETPGT_OK;
  const TcPlan p = tc_plan(batch, num_items);
  Workspace w(ws, ws_bytes);
  const size_t m = (size_t)batch * p.splits * k;
  float* cand_val = w.take<float>(m);
  int64_t* cand_idx = w.take<int64_t>(m);
  CUtensorMap map_sess, map_items;
  if (!make_map(&map_sess, sess_bf16, batch, dim, BLOCK_M) || !make_map(&map_items, table_bf16, num_items, dim, BLOCK_N)) {
    set_error("score_topk_bf16: cuTensorMapEncodeTiled failed");
    return ETPGT_ECUDA;
  }
  const int num_kb = dim / BLOCK_K;
  const size_t smem = tc_smem_bytes(num_kb);
  const dim3 grid(p.splits, p.m_tiles);
#define LAUNCH(NKB)                                                                                          \
  {                                                                                                          \
    cudaFuncSetAttribute(score_topk_tc_kernel<NKB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    score_topk_tc_kernel<NKB><<<grid, kThreads, smem, stream>>>(map_sess, map_items, batch, num_items, k,     \
                                                               p.tiles_per_split, p.splits, id_base, cand_val, \
                                                               cand_idx);                                    \
  }
  switch (num_kb) {
    case 1: LAUNCH(1) break;
    case 2: LAUNCH(2) break;
    case 3: LAUNCH(3) break;
    default: LAUNCH(4) break;
  }
#undef LAUNCH
  ETPGT_CHECK_LAUNCH("score_topk_tc");
  return etpgt_topk_merge(cand_val, cand_idx, batch, p.splits, k, top_val, top_idx, stream_);
}
