// a9: full-catalogue scoring on the 5th-generation tensor cores with a fused top-k epilogue.
//   scores[b, i] = <S[b, :], E[i, :]>  (bf16 operands, fp32 accumulation in TMEM), top-k per row
// etpgt/model/base.py:59-78 (`torch.matmul(S, E.t())` + `torch.topk`).  The [B, I] score matrix is
// never written: the epilogue warps read each 128 x 256 accumulator tile straight out of TMEM and
// keep per-row k-lists; only [B, parts, k] candidates leave the SM, merged by etpgt_topk_merge.
//
// Structure (one CTA = one unit of 128 sessions x a contiguous range of 256-item tiles):
//   warp 0      TMA producer: session tile once (DIM/64 k-blocks of 128x64 bf16, SWIZZLE_128B), then
//               a ring of item k-blocks (256x64 bf16) with mbarrier full/empty hand-shakes
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128x256x16, kind::f16),
//               accumulators double-buffered in the 512 TMEM columns
//   warps 2-5   epilogue, one warp per TMEM lane quarter: tcgen05.ld 32x32b.x32 (lane = session row),
//               threshold-pruned candidates, register-resident sorted k-list per row
// Tensor-bound: 2*128*256*DIM flop per tile against DIM/16 MMA instructions of 128 cycles each.
#include <math.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace etpgt {
namespace {

using namespace tc;

constexpr int BLOCK_M = 128;   // sessions per CTA (TMEM lanes)
constexpr int BLOCK_N = 256;   // items per accumulator tile (TMEM columns)
constexpr int kMaxStages = 3;  // ring of item k-blocks
constexpr int kAccStages = 2;  // TMEM accumulator double buffer
constexpr int kTmemCols = 512;
constexpr int kEpilogueWarps = 4;
constexpr int kThreads = 32 * (2 + kEpilogueWarps);
constexpr int kMaxKTc = 32;
constexpr uint32_t A_KBLOCK_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr uint32_t B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;   // 32 KB
constexpr uint32_t kInstrDesc = instr_desc_bf16(BLOCK_M, BLOCK_N);

struct __align__(8) Barriers {
  uint64_t a_full;
  uint64_t b_full[kMaxStages];
  uint64_t b_empty[kMaxStages];
  uint64_t acc_full[kAccStages];
  uint64_t acc_empty[kAccStages];
  uint32_t tmem_base;
};

// Per-row candidate list, REGISTER resident and sorted (score desc; equal scores keep arrival =
// ascending id order), KCAP >= k entries; the row's threshold is the last entry.
//
// Inserting straight from the scan would serialise the warp: the 32 rows of a warp accept
// candidates at different columns, and every accept would drag the whole warp through the insert
// (first version, shared-memory lists with a k-step rescan per accept: 21 ms for 23,861 x 82,174,
// 3 % of the tensor peak; pending buffers: 2.6 ms; the rescan's dependent shared-memory loads
// were still the largest cost).  Now a candidate that beats the row's (possibly stale) threshold
// is only APPENDED to a small pending buffer in shared memory, and the pending buffers of all 32
// rows are merged into the register lists together, in lockstep, by a branch-free insertion
// (compare, then shift everything behind the insertion point) when one of them is about to fill
// up or the range ends.  The threshold only ever rises, so a stale one admits extra candidates but
// never loses one; arrival order is ascending id, so strict '>' keeps the lower id on ties.
constexpr int kPendingCap = 40;  // a whole 32-column chunk always fits after the overflow check
constexpr int kScanBlock = 4;    // columns per second-level maximum

template <int KCAP>
struct RowList {
  float lv[KCAP];
  int32_t li[KCAP];
  float* pend_vals;
  int32_t* pend_idxs;
  int pending;
  float thr;
  __device__ __forceinline__ void init(float* pv, int32_t* pi) {
    pend_vals = pv; pend_idxs = pi; pending = 0; thr = -INFINITY;
#pragma unroll
    for (int t = 0; t < KCAP; ++t) { lv[t] = -INFINITY; li[t] = INT32_MAX; }
  }
  __device__ __forceinline__ void append(float v, int32_t id) {
    pend_vals[pending * BLOCK_M] = v;
    pend_idxs[pending * BLOCK_M] = id;
    ++pending;
  }
  __device__ __forceinline__ void insert(float v, int32_t id) {
    bool shifting = false;
#pragma unroll
    for (int t = 0; t < KCAP; ++t) {
      shifting = shifting || v > lv[t];
      const float ov = lv[t];
      const int32_t oi = li[t];
      lv[t] = shifting ? v : ov;
      li[t] = shifting ? id : oi;
      v = shifting ? ov : v;
      id = shifting ? oi : id;
    }
    thr = lv[KCAP - 1];
  }
  // all 32 lanes of the warp call this together
  __device__ __forceinline__ void flush() {
    int most = pending;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) most = max(most, __shfl_xor_sync(0xffffffffu, most, off));
    for (int p = 0; p < most; ++p) {
      const bool on = p < pending;
      const float v = on ? pend_vals[p * BLOCK_M] : -INFINITY;   // -inf never beats any entry
      const int32_t id = on ? pend_idxs[p * BLOCK_M] : INT32_MAX;
      insert(v, id);
    }
    pending = 0;
  }
};

template <int NUM_KB, int KCAP>  // DIM / 64, list capacity
__global__ void __launch_bounds__(kThreads, 1)
score_topk_tc_kernel(const __grid_constant__ CUtensorMap map_sess, const __grid_constant__ CUtensorMap map_items,
                     int64_t batch, int64_t num_items, int k, int stages_and_flags, int tiles_per_split, int splits,
                     int64_t id_base, float* __restrict__ cand_val, int64_t* __restrict__ cand_idx) {
  const int stages = stages_and_flags & 255;
  const bool debug_skip_scan = (stages_and_flags & 256) != 0;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;                                   // NUM_KB x 16 KB
  uint8_t* smem_b = smem_a + NUM_KB * A_KBLOCK_BYTES;       // stages x 32 KB
  float* pend_vals = reinterpret_cast<float*>(smem_b + stages * B_STAGE_BYTES);   // [cap][128]
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
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&bars->b_full[s], 1); mbar_init(&bars->b_empty[s], 1); }
    for (int s = 0; s < kAccStages; ++s) { mbar_init(&bars->acc_full[s], 1); mbar_init(&bars->acc_empty[s], kEpilogueWarps); }
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
          if (++stage == stages) { stage = 0; phase ^= 1; }
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
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&bars->acc_full[acc]);     // accumulator tile complete
      }
    }
  } else {
    // ===================== epilogue: 4 warps, TMEM lane quarter = warp % 4 =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;       // session row inside the tile == TMEM lane
    RowList<KCAP> list;
    list.init(pend_vals + row, pend_idxs + row);
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
        float v[32];
        {
          uint32_t raw[32];
          tmem_ld_32x32(taddr + (uint32_t)c0, raw);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
        }
        if (limit - c0 < 32) {  // only the table's last tile: columns past the end never qualify
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = c0 + j < limit ? v[j] : -INFINITY;
        }
        // two-level maximum; every branch below is warp-uniform (votes), the appends are predicated,
        // so the scan costs no divergence and the warp only touches the 4-column blocks that hold a
        // candidate of some row
        float gmax[32 / kScanBlock];
#pragma unroll
        for (int g = 0; g < 32 / kScanBlock; ++g)
          gmax[g] = fmaxf(fmaxf(v[4 * g], v[4 * g + 1]), fmaxf(v[4 * g + 2], v[4 * g + 3]));
        const float mx = fmaxf(fmaxf(fmaxf(gmax[0], gmax[1]), fmaxf(gmax[2], gmax[3])),
                               fmaxf(fmaxf(gmax[4], gmax[5]), fmaxf(gmax[6], gmax[7])));
        if (debug_skip_scan) { if (mx == 123456.75f) list.thr = mx; continue; }  // profiling only (ETPGT_SCORE_DEBUG=1): garbage results
        if (__any_sync(0xffffffffu, mx > list.thr)) {
          if (__any_sync(0xffffffffu, list.pending > kPendingCap - 32)) list.flush();  // room for a whole chunk
          const float thr = list.thr;
#pragma unroll
          for (int g = 0; g < 32 / kScanBlock; ++g) {
            if (__any_sync(0xffffffffu, gmax[g] > thr)) {
#pragma unroll
              for (int j = g * kScanBlock; j < (g + 1) * kScanBlock; ++j)
                if (v[j] > thr) list.append(v[j], base + c0 + j);
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
      const int64_t o = (grow * splits + split) * k;
#pragma unroll
      for (int t = 0; t < KCAP; ++t) {
        if (t < k) {
          cand_val[o + t] = list.li[t] == INT32_MAX ? -INFINITY : list.lv[t];
          cand_idx[o + t] = list.li[t] == INT32_MAX ? INT64_MAX : id_base + tile_begin * BLOCK_N + list.li[t];
        }
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

struct TcPlan {
  int m_tiles, splits, tiles_per_split;
};

TcPlan tc_plan(int64_t batch, int64_t num_items) {
  TcPlan p;
  p.m_tiles = (int)((batch + BLOCK_M - 1) / BLOCK_M);
  const int64_t total_tiles = (num_items + BLOCK_N - 1) / BLOCK_N;
  // One CTA per SM.  Pick the number of item-range splits whose CTA count fills whole waves of the
  // 148 SMs well, charging each extra split for its per-range warm-up (lists restart empty).
  int64_t max_splits = total_tiles / 8 > 1 ? total_tiles / 8 : 1;
  if (max_splits > 32) max_splits = 32;
  int best = 1;
  double best_score = -1.0;
  for (int64_t s = 1; s <= max_splits; ++s) {
    const int64_t per = (total_tiles + s - 1) / s;
    const int64_t real = (total_tiles + per - 1) / per;
    const int64_t ctas = real * p.m_tiles;
    const int64_t waves = (ctas + kNumSMs - 1) / kNumSMs;
    double eff = (double)ctas / (double)(waves * kNumSMs);
    const double score = eff - 0.03 * (double)s;
    if (score > best_score) { best_score = score; best = (int)s; }
  }
  if (const char* forced = getenv("ETPGT_SCORE_SPLITS")) {  // tuning knob
    const int f = atoi(forced);
    if (f >= 1 && f <= max_splits) best = f;
  }
  p.tiles_per_split = (int)((total_tiles + best - 1) / best);
  p.splits = (int)((total_tiles + p.tiles_per_split - 1) / p.tiles_per_split);
  return p;
}

int tc_stages(int, int) { return kMaxStages; }

size_t tc_smem_bytes(int num_kb, int, int stages) {
  return 1024 + (size_t)num_kb * A_KBLOCK_BYTES + (size_t)stages * B_STAGE_BYTES +
         (size_t)kPendingCap * BLOCK_M * 8 + sizeof(Barriers) + 64;
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
                "score_topk_bf16: dim %d must be a multiple of 64 up to 256", dim);
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
  if (!make_map_bf16(&map_sess, sess_bf16, batch, dim, dim, BLOCK_M) ||
      !make_map_bf16(&map_items, table_bf16, num_items, dim, dim, BLOCK_N)) {
    set_error("score_topk_bf16: cuTensorMapEncodeTiled failed");
    return ETPGT_ECUDA;
  }
  const int num_kb = dim / BLOCK_K;
  const int stages = tc_stages(num_kb, k);
  const size_t smem = tc_smem_bytes(num_kb, k, stages);
  const dim3 grid(p.splits, p.m_tiles);
#define LAUNCH2(NKB, KC)                                                                                        \
  {                                                                                                             \
    cudaFuncSetAttribute(score_topk_tc_kernel<NKB, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    score_topk_tc_kernel<NKB, KC><<<grid, kThreads, smem, stream>>>(map_sess, map_items, batch, num_items, k, stages | (getenv("ETPGT_SCORE_DEBUG") ? 256 : 0), \
                                                                   p.tiles_per_split, p.splits, id_base, cand_val, \
                                                                   cand_idx);                                   \
  }
#define LAUNCH(NKB)                    \
  {                                    \
    if (k <= 10) LAUNCH2(NKB, 10)      \
    else if (k <= 20) LAUNCH2(NKB, 20) \
    else LAUNCH2(NKB, 32)              \
  }
  switch (num_kb) {
    case 1: LAUNCH(1) break;
    case 2: LAUNCH(2) break;
    case 3: LAUNCH(3) break;
    default: LAUNCH(4) break;
  }
#undef LAUNCH
#undef LAUNCH2
  ETPGT_CHECK_LAUNCH("score_topk_tc");
  return etpgt_topk_merge(cand_val, cand_idx, batch, p.splits, k, top_val, top_idx, stream_);
}
