// a9: full-catalogue scoring on the 5th-generation tensor cores with a fused top-k epilogue.
//   scores[b, i] = <S[b, :], E[i, :]>  (bf16 operands, fp32 accumulation in TMEM), top-k per row
// etpgt/model/base.py:59-78 (`torch.matmul(S, E.t())` + `torch.topk`).  The [B, I] score matrix is
// never written.
//
// GEMM kernel (one CTA = 128 sessions x a contiguous range of 256-item tiles):
//   warp 0      TMA producer: session tile once (DIM/64 k-blocks of 128x64 bf16, SWIZZLE_128B), then
//               a ring of item k-blocks (256x64 bf16) with mbarrier full/empty hand-shakes
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128x256x16, kind::f16),
//               accumulators double-buffered in the 512 TMEM columns
//   warps 2-5   epilogue, one warp per TMEM lane quarter (lane = session row)
//
// Epilogue = "chunk dump".  Measured on B200: with an epilogue that only loads TMEM and takes
// maxima the pipeline runs at 84 % of the cuBLAS bf16 peak; every form of per-item candidate
// handling inside the loop (shared-memory k-lists: 3 %, pending buffers: 23 %, register lists with
// warp-uniform scans: 30 %) was bound by SIMT divergence — the 32 rows of a warp accept candidates
// at different columns, so each accepted item costs the whole warp an insert.  So the loop keeps,
// per row, only the K best 32-column CHUNK MAXIMA in a register-resident sorting network (values
// only, branch-free, all lanes in lockstep), and a lane whose chunk maximum beats its row's
// threshold dumps that chunk's 32 raw scores (one full 128-byte line) to a per-(row, range) slot
// buffer in HBM.  K chunk maxima >= thr prove K items >= thr, so thr is a valid lower bound of the
// row's K-th best score and no chunk holding a top-K item is ever skipped (chunks arrive in
// ascending id order and the test is a strict '>', so ties keep the lower id).  About
// K*ln(chunks/K) ~ 100 chunks (12 KB) per row are dumped.
//
// Select kernel (one warp per row): filters the dumped scores against the best range threshold and
// picks the exact top-k by (score desc, id asc).  Rows whose slot buffer overflowed (adversarial
// score orders) are recomputed exactly by a CUDA-core fallback kernel, so the result is always exact.
#include <math.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace etpgt {
namespace {

using namespace tc;

constexpr int BLOCK_M = 128;   // sessions per CTA (TMEM lanes)
constexpr int BLOCK_N = 256;   // items per accumulator tile (TMEM columns)
constexpr int CHUNK = 32;      // columns per tcgen05.ld
constexpr int DUMPW = 16;      // columns per dumped piece (64 bytes)
constexpr int kMaxStages = 3;  // ring of item k-blocks
constexpr int kAccStages = 2;  // TMEM accumulator double buffer
constexpr int kTmemCols = 512;
constexpr int kEpilogueWarps = 4;
constexpr int kThreads = 32 * (2 + kEpilogueWarps);
constexpr int kMaxKTc = 32;
constexpr int kPendingMerge = 10;  // capacity: maxima a row may have waiting (merge threshold - 1 + pieces per chunk)
constexpr int kDefaultPendingMerge = 8;
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

// Dump buffers, per (row, range): `cap` slots of 32 scores + the chunk's first column (relative to the
// range), a slot count (may exceed cap = overflow) and the range's final threshold.
struct DumpBuffers {
  float* scores;      // [rows*splits][cap][32]
  int32_t* chunk_col; // [rows*splits][cap]
  int32_t* count;     // [rows*splits]
  float* threshold;   // [rows*splits]
  int cap;
};

// Work units.  One CTA per SM and one long range per row is the cheapest (every range pays a warm-up
// while its thresholds rise from -inf), but the number of 128-row tiles is rarely a multiple of 148.
// So the first `m_full` row tiles (whole waves) each scan the WHOLE catalogue as one unit, and only
// the remaining row tiles are cut into `tail_splits` item ranges to fill the last wave.
struct Schedule {
  int m_full;           // row tiles with a single full-catalogue unit (a multiple of 148, may be 0)
  int tail_splits;      // item ranges per remaining row tile
  int tiles_per_split;  // 256-item tiles per tail range
  int total_tiles;
  __host__ __device__ int64_t full_rows() const { return (int64_t)m_full * BLOCK_M; }
  __host__ __device__ int parts_of_row(int64_t row) const { return row < full_rows() ? 1 : tail_splits; }
  // index of (row, part) in the count / threshold arrays; x cap in the slot arrays
  __host__ __device__ int64_t part_index(int64_t row, int part) const {
    return row < full_rows() ? row : full_rows() + (row - full_rows()) * tail_splits + part;
  }
  __host__ __device__ int first_tile(int64_t row, int part) const { return row < full_rows() ? 0 : part * tiles_per_split; }
  __host__ __device__ int64_t num_parts(int64_t batch) const {
    const int64_t f = batch < full_rows() ? batch : full_rows();
    return f + (batch - f) * tail_splits;
  }
};

template <int NUM_KB, int KCAP>  // DIM / 64, number of chunk maxima tracked (>= k)
__global__ void __launch_bounds__(kThreads, 1)
score_dump_tc_kernel(const __grid_constant__ CUtensorMap map_sess, const __grid_constant__ CUtensorMap map_items,
                     int64_t batch, int64_t num_items, int stages, Schedule sch, DumpBuffers dump,
                     int pending_merge /* <= kPendingMerge */) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;                                   // NUM_KB x 16 KB
  uint8_t* smem_b = smem_a + NUM_KB * A_KBLOCK_BYTES;       // stages x 32 KB
  float* pend_max = reinterpret_cast<float*>(smem_b + stages * B_STAGE_BYTES);  // [kPendingMerge][128]
  Barriers* bars = reinterpret_cast<Barriers*>(pend_max + kPendingMerge * BLOCK_M);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int m_tile, part;
  int64_t tile_begin, tile_end;
  if ((int)blockIdx.x < sch.m_full) {
    m_tile = blockIdx.x; part = 0; tile_begin = 0; tile_end = sch.total_tiles;
  } else {
    const int v = blockIdx.x - sch.m_full;
    m_tile = sch.m_full + v / sch.tail_splits;
    part = v % sch.tail_splits;
    tile_begin = (int64_t)part * sch.tiles_per_split;
    tile_end = tile_begin + sch.tiles_per_split < sch.total_tiles ? tile_begin + sch.tiles_per_split : sch.total_tiles;
  }
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
    const int64_t grow = (int64_t)m_tile * BLOCK_M + row;
    const bool live = grow < batch;
    const int64_t pidx = live ? sch.part_index(grow, part) : 0;
    const int64_t slot0 = pidx * (int64_t)dump.cap;
    float* my_scores = dump.scores + slot0 * DUMPW;
    int32_t* my_cols = dump.chunk_col + slot0;
    float best[KCAP];   // the KCAP largest chunk maxima of this row so far, descending
#pragma unroll
    for (int t = 0; t < KCAP; ++t) best[t] = -INFINITY;
    float thr = -INFINITY;
    int count = 0, pending = 0;
    const uint32_t my_pending = smem_u32(pend_max + row);  // [kPendingMerge][128], row fastest (shared-space address)
    for (int t = 0; t < num_tiles; ++t) {
      const int acc = t & 1;
      const uint32_t acc_phase = (uint32_t)(t >> 1) & 1;
      mbar_wait(&bars->acc_full[acc], acc_phase);
      tc_fence_after();
      const int64_t item0 = (tile_begin + t) * BLOCK_N;
      const int limit = num_items - item0 < BLOCK_N ? (int)(num_items - item0) : BLOCK_N;  // real columns
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BLOCK_N);
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += CHUNK) {
        float v[CHUNK];
        {
          uint32_t raw[CHUNK];
          tmem_ld_32x32(taddr + (uint32_t)c0, raw);
#pragma unroll
          for (int j = 0; j < CHUNK; ++j) v[j] = __uint_as_float(raw[j]);
        }
        if (limit - c0 < CHUNK) {  // only the table's last tile: columns past the end never qualify
#pragma unroll
          for (int j = 0; j < CHUNK; ++j) v[j] = c0 + j < limit ? v[j] : -INFINITY;
        }
        float g[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) g[q] = fmaxf(fmaxf(v[4 * q], v[4 * q + 1]), fmaxf(v[4 * q + 2], v[4 * q + 3]));
        // dump granularity = DUMPW (16) columns: one 64-byte piece per hit instead of the whole 128-byte
        // chunk — a third less dump traffic and half the store instructions on the divergent path
        float mx[CHUNK / DUMPW];
        bool hit[CHUNK / DUMPW];
        bool any_hit = false;
#pragma unroll
        for (int h = 0; h < CHUNK / DUMPW; ++h) {
          mx[h] = fmaxf(fmaxf(g[4 * h], g[4 * h + 1]), fmaxf(g[4 * h + 2], g[4 * h + 3]));
          hit[h] = live && mx[h] > thr;
          any_hit = any_hit || hit[h];
        }
        if (__any_sync(0xffffffffu, any_hit)) {  // warp-uniform
#pragma unroll
          for (int h = 0; h < CHUNK / DUMPW; ++h) {
            if (hit[h]) {
              const int slot = count++;
              st_shared_f32(my_pending + pending * BLOCK_M * sizeof(float), mx[h]);  // waits for the lockstep merge
              ++pending;
              if (slot < dump.cap) {
                my_cols[slot] = t * BLOCK_N + c0 + h * DUMPW;
                float4* dst = reinterpret_cast<float4*>(my_scores + (int64_t)slot * DUMPW);
#pragma unroll
                for (int q = 0; q < DUMPW / 4; ++q) {
                  const int j = h * DUMPW + 4 * q;
                  __stcs(dst + q, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                }
              }
            }
          }
          // Merge the pending chunk maxima of all 32 rows into the sorted lists TOGETHER: one pass of
          // the sorting network then serves up to 32 rows at once (run per hit it would serve ~1).  The
          // threshold is a little stale in between, which only dumps a few extra chunks.
          if (__any_sync(0xffffffffu, pending >= pending_merge)) {
            int most = pending;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) most = max(most, __shfl_xor_sync(0xffffffffu, most, off));
            for (int p = 0; p < most; ++p) {
              float x = p < pending ? ld_shared_f32(my_pending + p * BLOCK_M * sizeof(float)) : -INFINITY;  // -inf: no-op
#pragma unroll
              for (int s = 0; s < KCAP; ++s) {
                const float hi = fmaxf(best[s], x);
                x = fminf(best[s], x);
                best[s] = hi;
              }
            }
            pending = 0;
            thr = best[KCAP - 1];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc_empty[acc]);
    }
    if (live) {
      dump.count[pidx] = count;
      dump.threshold[pidx] = thr;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

__device__ __forceinline__ bool better(float v, int64_t i, float bv, int64_t bi) {
  return v > bv || (v == bv && i < bi);
}

// (score, column) packed so that an unsigned 64-bit comparison orders candidates by score descending,
// then column ascending: high word = order-preserving image of the float (-0.0 canonicalised to +0.0,
// like the oracle's comparison), low word = ~column.  0 is never a valid key.
__device__ __forceinline__ uint64_t candidate_key(float v, int32_t col) {
  uint32_t f = __float_as_uint(v + 0.0f);
  f = (f & 0x80000000u) ? ~f : (f | 0x80000000u);
  return ((uint64_t)f << 32) | (uint64_t)(0xFFFFFFFFu - (uint32_t)col);
}
__device__ __forceinline__ void decode_key(uint64_t key, float& v, int64_t& col) {
  if (key == 0) { v = -INFINITY; col = 0; return; }
  const uint32_t f = (uint32_t)(key >> 32);
  v = __uint_as_float((f & 0x80000000u) ? (f ^ 0x80000000u) : ~f);
  col = (int64_t)(0xFFFFFFFFu - (uint32_t)key);
}

// One warp per row.  Survivors (scores >= the best range threshold) are collected in shared memory as
// packed keys, then sorted (bitonic network over 64 keys; selection passes when there are more).
constexpr int kSelectWarps = 4;
constexpr int kSurvivorCap = 512;

__global__ void __launch_bounds__(kSelectWarps * 32)
score_select_kernel(DumpBuffers dump, int64_t batch, Schedule sch, int k, int64_t id_base,
                    float* __restrict__ top_val, int64_t* __restrict__ top_idx, int32_t* __restrict__ redo,
                    const int64_t* __restrict__ targets, int32_t* __restrict__ hit_pos) {
  __shared__ uint64_t s_key[kSelectWarps][kSurvivorCap];  // (orderable score, ~column): larger = better
  __shared__ int s_count[kSelectWarps];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)kSelectWarps + w;
  if (row >= batch) return;
  if (lane == 0) s_count[w] = 0;
  float tau = -INFINITY;
  bool overflow = false;
  const int splits = sch.parts_of_row(row);
  for (int s = 0; s < splits; ++s) {
    tau = fmaxf(tau, dump.threshold[sch.part_index(row, s)]);
    overflow = overflow || dump.count[sch.part_index(row, s)] > dump.cap;
  }
  __syncwarp();
  if (!overflow) {
    for (int s = 0; s < splits; ++s) {
      const int n = dump.count[sch.part_index(row, s)];
      const int64_t slot0 = sch.part_index(row, s) * (int64_t)dump.cap;
      const int range_col0 = sch.first_tile(row, s) * BLOCK_N;
      // 128-bit loads: DUMPW/4 lanes cover one dumped piece, a warp load covers 128/DUMPW pieces, eight
      // independent loads in flight per lane (4 KB per warp and iteration)
      constexpr int LPP = DUMPW / 4;            // lanes per piece
      constexpr int PPL = 32 / LPP;             // pieces per warp load
      const int sub = lane / LPP, quad = lane % LPP;
      const float4* base = reinterpret_cast<const float4*>(dump.scores + slot0 * DUMPW) + quad;
      for (int c0 = 0; c0 < n; c0 += 8 * PPL) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int piece = c0 + PPL * u + sub;
          v[u] = piece < n ? __ldcs(base + (int64_t)piece * LPP) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float e4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
          for (int comp = 0; comp < 4; ++comp) {
            if (e4[comp] >= tau && e4[comp] > -INFINITY) {  // -inf marks columns past the end of the table
              const int pos = atomicAdd(&s_count[w], 1);
              if (pos < kSurvivorCap) {
                s_key[w][pos] = candidate_key(
                    e4[comp], range_col0 + dump.chunk_col[slot0 + c0 + PPL * u + sub] + 4 * quad + comp);
              }
            }
          }
        }
      }
    }
    __syncwarp();
    overflow = s_count[w] > kSurvivorCap;  // a huge tie at the threshold
  }
  if (overflow) {
    if (lane == 0) redo[row] = 1;
    return;
  }
  if (lane == 0) redo[row] = 0;
  const int n = s_count[w];
  uint64_t mine = 0;   // lane t ends up with the t-th best candidate (0 = none)
  if (n <= 64) {
    // the usual case (a few dozen survivors): bitonic sort of 64 keys, two per lane, descending
    uint64_t k0 = lane < n ? s_key[w][lane] : 0, k1 = lane + 32 < n ? s_key[w][lane + 32] : 0;
#pragma unroll
    for (int size = 2; size <= 64; size <<= 1) {
#pragma unroll
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        if (stride == 32) {          // partner of element i (< 32) is i + 32: the lane's own second key
          const uint64_t hi = k0 > k1 ? k0 : k1, lo = k0 > k1 ? k1 : k0;   // size == 64: descending everywhere
          k0 = hi; k1 = lo;
        } else {
          // element index e0 = lane, e1 = lane + 32; direction of the bitonic block containing the element
          const uint64_t p0 = __shfl_xor_sync(0xffffffffu, k0, stride), p1 = __shfl_xor_sync(0xffffffffu, k1, stride);
          const bool lower = (lane & stride) == 0;                       // this lane holds the lower index of the pair
          const bool desc0 = size == 64 || ((lane & size) == 0);        // block direction (final merge: descending)
          const bool desc1 = size == 64 || (((lane + 32) & size) == 0);
          const bool take_max0 = lower == desc0, take_max1 = lower == desc1;
          k0 = take_max0 ? (k0 > p0 ? k0 : p0) : (k0 < p0 ? k0 : p0);
          k1 = take_max1 ? (k1 > p1 ? k1 : p1) : (k1 < p1 ? k1 : p1);
        }
      }
    }
    mine = k0;                       // ranks 0..31 are the first keys of lanes 0..31 (k <= 32)
  } else {
    // many survivors (large ties at the threshold): k selection passes over the shared-memory list
    uint64_t prev = ~uint64_t(0);
    for (int t = 0; t < k; ++t) {
      uint64_t best = 0;
      for (int c = lane; c < n; c += 32) {
        const uint64_t key = s_key[w][c];
        if (key < prev && key > best) best = key;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const uint64_t o = __shfl_xor_sync(0xffffffffu, best, off);
        best = o > best ? o : best;
      }
      if (lane == t) mine = best;
      prev = best;
    }
  }
  float v;
  int64_t col;
  decode_key(mine, v, col);
  const int64_t id = mine == 0 ? INT64_MAX : id_base + col;
  if (lane < k) {   // one coalesced store of the row's k results
    top_val[row * k + lane] = v;
    top_idx[row * k + lane] = id;
  }
  if (targets != nullptr) {
    // Recall / NDCG input (etpgt/utils/metrics.py:6-66) straight from the registers that hold the sorted list:
    // position of the row's target among its k results (first match), or -1
    const unsigned found = __ballot_sync(0xffffffffu, lane < k && id == targets[row]);
    if (lane == 0) hit_pos[row] = found ? __ffs(found) - 1 : -1;
  }
}

// Exact CUDA-core recomputation of the rows flagged in `redo` (slot-buffer or tie overflow): one CTA
// per row, every thread scores a strided share of the items (fp32 accumulation of the same bf16
// operands) and keeps a private sorted k-list; lists are merged through shared memory.
constexpr int kRedoThreads = 256;

__global__ void __launch_bounds__(kRedoThreads)
score_redo_kernel(const __nv_bfloat16* __restrict__ sess, const __nv_bfloat16* __restrict__ table, int64_t batch,
                  int64_t num_items, int dim, int k, int64_t id_base, const int32_t* __restrict__ redo,
                  float* __restrict__ top_val, int64_t* __restrict__ top_idx, const int64_t* __restrict__ targets,
                  int32_t* __restrict__ hit_pos) {
  extern __shared__ float redo_smem[];       // session row [dim], then lists
  // a small persistent grid; the flags of kRedoThreads rows are read at once, so rows that need no
  // recomputation (normally all of them) cost one coalesced load per 256 rows
  __shared__ int32_t s_flag[kRedoThreads];
  for (int64_t row0 = (int64_t)blockIdx.x * kRedoThreads; row0 < batch; row0 += (int64_t)gridDim.x * kRedoThreads) {
  const int32_t mine = row0 + threadIdx.x < batch ? redo[row0 + threadIdx.x] : 0;
  if (__syncthreads_or(mine) == 0) continue;  // CTA-uniform
  s_flag[threadIdx.x] = mine;
  __syncthreads();
  for (int rr = 0; rr < kRedoThreads; ++rr) {
  if (s_flag[rr] == 0) continue;              // CTA-uniform
  const int64_t row = row0 + rr;
  __syncthreads();                           // the previous row's lists are no longer read
  float* s_row = redo_smem;
  float* l_val = redo_smem + dim;            // [kRedoThreads][k]
  int32_t* l_idx = reinterpret_cast<int32_t*>(l_val + kRedoThreads * k);
  for (int d = threadIdx.x; d < dim; d += kRedoThreads) s_row[d] = __bfloat162float(sess[row * dim + d]);
  float* mv = l_val + threadIdx.x * k;
  int32_t* mi = l_idx + threadIdx.x * k;
  for (int t = 0; t < k; ++t) { mv[t] = -INFINITY; mi[t] = INT32_MAX; }
  __syncthreads();
  for (int64_t item = threadIdx.x; item < num_items; item += kRedoThreads) {
    const __nv_bfloat16* e = table + item * dim;
    float acc = 0.f;
    for (int d = 0; d < dim; d += 2) {
      const float2 p = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(e + d));
      acc = fmaf(s_row[d], p.x, acc);
      acc = fmaf(s_row[d + 1], p.y, acc);
    }
    if (acc > mv[k - 1]) {  // ascending ids per thread: strict '>' keeps the lower id
      int t = k - 1;
      while (t > 0 && mv[t - 1] < acc) { mv[t] = mv[t - 1]; mi[t] = mi[t - 1]; --t; }
      mv[t] = acc;
      mi[t] = (int32_t)item;
    }
  }
  __syncthreads();
  // k selection passes over the kRedoThreads*k candidates by (score desc, id asc)
  __shared__ float r_val[kRedoThreads / 32];
  __shared__ int32_t r_idx[kRedoThreads / 32];
  float prev_v = INFINITY;
  int32_t prev_i = -1;
  int found = -1;
  for (int t = 0; t < k; ++t) {
    float bv = -INFINITY;
    int32_t bi = INT32_MAX;
    for (int c = threadIdx.x; c < kRedoThreads * k; c += kRedoThreads) {
      const float v = l_val[c];
      const int32_t i = l_idx[c];
      const bool after_prev = v < prev_v || (v == prev_v && i > prev_i);
      if (after_prev && (v > bv || (v == bv && i < bi))) { bv = v; bi = i; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
      const int32_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { r_val[threadIdx.x >> 5] = bv; r_idx[threadIdx.x >> 5] = bi; }
    __syncthreads();
    bv = r_val[0]; bi = r_idx[0];
    for (int q = 1; q < kRedoThreads / 32; ++q)
      if (r_val[q] > bv || (r_val[q] == bv && r_idx[q] < bi)) { bv = r_val[q]; bi = r_idx[q]; }
    if (threadIdx.x == 0) {
      top_val[row * k + t] = bv;
      top_idx[row * k + t] = bi == INT32_MAX ? INT64_MAX : id_base + bi;
      if (targets != nullptr && found < 0 && bi != INT32_MAX && id_base + bi == targets[row]) found = t;
    }
    prev_v = bv;
    prev_i = bi;
    __syncthreads();
  }
  if (threadIdx.x == 0 && targets != nullptr) hit_pos[row] = found;
  }
  __syncthreads();                           // s_flag is rewritten by the next block of rows
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
  Schedule sch;
  int grid;
  int cap;
};

TcPlan tc_plan(int64_t batch, int64_t num_items, int k) {
  TcPlan p;
  const int m_tiles = (int)((batch + BLOCK_M - 1) / BLOCK_M);
  const int total_tiles = (int)((num_items + BLOCK_N - 1) / BLOCK_N);
  // tail ranges stay >= 64 tiles (512 chunks) so that every range's own threshold is tight enough for
  // the select kernel's survivor buffer
  int max_splits = total_tiles / 64 > 1 ? total_tiles / 64 : 1;
  if (max_splits > 32) max_splits = 32;
  const int m_full = (m_tiles / kNumSMs) * kNumSMs;   // whole waves: one full-catalogue unit per row tile
  const int tail = m_tiles - m_full;
  int splits = 1;
  if (tail > 0) {
    // cost model fitted on B200: time ~ waves * (range length + a warm-up worth ~half a full range)
    double best_cost = 1e30;
    for (int s = 1; s <= max_splits; ++s) {
      const int per = (total_tiles + s - 1) / s;
      const int real = (total_tiles + per - 1) / per;
      const int waves = (real * tail + kNumSMs - 1) / kNumSMs;
      const double cost = (double)waves * ((double)per / (double)total_tiles + 0.5);
      if (cost < best_cost - 1e-9) { best_cost = cost; splits = s; }
    }
    if (const char* forced = getenv("ETPGT_SCORE_SPLITS")) {  // tuning knob
      const int f = atoi(forced);
      if (f >= 1 && f <= max_splits) splits = f;
    }
  }
  p.sch.m_full = m_full;
  p.sch.total_tiles = total_tiles;
  p.sch.tiles_per_split = (total_tiles + splits - 1) / splits;
  p.sch.tail_splits = (total_tiles + p.sch.tiles_per_split - 1) / p.sch.tiles_per_split;
  p.grid = m_full + tail * p.sch.tail_splits;
  // slots per (row, range): ~ K*(1 + ln(chunks/K)) chunks are expected; 1.6x head-room (sized for the
  // longest range), overflow is handled exactly by the fallback kernel
  const double chunks = (double)(m_full > 0 ? total_tiles : p.sch.tiles_per_split) * (BLOCK_N / DUMPW);
  const int kc = k <= 10 ? 10 : k <= 20 ? 20 : 32;
  double expect = kc * (1.0 + (chunks > kc ? log(chunks / kc) : 0.0));
  int cap = (int)(1.6 * expect) + 8;
  if (cap > (int)chunks) cap = (int)chunks;
  if (const char* forced = getenv("ETPGT_SCORE_CAP")) {  // test hook: force slot-buffer overflow
    const int f = atoi(forced);
    if (f >= 1) cap = f;
  }
  p.cap = cap < 1 ? 1 : cap;
  return p;
}

size_t tc_smem_bytes(int num_kb, int stages) {
  return 1024 + (size_t)num_kb * A_KBLOCK_BYTES + (size_t)stages * B_STAGE_BYTES +
         (size_t)kPendingMerge * BLOCK_M * sizeof(float) + sizeof(Barriers) + 64;
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
  const TcPlan p = tc_plan(batch, num_items, k);
  const size_t units = (size_t)p.sch.num_parts(batch);
  return align_up(units * p.cap * DUMPW * sizeof(float)) + align_up(units * p.cap * sizeof(int32_t)) +
         2 * align_up(units * sizeof(float)) + 2 * align_up((size_t)batch * sizeof(int32_t)) + 256;
}

extern "C" int etpgt_score_topk_bf16_eval(const void* sess_bf16, const void* table_bf16, int64_t batch,
                                          int64_t num_items, int dim, int k, int64_t id_base, float* top_val,
                                          int64_t* top_idx, const int64_t* targets, int32_t* hit_pos, void* ws,
                                          size_t ws_bytes, etpgt_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ETPGT_REQUIRE((targets == nullptr) == (hit_pos == nullptr), "score_topk_bf16: targets and hit_pos come together");
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
  const TcPlan p = tc_plan(batch, num_items, k);
  Workspace w(ws, ws_bytes);
  const size_t units = (size_t)p.sch.num_parts(batch);
  DumpBuffers dump;
  dump.scores = w.take<float>(units * p.cap * DUMPW);
  dump.chunk_col = w.take<int32_t>(units * p.cap);
  dump.count = w.take<int32_t>(units);
  dump.threshold = w.take<float>(units);
  dump.cap = p.cap;
  int32_t* redo = w.take<int32_t>(batch);
  CUtensorMap map_sess, map_items;
  if (!make_map_bf16(&map_sess, sess_bf16, batch, dim, dim, BLOCK_M) ||
      !make_map_bf16(&map_items, table_bf16, num_items, dim, dim, BLOCK_N)) {
    set_error("score_topk_bf16: cuTensorMapEncodeTiled failed");
    return ETPGT_ECUDA;
  }
  const int num_kb = dim / BLOCK_K;
  const int stages = kMaxStages;
  int pending_merge = kDefaultPendingMerge;
  if (const char* forced = getenv("ETPGT_SCORE_PENDING")) {  // tuning knob
    const int f = atoi(forced);
    if (f >= 1 && f <= kPendingMerge + 1 - CHUNK / DUMPW) pending_merge = f;
  }
  const size_t smem = tc_smem_bytes(num_kb, stages);
  const dim3 grid(p.grid);
#define LAUNCH2(NKB, KC)                                                                                        \
  {                                                                                                             \
    cudaFuncSetAttribute(score_dump_tc_kernel<NKB, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    score_dump_tc_kernel<NKB, KC><<<grid, kThreads, smem, stream>>>(map_sess, map_items, batch, num_items, stages, \
                                                                   p.sch, dump, pending_merge);                \
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
  ETPGT_CHECK_LAUNCH("score_dump_tc");
  score_select_kernel<<<(unsigned)((batch + kSelectWarps - 1) / kSelectWarps), kSelectWarps * 32, 0, stream>>>(
      dump, batch, p.sch, k, id_base, top_val, top_idx, redo, targets, hit_pos);
  ETPGT_CHECK_LAUNCH("score_select");
  const size_t redo_smem = ((size_t)dim + (size_t)kRedoThreads * k * 2) * sizeof(float);
  if (redo_smem > 48 * 1024)
    cudaFuncSetAttribute(score_redo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)redo_smem);
  const int64_t redo_blocks = (batch + kRedoThreads - 1) / kRedoThreads;
  const unsigned redo_grid = (unsigned)(redo_blocks < 2 * kNumSMs ? redo_blocks : 2 * kNumSMs);
  score_redo_kernel<<<redo_grid, kRedoThreads, redo_smem, stream>>>(
      static_cast<const __nv_bfloat16*>(sess_bf16), static_cast<const __nv_bfloat16*>(table_bf16), batch, num_items,
      dim, k, id_base, redo, top_val, top_idx, targets, hit_pos);
  ETPGT_CHECK_LAUNCH("score_redo");
  return ETPGT_OK;
}

extern "C" int etpgt_score_topk_bf16(const void* sess_bf16, const void* table_bf16, int64_t batch,
                                     int64_t num_items, int dim, int k, int64_t id_base, float* top_val,
                                     int64_t* top_idx, void* ws, size_t ws_bytes, etpgt_stream_t stream) {
  return etpgt_score_topk_bf16_eval(sess_bf16, table_bf16, batch, num_items, dim, k, id_base, top_val, top_idx, nullptr,
                                    nullptr, ws, ws_bytes, stream);
}
