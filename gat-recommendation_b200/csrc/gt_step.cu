// Step driver: the whole training step of graph_transformer_optimized (forward, sampled loss, backward)
// as ONE host call that launches the path's kernels back to back from C++.
//
// The reference strings these stages together in Python (Trainer.train_epoch, etpgt/train/trainer.py:69-131:
// model(batch) -> loss -> backward), one framework dispatch per operator.  The kernels of this library are
// short (a step is ~80 launches of 5-300 us), so at small and medium batches the step time is the HOST's time to
// walk autograd and marshal arguments, not the device's.  This driver calls the same entry points, in the same
// order and with the same arguments as the autograd nodes of etpgt_b200/ops.py (so its results are bit-identical
// to that path), from compiled code and out of one caller-provided arena: no Python, no autograd graph, no
// per-tensor allocation between launches.
//
// Data parallelism: BatchNorm needs whole-batch statistics, i.e. an all-reduce of 2*dim+1 doubles between the
// statistics kernel and the apply kernel of every layer, forward and backward.  The step is therefore cut into
// phases at exactly those points; the caller runs [phase_begin, phase_end) per call and all-reduces the exchanged
// rows of `bn_sums` in between (one call with all phases on a single GPU).  One more cut follows the last table-
// gradient kernel: phase 2*layers+1 (weight gradient of layer 0, PE projection gradient) touches neither the table
// gradient nor anything a collective needs, so the table's all-reduce can run underneath it.  2*layers+2 phases.
#include "common.cuh"

namespace etpgt {
namespace {

__global__ void fill_tail_kernel(double* a, double va, float* b, float vb) {
  if (threadIdx.x == 0) {
    if (a) *a = va;
    if (b) *b = vb;
  }
}
__global__ void copy_doubles_kernel(double* dst, const double* src, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}
__global__ void bump_counters_kernel(int64_t* c0, int64_t* c1, int64_t* c2, int64_t* c3) {
  if (threadIdx.x == 0) {
    if (c0) *c0 += 1;
    if (c1) *c1 += 1;
    if (c2) *c2 += 1;
    if (c3) *c3 += 1;
  }
}

size_t max_sz(size_t a, size_t b) { return a > b ? a : b; }

// Everything the step keeps between its kernels, carved from the arena in a fixed order (the same layout is
// recomputed by every phase call).
struct LayerBuffers {
  void *x_hi, *x_lo;        // bf16 [n, in]: split of the layer input (layer > 0: written by the previous BN apply)
  void *w_hi, *w_lo;        // bf16 [4*dim, in]
  float* qkvs;              // [n, 4*dim]
  float* alpha_mask;        // [e, heads] or null
  float *conv_out, *agg;    // [n, dim]
  float *beta, *m, *inv_l;  // [n], [n, heads] x2
  float *mean, *invstd;     // [dim]
  float* y;                 // [n, dim] layer output
  double* local;            // [2*dim+1] backward statistics of this rank
  float *d_conv, *d_res;    // [n, dim]
  void *g_hi, *g_lo;        // bf16 [n, 4*dim]
};

struct Layout {
  LayerBuffers layer[ETPGT_GT_MAX_LAYERS];
  float* x0;          // [n, dim]
  int32_t* seg_ptr;   // [b+1]
  void* readout_aux;  // [b, dim] int32 (max readout)
  float* scores;      // [b, 1+num_neg]
  float* d_loss;      // [1]
  float* d_sess;      // [b, dim]
  float* d_last;      // [n, dim] gradient of the last layer's output
  void* scratch;
  size_t scratch_bytes;
  size_t total;
};

size_t scratch_bytes_of(const etpgt_gt_step_t& s) {
  const int64_t n = s.num_nodes, e = s.num_edges, b = s.num_sessions;
  const int dim = s.dim, width = 4 * s.dim;
  size_t need = 256;
  need = max_sz(need, etpgt_gemm_bf16x3_workspace_bytes(n, width, dim, 1));
  need = max_sz(need, etpgt_gemm_bf16x3_workspace_bytes(n, dim, width, 1));
  need = max_sz(need, etpgt_gemm_bf16x3_workspace_bytes(width, dim, n, 0));
  need = max_sz(need, etpgt_bn_workspace_bytes(n, dim));
  need = max_sz(need, etpgt_tconv_fwd_bn_workspace_bytes(dim));
  need = max_sz(need, etpgt_tconv_bwd_workspace_bytes(n, e, dim, s.heads));
  need = max_sz(need, etpgt_sampled_loss_workspace_bytes(b, s.num_neg, dim));
  need = max_sz(need, etpgt_embed_pe_bwd_workspace_bytes(n, dim, s.k_pe > 0 ? s.k_pe : 1));
  return align_up(need);
}

Layout carve(const etpgt_gt_step_t& s, void* arena) {
  Workspace w(arena, ~size_t(0));
  Layout L{};
  const size_t n = (size_t)s.num_nodes, e = (size_t)s.num_edges, b = (size_t)s.num_sessions;
  const size_t dim = (size_t)s.dim, width = 4 * dim, heads = (size_t)s.heads;
  const bool alpha_drop = s.training && s.alpha_p > 0.0 && e > 0;
  L.x0 = w.take<float>(n * dim);
  for (int l = 0; l < s.num_layers; ++l) {
    LayerBuffers& B = L.layer[l];
    if (l == 0) {
      B.x_hi = w.take<uint16_t>(n * dim);
      B.x_lo = w.take<uint16_t>(n * dim);
    }
    B.w_hi = w.take<uint16_t>(width * dim);
    B.w_lo = w.take<uint16_t>(width * dim);
    B.qkvs = w.take<float>(n * width);
    B.alpha_mask = alpha_drop ? w.take<float>(e * heads) : nullptr;
    B.conv_out = w.take<float>(n * dim);
    B.agg = w.take<float>(n * dim);
    B.beta = w.take<float>(n);
    B.m = w.take<float>(n * heads);
    B.inv_l = w.take<float>(n * heads);
    B.mean = w.take<float>(dim);
    B.invstd = w.take<float>(dim);
    B.y = w.take<float>(n * dim);
    if (l + 1 < s.num_layers) {   // bf16 hand-over to the next layer's projection
      L.layer[l + 1].x_hi = w.take<uint16_t>(n * dim);
      L.layer[l + 1].x_lo = w.take<uint16_t>(n * dim);
    }
    B.local = w.take<double>(2 * dim + 1);
    B.d_conv = w.take<float>(n * dim);
    B.d_res = w.take<float>(n * dim);
    B.g_hi = w.take<uint16_t>(n * width);
    B.g_lo = w.take<uint16_t>(n * width);
  }
  L.seg_ptr = w.take<int32_t>(b + 1);
  L.readout_aux = s.readout_mode == ETPGT_READOUT_MAX ? (void*)w.take<int32_t>(b * dim) : nullptr;
  L.scores = w.take<float>(b * (size_t)(s.num_neg + 1));
  L.d_loss = w.take<float>(1);
  L.d_sess = w.take<float>(b * dim);
  L.d_last = w.take<float>(n * dim);
  L.scratch_bytes = scratch_bytes_of(s);
  L.scratch = w.take<char>(L.scratch_bytes);
  L.total = w.used;
  return L;
}

int check(const etpgt_gt_step_t* s) {
  ETPGT_REQUIRE(s != nullptr, "gt_step: null descriptor");
  ETPGT_REQUIRE(s->struct_bytes == sizeof(etpgt_gt_step_t), "gt_step: descriptor size %lld != %zu (header mismatch)",
                (long long)s->struct_bytes, sizeof(etpgt_gt_step_t));
  ETPGT_REQUIRE(supported_dim(s->dim), "gt_step: unsupported dim %d", s->dim);
  ETPGT_REQUIRE(s->num_layers >= 1 && s->num_layers <= ETPGT_GT_MAX_LAYERS, "gt_step: layers must be 1..%d",
                ETPGT_GT_MAX_LAYERS);
  ETPGT_REQUIRE(s->num_nodes >= 1 && s->num_edges >= 0 && s->num_sessions >= 1, "gt_step: empty batch");
  ETPGT_REQUIRE(s->readout_mode == ETPGT_READOUT_MEAN || s->readout_mode == ETPGT_READOUT_MAX ||
                    s->readout_mode == ETPGT_READOUT_LAST,
                "gt_step: readout mode %d is not driven here (mean / max / last)", s->readout_mode);
  ETPGT_REQUIRE(s->loss_mode >= ETPGT_LOSS_BPR && s->loss_mode <= ETPGT_LOSS_DUAL, "Unknown loss type: %d",
                s->loss_mode);
  ETPGT_REQUIRE(s->num_neg >= 1, "gt_step: num_neg must be >= 1");
  ETPGT_REQUIRE(s->distributed || s->num_nodes >= 2 || !s->training,
                "Expected more than 1 value per channel when training");
  return ETPGT_OK;
}

#define TRY(expr)                    \
  do {                               \
    int rc__ = (expr);               \
    if (rc__ != ETPGT_OK) return rc__; \
  } while (0)

}  // namespace
}  // namespace etpgt

using namespace etpgt;

extern "C" size_t etpgt_gt_step_arena_bytes(const etpgt_gt_step_t* s) {
  if (check(s) != ETPGT_OK) return 0;
  return carve(*s, nullptr).total + 256;
}

extern "C" int etpgt_gt_step_num_phases(const etpgt_gt_step_t* s) {
  if (check(s) != ETPGT_OK) return 0;
  return 2 * s->num_layers + 2;
}

extern "C" int etpgt_gt_step_run(const etpgt_gt_step_t* sp, int phase_begin, int phase_end, etpgt_stream_t stream_) {
  TRY(check(sp));
  const etpgt_gt_step_t& s = *sp;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int L = s.num_layers, phases = 2 * L + 2;
  ETPGT_REQUIRE(phase_begin >= 0 && phase_end <= phases && phase_begin < phase_end, "gt_step: bad phase range [%d, %d)",
                phase_begin, phase_end);
  ETPGT_REQUIRE(s.arena != nullptr && s.arena_bytes >= etpgt_gt_step_arena_bytes(sp), "gt_step: arena %zu < %zu",
                s.arena_bytes, etpgt_gt_step_arena_bytes(sp));
  ETPGT_REQUIRE((reinterpret_cast<uintptr_t>(s.arena) & 255) == 0, "gt_step: arena must be 256-byte aligned");
  ETPGT_REQUIRE(s.ids && s.batch_vec && s.rowptr && s.colptr && s.targets && s.negatives && s.table && s.bn_sums &&
                    s.sess && s.losses,
                "gt_step: null pointer in the descriptor");
  ETPGT_REQUIRE(s.num_edges == 0 || (s.col && s.eperm && s.row && s.cpos), "gt_step: null edge arrays");
  const Layout lay = carve(s, s.arena);
  const int64_t n = s.num_nodes, e = s.num_edges, b = s.num_sessions;
  const int dim = s.dim, width = 4 * s.dim, heads = s.heads;
  const int sums_len = 2 * dim + 1;
  const double drop_p = s.training ? s.drop_p : 0.0;
  auto run = [&](int p) { return p >= phase_begin && p < phase_end; };
  auto fwd_sums = [&](int l) { return s.bn_sums + (size_t)l * sums_len; };
  auto bwd_sums = [&](int l) { return s.bn_sums + (size_t)(L + l) * sums_len; };
  // count argument of the BatchNorm kernels: 0 = read the (all-reduced) row count from the tail of the sums
  const double count = s.distributed ? 0.0 : (double)n;

  // ---- forward of layer l up to its BatchNorm statistics
  auto layer_forward_a = [&](int l) -> int {
    const etpgt_gt_layer_t& P = s.layer[l];
    const LayerBuffers& B = lay.layer[l];
    const float* x = l == 0 ? lay.x0 : lay.layer[l - 1].y;
    (void)x;   // layer 0: the embedding kernel wrote the split of x0; layer > 0: the previous BatchNorm apply did
    TRY(etpgt_split_bf16(P.weight, width, dim, dim, B.w_hi, B.w_lo, dim, nullptr, nullptr, (width + 7) / 8 * 8,
                         nullptr, lay.scratch, 256, stream_));
    TRY(etpgt_gemm_bf16x3_ex(B.x_hi, B.x_lo, B.w_hi, B.w_lo, n, width, dim, dim, dim, 0, 0, P.bias, 0, B.qkvs, width,
                             1, lay.scratch, etpgt_gemm_bf16x3_workspace_bytes(n, width, dim, 1), stream_));
    if (B.alpha_mask) TRY(etpgt_dropout_mask(P.alpha_seed, s.alpha_p, e * heads, B.alpha_mask, stream_));
    if (s.training) {
      // the BatchNorm statistics of the conv output come out of the conv kernel itself (no second pass over it)
      TRY(etpgt_tconv_fwd_bn(B.qkvs, n, dim, heads, s.rowptr, s.col, s.eperm, e, P.w_beta, B.alpha_mask, B.conv_out,
                             B.agg, B.beta, B.m, B.inv_l, nullptr, nullptr, 0, fwd_sums(l), lay.scratch,
                             etpgt_tconv_fwd_bn_workspace_bytes(dim), stream_));
      if (s.distributed) {
        fill_tail_kernel<<<1, 32, 0, stream>>>(fwd_sums(l) + 2 * dim, (double)n, nullptr, 0.f);
        ETPGT_CHECK_LAUNCH("gt_step count");
        // peer memory: the exchange is one kernel of this stream instead of a phase cut + host collective
        if (s.comm) TRY(etpgt_comm_allreduce_f64(s.comm, fwd_sums(l), fwd_sums(l), sums_len, stream_));
      }
    } else {
      TRY(etpgt_tconv_fwd(B.qkvs, n, dim, heads, s.rowptr, s.col, s.eperm, e, P.w_beta, B.alpha_mask, B.conv_out,
                          B.agg, B.beta, B.m, B.inv_l, stream_));
    }
    return ETPGT_OK;
  };
  // ---- BatchNorm finalize + apply (+ residual, dropout, bf16 split for the next layer)
  auto layer_forward_b = [&](int l) -> int {
    const etpgt_gt_layer_t& P = s.layer[l];
    const LayerBuffers& B = lay.layer[l];
    const float* x = l == 0 ? lay.x0 : lay.layer[l - 1].y;
    if (s.training)
      TRY(etpgt_bn_finalize(fwd_sums(l), count, dim, (float)P.eps, (float)P.momentum, B.mean, B.invstd,
                            P.running_mean, P.running_var, stream_));
    else
      TRY(etpgt_bn_from_running(P.running_mean, P.running_var, dim, (float)P.eps, B.mean, B.invstd, stream_));
    void* y_hi = l + 1 < L ? lay.layer[l + 1].x_hi : nullptr;
    void* y_lo = l + 1 < L ? lay.layer[l + 1].x_lo : nullptr;
    TRY(etpgt_bn_apply_ex(B.conv_out, n, dim, B.mean, B.invstd, P.bn_weight, P.bn_bias, x, 0, drop_p, P.drop_seed,
                          B.y, y_hi, y_lo, stream_));
    return ETPGT_OK;
  };
  // ---- backward: BatchNorm statistics of layer l
  auto layer_backward_a = [&](int l) -> int {
    const etpgt_gt_layer_t& P = s.layer[l];
    const LayerBuffers& B = lay.layer[l];
    const float* d_y = l == L - 1 ? lay.d_last : lay.layer[l + 1].d_res;
    TRY(etpgt_bn_bwd_stats_ex(B.conv_out, nullptr, d_y, n, dim, B.mean, B.invstd, 0, drop_p, P.drop_seed, B.local,
                              lay.scratch, etpgt_bn_workspace_bytes(n, dim), stream_));
    if (s.training && s.distributed) {
      fill_tail_kernel<<<1, 32, 0, stream>>>(B.local + 2 * dim, (double)n, nullptr, 0.f);
      ETPGT_CHECK_LAUNCH("gt_step count");
      if (s.comm) {
        TRY(etpgt_comm_allreduce_f64(s.comm, B.local, bwd_sums(l), sums_len, stream_));
      } else {
        copy_doubles_kernel<<<(sums_len + 255) / 256, 256, 0, stream>>>(bwd_sums(l), B.local, sums_len);
        ETPGT_CHECK_LAUNCH("gt_step sums copy");
      }
    }
    return ETPGT_OK;
  };
  // ---- backward: the rest of layer l
  auto layer_backward_b = [&](int l) -> int {
    const etpgt_gt_layer_t& P = s.layer[l];
    const LayerBuffers& B = lay.layer[l];
    const float* d_y = l == L - 1 ? lay.d_last : lay.layer[l + 1].d_res;
    const double* sums = (s.training && s.distributed) ? bwd_sums(l) : B.local;
    TRY(etpgt_bn_bwd_apply_ex(B.conv_out, nullptr, d_y, n, dim, B.mean, B.invstd, P.bn_weight, 0, s.training, sums,
                              count, B.local, drop_p, P.drop_seed, B.d_conv, B.d_res, P.d_bn_weight, P.d_bn_bias,
                              stream_));
    TRY(etpgt_tconv_bwd_split(B.qkvs, B.d_conv, n, dim, heads, s.rowptr, s.col, s.eperm, s.colptr, s.row, s.cpos, e,
                              P.w_beta, B.alpha_mask, B.agg, B.beta, B.m, B.inv_l, nullptr, B.g_hi, B.g_lo, P.d_bias,
                              P.d_w_beta, lay.scratch, etpgt_tconv_bwd_workspace_bytes(n, e, dim, heads), stream_));
    // dX = d_res + dQKVS x W (residual branch merged by the GEMM's TMA reduce-add)
    TRY(etpgt_gemm_bf16x3_ex(B.g_hi, B.g_lo, B.w_hi, B.w_lo, n, dim, width, width, dim, 0, 1, nullptr, 1, B.d_res, dim,
                             1, lay.scratch, etpgt_gemm_bf16x3_workspace_bytes(n, dim, width, 1), stream_));
    return ETPGT_OK;
  };
  // ---- backward: weight gradient of layer l, dW = dQKVS^T x X (split-K over the nodes)
  auto layer_backward_w = [&](int l) -> int {
    const etpgt_gt_layer_t& P = s.layer[l];
    const LayerBuffers& B = lay.layer[l];
    TRY(etpgt_gemm_bf16x3_ex(B.g_hi, B.g_lo, B.x_hi, B.x_lo, width, dim, n, width, dim, 1, 1, nullptr, 0, P.d_weight,
                             dim, 0, lay.scratch, etpgt_gemm_bf16x3_workspace_bytes(width, dim, n, 0), stream_));
    return ETPGT_OK;
  };

  for (int p = phase_begin; p < phase_end; ++p) {
    if (p == 0) {
      TRY(etpgt_embed_pe_fwd_split(s.ids, n, s.table, s.num_items, s.pe, 0, s.w_pe, s.b_pe, s.pe ? s.k_pe : 0, dim,
                                   lay.x0, lay.layer[0].x_hi, lay.layer[0].x_lo, stream_));
      if (s.training) {
        static_assert(ETPGT_GT_MAX_LAYERS == 4, "bump_counters_kernel takes four counters");
        int64_t* c[4] = {nullptr, nullptr, nullptr, nullptr};
        for (int l = 0; l < L; ++l) c[l] = s.layer[l].num_batches_tracked;
        bump_counters_kernel<<<1, 32, 0, stream>>>(c[0], c[1], c[2], c[3]);
        ETPGT_CHECK_LAUNCH("gt_step counters");
      }
      TRY(layer_forward_a(0));
    } else if (p < L) {
      TRY(layer_forward_b(p - 1));
      TRY(layer_forward_a(p));
    } else if (p == L) {
      TRY(layer_forward_b(L - 1));
      const float* x_last = lay.layer[L - 1].y;
      TRY(etpgt_segment_ptr(s.batch_vec, n, b, lay.seg_ptr, stream_));
      TRY(etpgt_readout_fwd(x_last, lay.seg_ptr, b, dim, s.readout_mode, nullptr, s.sess, lay.readout_aux, stream_));
      TRY(etpgt_sampled_loss_fwd(s.sess, s.table, s.targets, s.negatives, b, s.num_neg, dim, s.loss_mode, s.alpha,
                                 s.temperature, s.total_sessions, lay.scores, s.losses, lay.scratch,
                                 etpgt_sampled_loss_workspace_bytes(b, s.num_neg, dim), stream_));
      if (s.backward) {
        fill_tail_kernel<<<1, 32, 0, stream>>>(nullptr, 0.0, lay.d_loss, 1.0f);
        ETPGT_CHECK_LAUNCH("gt_step d_loss");
        TRY(etpgt_sampled_loss_bwd_planned(s.sess, s.table, s.targets, s.negatives, b, s.num_neg, dim, s.loss_mode,
                                           s.alpha, s.temperature, s.total_sessions, lay.scores, lay.d_loss,
                                           s.num_items, s.padding_idx, s.plan_loss_key, s.plan_loss_perm, lay.d_sess,
                                           s.d_table, lay.scratch,
                                           etpgt_sampled_loss_workspace_bytes(b, s.num_neg, dim), stream_));
        TRY(etpgt_readout_bwd(x_last, s.sess, lay.d_sess, lay.seg_ptr, n, b, dim, s.readout_mode, lay.readout_aux,
                              lay.d_last, nullptr, stream_));
        TRY(layer_backward_a(L - 1));
      }
    } else if (s.backward && p <= 2 * L) {
      const int l = 2 * L - p;   // p = L+1 .. 2L  ->  l = L-1 .. 0
      TRY(layer_backward_b(l));
      if (l > 0) {
        TRY(layer_backward_w(l));
        TRY(layer_backward_a(l - 1));
      } else {
        // the table gradient is complete as early as possible: its rows need dX of layer 0 only, so the
        // scatter runs before that layer's weight gradient and the caller can start the table's all-reduce
        // while the last phase computes
        TRY(etpgt_embed_pe_bwd_planned(s.ids, n, lay.layer[0].d_res, s.num_items, nullptr, 0, 0, dim, s.padding_idx,
                                       s.plan_nodes_key, s.plan_nodes_perm, s.d_table, nullptr, nullptr, lay.scratch,
                                       etpgt_embed_pe_bwd_workspace_bytes(n, dim, s.k_pe > 0 ? s.k_pe : 1), stream_));
      }
    } else if (s.backward) {   // p == 2L+1: what no collective waits for
      TRY(layer_backward_w(0));
      if (s.pe != nullptr)
        TRY(etpgt_embed_pe_bwd_planned(s.ids, n, lay.layer[0].d_res, s.num_items, s.pe, 0, s.k_pe, dim, s.padding_idx,
                                       nullptr, nullptr, nullptr, s.d_w_pe, s.d_b_pe, lay.scratch,
                                       etpgt_embed_pe_bwd_workspace_bytes(n, dim, s.k_pe), stream_));
    }
  }
  return ETPGT_OK;
}

// ---- the same step as ONE CUDA graph launch -------------------------------------------------------------------
// At small batches (the reference trains at 32 sessions per step, params.yaml:6) every kernel of the step runs for a
// few microseconds and the step time is the ~55 launch gaps, not the work.  The driver's launches are therefore
// captured from the caller's stream (relaxed mode: nothing else of the process is disturbed), the captured graph
// UPDATES the executable graph kept in `cache` (same topology step after step; only kernel arguments, grids and — for
// the density-dependent kernel variants — functions change, which cudaGraphExecUpdate applies in place), and the
// step runs as one graph launch: dependent kernels start back to back without a launch round trip each.  When the
// topology did change (another model, dropout switched on) the executable graph is rebuilt.  Same kernels, same
// arguments, same order: results are bit-identical to etpgt_gt_step_run.
struct etpgt_graph {
  cudaGraphExec_t exec = nullptr;
  long long launches = 0, rebuilds = 0;
};

extern "C" int etpgt_graph_create(etpgt_graph_t** out) {
  ETPGT_REQUIRE(out != nullptr, "graph_create: null output");
  *out = new etpgt_graph();
  return ETPGT_OK;
}

extern "C" int etpgt_graph_destroy(etpgt_graph_t* g) {
  if (g == nullptr) return ETPGT_OK;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  (void)cudaGetLastError();
  delete g;
  return ETPGT_OK;
}

extern "C" int64_t etpgt_graph_rebuilds(const etpgt_graph_t* g) { return g ? g->rebuilds : 0; }

extern "C" int etpgt_gt_step_run_graph(const etpgt_gt_step_t* sp, int phase_begin, int phase_end, etpgt_graph_t* cache,
                                       etpgt_stream_t stream_) {
  ETPGT_REQUIRE(cache != nullptr, "gt_step_run_graph: null graph cache");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ETPGT_REQUIRE(stream != nullptr, "gt_step_run_graph: the legacy default stream cannot be captured; run on a stream");
  TRY(check(sp));   // argument errors before a capture is open
  cudaError_t e = cudaStreamBeginCapture(stream, cudaStreamCaptureModeRelaxed);
  if (e != cudaSuccess) {
    set_error("gt_step_run_graph: cudaStreamBeginCapture: %s", cudaGetErrorString(e));
    return ETPGT_ECUDA;
  }
  const int rc = etpgt_gt_step_run(sp, phase_begin, phase_end, stream_);
  cudaGraph_t graph = nullptr;
  e = cudaStreamEndCapture(stream, &graph);
  if (rc != ETPGT_OK) {
    if (graph) cudaGraphDestroy(graph);
    (void)cudaGetLastError();
    return rc;
  }
  if (e != cudaSuccess || graph == nullptr) {
    set_error("gt_step_run_graph: cudaStreamEndCapture: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return ETPGT_ECUDA;
  }
  bool ready = false;
  if (cache->exec != nullptr) {
    cudaGraphExecUpdateResultInfo info;
    if (cudaGraphExecUpdate(cache->exec, graph, &info) == cudaSuccess) {
      ready = true;
    } else {
      (void)cudaGetLastError();
      cudaGraphExecDestroy(cache->exec);
      cache->exec = nullptr;
    }
  }
  if (!ready) {
    e = cudaGraphInstantiate(&cache->exec, graph, 0);
    if (e != cudaSuccess) {
      set_error("gt_step_run_graph: cudaGraphInstantiate: %s", cudaGetErrorString(e));
      cudaGraphDestroy(graph);
      cache->exec = nullptr;
      (void)cudaGetLastError();
      return ETPGT_ECUDA;
    }
    ++cache->rebuilds;
  }
  cudaGraphDestroy(graph);
  e = cudaGraphLaunch(cache->exec, stream);
  if (e != cudaSuccess) {
    set_error("gt_step_run_graph: cudaGraphLaunch: %s", cudaGetErrorString(e));
    return ETPGT_ECUDA;
  }
  ++cache->launches;
  return ETPGT_OK;
}
