/*
 * twotower_b200 — C ABI of the B200-native (sm_100a) two-tower training step.
 *
 * Drop-in boundary for the hot path of zoongahn/jodalroB-twoTower
 * (TwoTowerTrainTask.forward -> loss.backward() -> Adam.step()).  The reference is
 * pure Python/PyTorch and has no FFI of its own, so every entry point below names the reference
 * function (file:line, relative to the reference root) whose arithmetic it replaces; the
 * reference-side binding is a ctypes stub.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.  All data pointers are DEVICE pointers
 *     (tensor.data_ptr()); the caller owns every buffer; the library never allocates, frees or
 *     retains a caller pointer past the call.  Scratch space is passed in (query *_workspace_bytes).
 *   - every call is stream-ordered on `stream` (a cudaStream_t) and returns immediately.
 *   - return value: TT_OK (0) or a negative TT_ERR_*; tt_last_error() gives the message
 *     (thread-local).  No exceptions, no silent fallbacks (the reference's zero-vector fallback,
 *     src/towers/cat_embed.py:180-190, is deliberately NOT replicated).
 *   - matrices are row-major; `ld*` are leading dimensions in ELEMENTS.
 */
#ifndef TWOTOWER_B200_H_
#define TWOTOWER_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TT_ABI_VERSION 1

#define TT_OK 0
#define TT_ERR_INVALID_ARG (-1)
#define TT_ERR_CUDA (-2)
#define TT_ERR_UNSUPPORTED (-3)
#define TT_ERR_WORKSPACE (-4)

/* element types of activations / gradients */
#define TT_F32 0
#define TT_BF16 1
#define TT_F16 2 /* feature stores only (tt_batch_gather) */
#define TT_I64 3
#define TT_I32 4

/* KeyedJaggedTensor value order.  The reference builds SAMPLE-major values ([B,K] flattened,
 * src/towers/pairs/unified_bid_data_loader.py:827-841) and decodes values[b*K+k]
 * (src/towers/cat_embed.py:109-110); canonical TorchRec order is KEY-major ([K][B]). */
#define TT_LAYOUT_SAMPLE_MAJOR 0
#define TT_LAYOUT_KEY_MAJOR 1

/* optimizer row-update modes of tt_ebc_backward_adam */
#define TT_ADAM_DENSE_EXACT 0 /* every arena row every step == torch.optim.Adam (scripts/train.py:231) */
#define TT_ADAM_SPARSE 1      /* only rows touched by the batch (TorchRec/SparseAdam-style; NOT the reference semantics) */
#define TT_ADAM_DENSE_UNTOUCHED 2 /* only rows the batch does NOT touch (g = wd*w): first half of a split exact-dense step. It
                                   * needs the stamps of tt_ebc_sort_rows but no gradients, so it can run underneath the towers;
                                   * TT_ADAM_SPARSE on the same sort then completes the step, bit-identical to TT_ADAM_DENSE_EXACT */

typedef void* tt_stream_t; /* cudaStream_t */

/* Device-resident optimizer/RNG state (64 bytes, zero-initialised by the caller).  Keeping it on
 * the device lets the whole step be captured in a CUDA graph: tt_hyper_advance() bumps the step
 * and derives lr_t / bias corrections on the GPU.  Replaces torch.optim.Adam's python-side step
 * bookkeeping and LambdaLR (scripts/train.py:231-242,327-330). */
typedef struct tt_adam_hyper {
  int64_t step;       /* optimizer steps taken so far (t) */
  int64_t sched_step; /* scheduler.step() calls so far (LambdaLR counter, starts at 0) */
  float lr;           /* lr_t used by the current step */
  float step_size;    /* lr_t / (1 - beta1^t) */
  float bc2_sqrt;     /* sqrt(1 - beta2^t) */
  float beta1, beta2, eps, weight_decay;
  float reserved;
  uint64_t rng_seed; /* dropout Philox key */
  uint64_t rng_step; /* dropout Philox step counter (== step) */
} tt_adam_hyper;

/* Lazy exact-replay Adam (device-resident context, built once per engine; kernels take a DEVICE pointer to it).
 * The reference's optimizer is dense: every embedding row moves every step, touched by the batch or not (g = wd * w;
 * scripts/train.py:231, nn.Embedding(sparse=False)). Instead of streaming every row every step, a row's (w, m, v) stay as they are
 * in HBM together with row_last[row] = the optimizer step they are current to; when the row is next looked up or updated, the
 * steps it sat out are replayed in registers, each with its own (step_size, bc2_sqrt) from the history ring written by
 * tt_hyper_advance_hist — the same operations in the same order as the dense update, bit for bit. tt_ebc_lazy_flush brings
 * every row up to date (before parameters are read from outside, and at least every hist_len - 1 steps). beta1 / beta2 / eps /
 * weight_decay must not change between two flushes. */
typedef struct tt_lazy_ctx {
  const float* arena_w; /* [rows, E] the sorted (large) tables */
  float* arena_m;
  float* arena_v;
  int32_t* row_last;    /* [rows] */
  const float* hist;    /* [hist_len][2]: (step_size_t, bc2_sqrt_t) at index t & (hist_len - 1) */
  int64_t hist_len;     /* power of two */
  const tt_adam_hyper* hyper;
} tt_lazy_ctx;

const char* tt_last_error(void);
int tt_version(void);
int tt_device_sm_count(void);

/* ---------------------------------------------------------------------------------------------
 * Optimizer bookkeeping.  t += 1; lr_t = base_lr * (s < warmup ? s / warmup : 1) with s the number
 * of scheduler steps so far (lr_1 = 0 when warmup > 0, scripts/train.py:235-242); then s += 1.
 * warmup_steps <= 0 disables the schedule. */
int tt_hyper_advance(tt_adam_hyper* d_hyper, float base_lr, int64_t warmup_steps, float beta1, float beta2,
                     float eps, float weight_decay, uint64_t rng_seed, tt_stream_t stream);
/* ... and records (step_size_t, bc2_sqrt_t) in the lazy-Adam history ring */
int tt_hyper_advance_hist(tt_adam_hyper* d_hyper, float base_lr, int64_t warmup_steps, float beta1, float beta2, float eps,
                          float weight_decay, uint64_t rng_seed, float* d_hist, int64_t hist_len, tt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Pooled embedding-bag lookup for all K tables of a tower in one launch.
 * Replaces CategoricalEmbedder._kjt_to_dict + per-key nn.Embedding + torch.cat
 * (src/towers/cat_embed.py:88-123,158-177) and writes straight into the tower's concat buffer
 * (src/towers/tower/base_tower.py:139).
 *   arena        f32 [total_rows, E]  all tables of all towers, table k starts at row d_row_base[k]
 *   d_row_base   i64 [K]   d_vocab i32 [K]    (device arrays; staged through shared memory)
 *   values       i64 [n_values] raw ids; clamped to [0, vocab-1] in-kernel (cat_embed.py:115-117)
 *   bag_offsets  i64 [B*K+1] or NULL.  NULL: one id per (sample,key), the reference's only case.
 *                Otherwise bag j (in `layout` order) owns values[bag_offsets[j] .. bag_offsets[j+1])
 *                and is sum-pooled (generalisation; oracle torch embedding_bag(mode="sum")).
 *   out          [B, out_ld] of out_dtype; bag (b,k) is written at columns out_col0 + k*E ..+E
 *   rows_out     i32 or NULL: clamped arena row of every value (bit-exact integer stage).  Without
 *                bag_offsets it is indexed by gradient slot b*K+k (whatever the value layout), so the
 *                backward can use identity slots; with bag_offsets it is indexed by value position.
 *   slots_out    i32 or NULL, same indexing: slot_base + b*K+k (the pooled-gradient vector the value
 *                contributes to).  slot_base lets two towers share one slot space (tower 1 starts at
 *                tower 0's B*K).
 */
int tt_ebc_forward(const float* arena, const int64_t* d_row_base, const int32_t* d_vocab, const int64_t* values,
                   const int64_t* bag_offsets, int64_t n_values, int32_t B, int32_t K, int32_t E, int32_t layout,
                   void* out, int32_t out_dtype, int64_t out_ld, int64_t out_col0, int32_t* rows_out,
                   int32_t* slots_out, int64_t slot_base, tt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Sparse backward, stage 1: deterministic sort-by-row dedup.  Stable LSD radix sort of
 * (row, slot) by row, then per-row segment bounds.  Replaces the index side of autograd's
 * embedding_dense_backward for every nn.Embedding (scripts/train.py:326).
 *   rows, slots      i32 [n]   (slots NULL => identity)
 *   sorted_rows/slots i32 [n]  out; inside one row the slots ascend (stable) => fixed summation order
 *   row_begin/row_end/row_stamp i32 [total_rows]: for every row present, [begin,end) into the sorted
 *                    arrays and row_stamp[row] = stamp; rows absent keep an older stamp (no memset).
 *   stamp            effective stamp = stamp + (d_stamp ? (int32)*d_stamp : 0); d_stamp is a device
 *                    counter so a captured CUDA graph sees a fresh stamp on every replay.
 *   d_n              optional device count: the number of valid pairs is *d_n (<= n, n = capacity). Used by
 *                    the row-wise sharded path where the owner only learns on the device how many ids it owns.
 */
int64_t tt_ebc_sort_workspace_bytes(int64_t n, int64_t total_rows);
int tt_ebc_sort_rows(const int32_t* rows, const int32_t* slots, int64_t n, int64_t total_rows, void* workspace,
                     int64_t workspace_bytes, int32_t* sorted_rows, int32_t* sorted_slots, int32_t* row_begin,
                     int32_t* row_end, int32_t* row_stamp, int32_t stamp, const int64_t* d_stamp,
                     const int32_t* d_n, tt_stream_t stream);
/* Same, for a captured train step: the device epoch counter *d_stamp is advanced by one inside the first kernel of the sort
 * and the new value is the stamp (stamp argument = 0), so no separate "counter += 1" node is needed in the graph. Every later
 * consumer of d_stamp on the same stream (tt_ebc_backward_adam) sees the advanced value. n must be > 0. */
int tt_ebc_sort_rows_advance(const int32_t* rows, const int32_t* slots, int64_t n, int64_t total_rows, void* workspace,
                             int64_t workspace_bytes, int32_t* sorted_rows, int32_t* sorted_slots, int32_t* row_begin,
                             int32_t* row_end, int32_t* row_stamp, int64_t* d_stamp, const int32_t* d_n,
                             tt_stream_t stream);

/* Sparse backward, stage 2: segment-sum of the pooled-output gradients in fixed order, fused with
 * the Adam row update.  Gradient slot s < n0 belongs to source 0 (sample s / K0, key s % K0, vector
 * at grad0 + b*ld0 + col0_0 + k*E); slots >= n0 to source 1 likewise.
 *   mode TT_ADAM_DENSE_EXACT: every row of [row_lo,row_hi) is updated with g = segsum + wd*w —
 *        bit-for-bit the update rule of torch.optim.Adam over dense nn.Embedding grads
 *        (scripts/train.py:231,327).
 *   mode TT_ADAM_SPARSE: only stamped rows are updated.
 *   grad_arena != NULL: also/only materialise the dense gradient ([total_rows,E], zeros elsewhere)
 *        for the autograd-compatible path (w/m/v may then be NULL).
 * piece_ws: f32 scratch of tt_ebc_piece_workspace_bytes(n, E) for rows with long segments.
 * Sharded tables (multi-GPU): grad_table != NULL is a device array [world*2] of peer-mapped base pointers
 * (rank r's source 0 / source 1 buffers); slot = r * n_per_rank + local slot and the segment-sum PULLS the
 * gradient vectors over NVLink — the gradient all-to-all is fused into this kernel. d_n: see tt_ebc_sort_rows. */
int64_t tt_ebc_piece_workspace_bytes(int64_t n, int32_t E);
int tt_ebc_backward_adam(float* arena_w, float* arena_m, float* arena_v, float* grad_arena, int64_t total_rows,
                         int32_t E, const int32_t* sorted_rows, const int32_t* sorted_slots, int64_t n,
                         const int32_t* row_begin, const int32_t* row_end, const int32_t* row_stamp, int32_t stamp,
                         const int64_t* d_stamp, const void* grad0, int64_t ld0, int64_t col0_0, int32_t K0, int64_t n0, const void* grad1,
                         int64_t ld1, int64_t col0_1, int32_t K1, int64_t n1, int32_t grad_dtype, void* piece_ws,
                         int64_t piece_ws_bytes, const tt_adam_hyper* d_hyper, int32_t mode, const int32_t* d_n,
                         const void* const* grad_table, int64_t n_per_rank, const int32_t* d_colmap0, const int32_t* d_colmap1,
                         tt_stream_t stream);
/* Lazy exact-replay variants (see tt_lazy_ctx). tt_ebc_forward_mixed_lazy: rows of the sorted tables that are behind are caught up
 * on the fly, read-only, through step hyper->step + upto_offset (-1 inside a train step, whose own update is still to come; 0 for a
 * forward outside of one). tt_ebc_backward_adam_lazy: only the rows the batch touched — replay of the steps they sat out, then this
 * step with its gradient — and row_last = step. Results equal TT_ADAM_DENSE_EXACT bit for bit once flushed. */
int tt_ebc_forward_mixed_lazy(const float* arena, const int64_t* d_row_base, const int32_t* d_vocab, const int64_t* values, int32_t B,
                              int32_t K, int32_t E, int32_t layout, void* out, int32_t out_dtype, int64_t out_ld, int64_t out_col0,
                              const int32_t* d_large_rank, int32_t K_large, int64_t row_sub, int32_t* rows_out,
                              const tt_lazy_ctx* d_lazy, int32_t upto_offset, tt_stream_t stream);
int tt_ebc_backward_adam_lazy(float* arena_w, float* arena_m, float* arena_v, int64_t total_rows, int32_t E,
                              const int32_t* sorted_rows, const int32_t* sorted_slots, int64_t n, const int32_t* row_begin,
                              const int32_t* row_end, const int32_t* row_stamp, const int64_t* d_stamp, const void* grad0, int64_t ld0,
                              int64_t col0_0, int32_t K0, int64_t n0, const void* grad1, int64_t ld1, int64_t col0_1, int32_t K1,
                              int64_t n1, int32_t grad_dtype, void* piece_ws, int64_t piece_ws_bytes, const tt_adam_hyper* d_hyper,
                              const int32_t* d_n, const void* const* grad_table, int64_t n_per_rank, const int32_t* d_colmap0,
                              const int32_t* d_colmap1, const tt_lazy_ctx* d_lazy, tt_stream_t stream);
int tt_ebc_lazy_flush(const tt_lazy_ctx* d_lazy, int64_t rows, int32_t E, tt_stream_t stream);

/* Two table classes.  "Large" tables go through the sort-dedup path above.  "Small" tables (a few dozen rows: 24 of the
 * reference's 32 notice keys have <= 25) are hit by every sample, so sorting them is wasted work and sharding them across
 * GPUs only creates hot rows: they live next to the dense parameters, their gradient is a dense one-hot reduction
 * (deterministic two-level sum, no atomics) and tt_dense_adam updates them (same exact-dense semantics).
 *   tt_ebc_forward_mixed   like tt_ebc_forward (one id per bag) for a tower whose keys are split into classes: keys with
 *                          d_vocab[k] <= 0 are skipped; d_large_rank[k] is the key's index among the large tables (or -1);
 *                          rows_out[b*K_large + rank] = arena row - row_sub (compact keys for the sort).
 *   tt_ebc_small_backward  gout[g_off .. g_off + vocab*E) = sum_b [id(b,key) == row] * grad[b, col .. col+E) for every
 *                          descriptor {int32 key, int32 vocab, int32 col, int64 g_off} (24 bytes, 8-byte aligned).
 *   d_colmap0/1 (tt_ebc_backward_adam): per large key, the column offset of its slice in the pooled-gradient row. */
int tt_ebc_forward_mixed(const float* arena, const int64_t* d_row_base, const int32_t* d_vocab, const int64_t* values, int32_t B,
                         int32_t K, int32_t E, int32_t layout, void* out, int32_t out_dtype, int64_t out_ld, int64_t out_col0,
                         const int32_t* d_large_rank, int32_t K_large, int64_t row_sub, int32_t* rows_out, tt_stream_t stream);
int64_t tt_ebc_small_workspace_bytes(int32_t B, int64_t small_floats);
int tt_ebc_small_backward(const int64_t* values, int32_t B, int32_t K, int32_t layout, const void* grad, int64_t ld,
                          int32_t grad_dtype, int32_t E, const void* d_small_tabs, int32_t n_small, int32_t max_vocab,
                          int64_t small_floats, float* gout, void* workspace, int64_t workspace_bytes, tt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Row-wise sharded tables across GPUs (no counterpart in the single-GPU reference; TorchRec RW convention:
 * block_k = ceil(V_k/world), owner = id / block_k, local row = local_base[k] + id - owner*block_k).
 *   tt_ebc_clamp_ids       ids_out i32 [B*K] in slot order b*K+k = clamp(values, 0, V_k-1)  (cat_embed.py:115-117)
 *   tt_shard_select_owned  ids_all i32 [world*n_per_rank] = every rank's clamped ids (rank-major; per rank n0 slots of
 *                          tower 0 with K0 keys, then tower 1 with K1 keys; table index = k resp. K0+k).
 *                          Deterministic stable compaction of the ids this rank owns: rows_out (local arena row),
 *                          slots_out (global slot = rank*n_per_rank + local slot), *d_n_owned.  Bit-exact integer stage.
 *   tt_ebc_push_rows       fused lookup + row all-to-all: for every owned id, gather the row and store it into the
 *                          requester's concat buffer dst_table[rank*2 + tower] (peer-mapped over NVLink) at
 *                          [b, col0 + k*E].  The matching gradient return is tt_ebc_backward_adam with grad_table. */
/* One-shot all-reduce (sum) of a small f64 vector over NVLink peer memory, in place, bit-identical on every rank (fixed rank
 * order). For the latency-bound reductions of the sharded step (SyncBN sums, loss partials, barriers) in place of a collective
 * library call. Every rank owns a mailbox of tt_peer_mailbox_bytes(world, n_max) bytes per call site, zero-initialised once
 * and peer-mapped; d_mailboxes is the device array of the world mailbox pointers (same order on every rank), d_epoch a
 * zero-initialised device counter per call site (graph-replay safe). One process per GPU only: kernels of different ranks wait
 * for each other. Rank skew is waited out; a watchdog (default 10 minutes; tt_peer_set_timeout_ms, 0 = never) traps on a peer
 * that never arrives instead of hanging the GPU for ever. Rank-local work between steps (checkpoint writes, evaluation) may be
 * as long as the watchdog allows. */
int64_t tt_peer_mailbox_bytes(int32_t world, int32_t n_max);
int tt_peer_set_timeout_ms(int64_t ms);
int tt_peer_allreduce_f64(double* data, int32_t n, int32_t rank, int32_t world, const void* const* d_mailboxes, int32_t n_max,
                          int64_t* d_epoch, tt_stream_t stream);
/* All-gather over NVLink peer memory: this rank's block (`bytes`, a multiple of 16) is stored into EVERY rank's destination
 * buffer (d_dst_bases: device array of the world peer-mapped base pointers, same order on every rank) at dst_offset_bytes
 * (normally rank * bytes); the call completes on the stream only when every rank's block has landed in this rank's buffer.
 * d_mailboxes / d_epoch: a mailbox site as for tt_peer_allreduce_f64 (only its flags are used); d_count: zero-initialised int32
 * per call site. tt_sum_slots_f32: slots f32 [world][n + n_tail]; out[i] = sum over r of slots[r][i] (i < n) in rank order —
 * with the all-gather, a deterministic all-reduce (the column sums of the logits) — and, for the n_tail values that ride behind
 * every slot, recip_out[r*n_tail + i] = 1 / slots[r][n + i] (the row scales of all ranks). One process per GPU only. */
int tt_peer_allgather(const void* src, int64_t bytes, int32_t rank, int32_t world, void* const* d_dst_bases, int64_t dst_offset_bytes,
                      const void* const* d_mailboxes, int64_t* d_epoch, int32_t* d_count, tt_stream_t stream);
int tt_sum_slots_f32(const float* slots, int32_t world, int64_t n, int64_t n_tail, float* out, float* recip_out, tt_stream_t stream);
int tt_ebc_clamp_ids(const int64_t* values, const int32_t* d_vocab, int32_t B, int32_t K, int32_t layout, int32_t* ids_out,
                     tt_stream_t stream);
int tt_ebc_clamp_ids_subset(const int64_t* values, const int32_t* d_vocab, const int32_t* d_keys, int32_t B, int32_t K, int32_t KL,
                            int32_t layout, int32_t* ids_out, tt_stream_t stream); /* ids_out[b*KL+kl] for keys d_keys[kl] only */
int64_t tt_shard_workspace_bytes(int64_t n_total);
int tt_shard_select_owned(const int32_t* ids_all, const int32_t* d_block, const int32_t* d_local_base, int64_t n_per_rank,
                          int64_t n0, int32_t K0, int32_t K1, int32_t world, int32_t me, int32_t* rows_out, int32_t* slots_out,
                          int32_t* d_n_owned, void* workspace, int64_t workspace_bytes, tt_stream_t stream);
/* All-to-all id routing: the SENDER buckets its ids by owner and pushes (local row, global slot) straight into each owner's receive
 * buffer over NVLink peer memory; the owner packs its W segments. Replaces all-gather + tt_shard_select_owned (whose per-rank work
 * and traffic grow with W): n * 8 bytes leave a rank instead of W * n * 4 arriving. The packed order (sender-major, slot order
 * inside) is the order tt_shard_select_owned produces: everything downstream is bit-identical.
 *   tt_shard_a2a_send     d_recv_rows / d_recv_slots: device arrays of `world` peer pointers to every rank's receive buffers
 *                         i32 [world][n_per_rank]; d_recv_counts: peer pointers to i32 [world] (segment lengths, index = sender).
 *                         A cross-rank barrier must separate it from tt_shard_a2a_compact on every rank.
 *   tt_shard_a2a_compact  this rank's receive buffers -> rows_out / slots_out (compact), *d_n_owned. */
int64_t tt_shard_a2a_workspace_bytes(int64_t n_per_rank, int32_t world);
int tt_shard_a2a_send(const int32_t* ids_local, const int32_t* d_block, const int32_t* d_local_base, int64_t n_per_rank, int64_t n0,
                      int32_t K0, int32_t K1, int32_t world, int32_t me, void* const* d_recv_rows, void* const* d_recv_slots,
                      void* const* d_recv_counts, void* workspace, int64_t workspace_bytes, tt_stream_t stream);
int tt_shard_a2a_compact(const int32_t* recv_rows, const int32_t* recv_slots, const int32_t* recv_counts, int32_t world,
                         int64_t n_per_rank, int32_t* rows_out, int32_t* slots_out, int32_t* d_n_owned, tt_stream_t stream);
int tt_ebc_push_rows(const float* arena, int32_t E, const int32_t* rows, const int32_t* slots, const int32_t* d_n_owned,
                     int64_t capacity, void* const* dst_table, int64_t n_per_rank, int64_t n0, int32_t K0, int32_t K1, int64_t ld0,
                     int64_t ld1, int64_t col0_0, int64_t col0_1, const int32_t* d_colmap0, const int32_t* d_colmap1,
                     int32_t out_dtype, tt_stream_t stream);
/* ... with the owner-side lazy catch-up of tt_ebc_forward_mixed_lazy */
int tt_ebc_push_rows_lazy(const float* arena, int32_t E, const int32_t* rows, const int32_t* slots, const int32_t* d_n_owned,
                          int64_t capacity, void* const* dst_table, int64_t n_per_rank, int64_t n0, int32_t K0, int32_t K1, int64_t ld0,
                          int64_t ld1, int64_t col0_0, int64_t col0_1, const int32_t* d_colmap0, const int32_t* d_colmap1,
                          int32_t out_dtype, const tt_lazy_ctx* d_lazy, int32_t upto_offset, tt_stream_t stream);
/* Multi-hot bags (pooling factor L, BASELINE configs[4]) over row-wise shards: every id-slot of a rank is slot = bag * L + l; the
 * routing entry points take the (key, position) pairs as K * L pseudo-keys. The OWNER pre-pools its ids of a bag and stores one
 * fp32 partial per (owner, bag) into the requester's peer-mapped partial buffer f32 [world][n_bags][E] (which the requester zeroes
 * at the start of the step, tt_zero_bytes: owners without an id of a bag write nothing); the REQUESTER sums the W partials in rank
 * order into its concat buffers (column of bag (b, k): d_colmap[k]). The gradient return is tt_ebc_backward_adam with the pseudo-
 * key column maps: every id of a bag pulls the bag's pooled gradient. */
int tt_ebc_push_pooled(const float* arena, int32_t E, const int32_t* rows, const int32_t* slots, const int32_t* d_n_owned,
                       int64_t capacity, void* const* part_table, int64_t n_per_rank, int32_t L, int64_t n_bags, int32_t me,
                       tt_stream_t stream);
int tt_ebc_sum_partials(const float* part, int32_t world, int64_t n_bags, int32_t E, void* c0, int64_t ld0, const int32_t* d_colmap0,
                        int64_t n0_bags, int32_t K0, void* c1, int64_t ld1, const int32_t* d_colmap1, int32_t K1, int32_t out_dtype,
                        tt_stream_t stream);
int tt_zero_bytes(void* p, int64_t bytes, tt_stream_t stream);

/* Dense-parameter Adam over a flat fp32 buffer (tower Linear/BatchNorm weights), same update rule.
 * Replaces torch.optim.Adam.step for the non-embedding parameters (scripts/train.py:327). */
int tt_dense_adam(float* w, const float* g, float* m, float* v, int64_t n, const tt_adam_hyper* d_hyper,
                  tt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Tower Linear layers (nn.Linear at src/towers/tower/base_tower.py:59,89,97).
 * forward : C[M,N] = act(A[M,K] · W[N,K]^T + bias)       act = ReLU iff relu != 0
 * dgrad   : dX[M,K] = dY[M,N] · W[N,K]
 * wgrad   : dW[N,K] = dY[M,N]^T · X[M,K];  dbias[N] = column sums of dY (NULL to skip)
 * W/bias/dW/dbias are fp32 masters; A/C/dY/dX/X are `act_dtype`.  workspace for wgrad split-K. */
int tt_linear_forward(const void* A, int64_t lda, const float* W, const float* bias, void* C, int64_t ldc, int64_t M,
                      int32_t N, int32_t K, int32_t act_dtype, int32_t relu, tt_stream_t stream);
int tt_linear_dgrad(const void* dY, int64_t ldy, const float* W, void* dX, int64_t ldx, int64_t M, int32_t N,
                    int32_t K, int32_t act_dtype, tt_stream_t stream);
int64_t tt_linear_wgrad_workspace_bytes(int64_t M, int32_t N, int32_t K);
int tt_linear_wgrad(const void* dY, int64_t ldy, const void* X, int64_t ldx, float* dW, float* dbias, int64_t M,
                    int32_t N, int32_t K, int32_t act_dtype, void* workspace, int64_t workspace_bytes,
                    tt_stream_t stream);

/* bf16 tensor-core variants (tcgen05.mma, fp32 accumulation in TMEM, TMA-fed): activations / gradients are bf16,
 * W_bf16 is a bf16 copy of the fp32 master weight (tt_cast_bf16), dW / dbias stay fp32.  All leading dimensions must
 * be multiples of 8 elements and all pointers 16-byte aligned (TMA).  out_dtype selects bf16 or fp32 for C / dX. */
int tt_cast_bf16(const float* src, void* dst, int64_t n, tt_stream_t stream);
/* up to four casts in one launch (segments with n <= 0 are skipped): the weight mirror and both dense inputs of a step */
int tt_cast_bf16_multi(const float* src0, void* dst0, int64_t n0, const float* src1, void* dst1, int64_t n1, const float* src2,
                       void* dst2, int64_t n2, const float* src3, void* dst3, int64_t n3, tt_stream_t stream);
int tt_linear_forward_bf16(const void* A, int64_t lda, const void* W_bf16, const float* bias, void* C, int64_t ldc,
                           int32_t out_dtype, int64_t M, int32_t N, int32_t K, int32_t relu, tt_stream_t stream);
int tt_linear_dgrad_bf16(const void* dY, int64_t ldy, const void* W_bf16, void* dX, int64_t ldx, int32_t out_dtype, int64_t M,
                         int32_t N, int32_t K, tt_stream_t stream);
int64_t tt_linear_wgrad_bf16_workspace_bytes(int64_t M, int32_t N, int32_t K);
int tt_linear_wgrad_bf16(const void* dY, int64_t ldy, const void* X, int64_t ldx, float* dW, float* dbias, int64_t M, int32_t N,
                         int32_t K, void* workspace, int64_t workspace_bytes, tt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * BatchNorm1d (train-mode batch statistics) + Dropout, the reference's Linear->ReLU->BN->Dropout
 * block (src/towers/tower/base_tower.py:88-93).  Split in three so data-parallel ranks can
 * all-reduce `sums` between stats and finalize (SyncBN == single-device BN on the global batch).
 *   tt_bn_stats     sums f64 [2N]: sum_b h, sum_b h^2 over the M local rows (deterministic order)
 *   tt_bn_finalize  mean/rstd f32 [N] from sums over `count` rows; running_mean/var momentum update
 *                   (unbiased var), num_batches_tracked += 1     (pass NULLs to skip running stats)
 *   tt_bn_apply     y = ((h - mean) * rstd * gamma + beta) * keep / (1 - p)
 *                   keep mask: mask_in u8 [M,N] if given (parity tests), else Philox(d_hyper, rng_stream)
 *   eval mode: call tt_bn_apply with mean = running_mean and rstd from tt_bn_rstd_from_var. */
int tt_bn_stats(const void* h, int64_t ldh, int64_t M, int32_t N, int32_t act_dtype, double* sums, void* workspace,
                int64_t workspace_bytes, tt_stream_t stream);
int64_t tt_bn_workspace_bytes(int64_t M, int32_t N);
int tt_bn_finalize(const double* sums, int64_t count, int32_t N, float eps, float momentum, float* mean, float* rstd,
                   float* running_mean, float* running_var, int64_t* num_batches_tracked, tt_stream_t stream);
int tt_bn_rstd_from_var(const float* var, int32_t N, float eps, float* rstd, tt_stream_t stream);
int tt_bn_apply(const void* h, int64_t ldh, int64_t M, int32_t N, int32_t act_dtype, const float* mean,
                const float* rstd, const float* gamma, const float* beta, float dropout_p, const uint8_t* mask_in,
                const tt_adam_hyper* d_hyper, uint32_t rng_stream, void* y, int64_t ldy, tt_stream_t stream);
/* backward: sums f64 [2N] = (sum_b dy', sum_b dy'*xhat) with dy' = dy * keep/(1-p);  then
 * dh = [h>0] * rstd*gamma*(dy' - sums0/count - xhat*sums1/count);  dgamma = affine_grad_scale * sums1, dbeta =
 * affine_grad_scale * sums0 (1 normally; with SyncBN the sums are already global on every rank, so exactly one rank passes 1
 * and the others 0 before the gradient all-reduce).
 * eval-mode backward (frozen stats) is dh = [h>0]*rstd*gamma*dy' : pass train = 0. */
int tt_bn_backward_reduce(const void* dy, int64_t lddy, const void* h, int64_t ldh, int64_t M, int32_t N,
                          int32_t act_dtype, const float* mean, const float* rstd, float dropout_p,
                          const uint8_t* mask_in, const tt_adam_hyper* d_hyper, uint32_t rng_stream, double* sums,
                          void* workspace, int64_t workspace_bytes, tt_stream_t stream);
int tt_bn_backward_apply(const void* dy, int64_t lddy, const void* h, int64_t ldh, int64_t M, int32_t N,
                         int32_t act_dtype, const float* mean, const float* rstd, const float* gamma,
                         float dropout_p, const uint8_t* mask_in, const tt_adam_hyper* d_hyper, uint32_t rng_stream,
                         const double* sums, int64_t count, int32_t train, float* dgamma, float* dbeta,
                         float affine_grad_scale, void* dh, int64_t lddh, tt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * F.normalize(p=2, dim=1, eps=1e-12) (src/towers/tower/base_tower.py:145) and its backward
 * dz = (de - e*(e·de)) / max(||z||, eps).  z/dz are fp32; e is out_dtype. */
int tt_l2norm_forward(const float* z, int64_t M, int32_t F, void* e, int32_t out_dtype, float* nrm,
                      tt_stream_t stream);
/* bf16 tower outputs twice: e (unit rows) and e_scaled = bf16(scale * z / |z|), rounded once from fp32 — the pre-scaled X operand of
 * the three-pass logits step (scale = log2(e) / temperature). */
int tt_l2norm_forward_scaled(const float* z, int64_t M, int32_t F, void* e_bf16, void* e_scaled_bf16, float scale, float* nrm,
                             tt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Fused tower kernels of the bf16 train step (tower_fused.cu): the MLP of base_tower.py:87-97,142-145 — Linear -> ReLU ->
 * BatchNorm1d -> Dropout -> Linear -> F.normalize — in two kernels forward and two backward (the BatchNorm batch statistics are the
 * only grid-wide dependencies). One hidden block, hidden width H1 and final dim F in {64, 128}, dense-projection width H0 a multiple
 * of 128 (tt_tower_fused_supported); other shapes use the per-stage entry points above.
 *   tt_linear_forward_bf16_stats  tt_linear_forward_bf16 + colstat f32 [tt_linear_stats_tiles(M)][2][N]: per 128-row tile the column
 *                                 sums of the stored (bf16-rounded) outputs and of their squares
 *   tt_tower_tail_forward         statistics (stat_part of n_tiles tiles, or — SyncBN — the all-reduced `sums` f64 [2][H1] over `count`
 *                                 rows) -> mean / rstd / running stats; y = dropout(bn(h)) (bf16, kept for Linear2's weight gradient);
 *                                 z = y W2^T + b2; nrm = |z|; e = z / nrm (bf16) and e_scaled = e * e_scale (may be NULL)
 *   tt_tower_head_backward        dz = (de - e (e.de)) / nrm -> bf16 (dz_bf16, and its per-tile column sums dz_colpart [n_tiles][F]);
 *                                 dy = dz W2; dyp = dy * keep/(1-p) (bf16); bn_part [n_tiles][2][H1] = column sums of dyp, dyp * xhat
 *   tt_tower_mid_backward         BatchNorm backward from bn_part (or all-reduced sums) + ReLU mask -> dh (bf16, dh_colpart
 *                                 [n_tiles][H1]); dc = dh W1 (bf16 [M, LD], all LD columns of the concat buffer) and dc_colpart
 *                                 [n_tiles][H0] (bias gradient of the dense projection); block 0 writes dgamma / dbeta (scaled by
 *                                 affine_grad_scale) and dbias2 = sum of dz_colpart
 *   tt_colpart_reduce_f64 / tt_colpart_reduce2   fold per-tile partials (fp64, tile order) into f64 sums (SyncBN hand-off) / into one
 *                                 or two fp32 bias gradients */
int tt_tower_fused_supported(int32_t H0, int32_t H1, int32_t F, int32_t LD);
int64_t tt_linear_stats_tiles(int64_t M);
int tt_linear_forward_bf16_stats(const void* A, int64_t lda, const void* W_bf16, const float* bias, void* C, int64_t ldc,
                                 int32_t out_dtype, int64_t M, int32_t N, int32_t K, int32_t relu, float* colstat,
                                 tt_stream_t stream);
int tt_tower_tail_forward(const void* h, int64_t M, int32_t H1, int32_t F, const float* stat_part, int32_t n_tiles, const double* sums,
                          int64_t count, const float* gamma, const float* beta, float eps, float momentum, float* mean, float* rstd,
                          float* running_mean, float* running_var, int64_t* num_batches_tracked, float dropout_p,
                          const uint8_t* mask_in, const tt_adam_hyper* d_hyper, uint32_t rng_stream, void* y, const void* W2_bf16,
                          const float* bias2, void* e, void* e_scaled, float e_scale, float* nrm, tt_stream_t stream);
int tt_tower_head_backward(const float* de, const void* e, const float* nrm, int64_t M, int32_t F, int32_t H1, const void* W2_bf16,
                           void* dz_bf16, float* dz_colpart, const void* h, const float* mean, const float* rstd, float dropout_p,
                           const uint8_t* mask_in, const tt_adam_hyper* d_hyper, uint32_t rng_stream, void* dyp, float* bn_part,
                           tt_stream_t stream);
int tt_tower_mid_backward(const void* dyp, const void* h, int64_t M, int32_t H1, int32_t LD, const void* W1_bf16, const float* bn_part,
                          int32_t n_tiles, const double* sums, int64_t count, const float* mean, const float* rstd, const float* gamma,
                          float* dgamma, float* dbeta, float affine_grad_scale, const float* dz_colpart, int32_t F, float* dbias2,
                          void* dh, float* dh_colpart, void* dc, int32_t H0, float* dc_colpart, tt_stream_t stream);
int tt_colpart_reduce_f64(const float* part, int32_t n_tiles, int32_t C, double* sums, tt_stream_t stream);
int tt_colpart_reduce2(const float* part0, int32_t C0, float* out0, const float* part1, int32_t C1, float* out1, int32_t n_tiles,
                       tt_stream_t stream);

int tt_l2norm_backward(const float* de, const void* e, int32_t e_dtype, const float* nrm, int64_t M, int32_t F,
                       float* dz, tt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * In-batch-negative similarity logits + symmetric softmax cross-entropy, logits never stored.
 * Replaces _compute_similarity_matrix / _compute_loss / _compute_metrics
 * (src/towers/two_tower_train_task.py:99-134,162-179) and their autograd.
 *   X [Bx,F] are the "row" embeddings held by this rank, Y [By,F] all "column" embeddings
 *   (By == Bx on one GPU; all-gathered item tower on several).  S = X·Y^T * inv_temp.
 *   Row r's positive is column diag_offset + r.  exp() is taken as exp(S - inv_temp) (rows are
 *   unit-norm so S <= inv_temp): sums are shifted by a constant, no running max needed.
 * forward outputs (all f32 unless noted):
 *   row_sumexp [Bx], col_sumexp [By] (over this rank's rows), row_sum_s [Bx], col_sum_s [By],
 *   diag [Bx] = S[r, diag_offset+r], row_argmax i32 [Bx] (first maximum, torch.argmax tie rule).
 * tt_simce_finalize turns (globally summed) sums into  inv_row [Bx], inv_col [By] for the backward
 * and partial f64[16][8]: 16 block partials of {sum_i l_row, sum_{j in my diag range} l_col, hits, sum diag, sum S, 0,0,0}
 * (the buffer must hold 128 doubles; distributed callers all-reduce all of it);
 * tt_simce_scalars maps (globally summed) partials to {loss, accuracy, pos_mean, neg_mean, gap}.
 * Workspace sizes (all tt_simce*_workspace_bytes) are NOT monotone in Bx: a short, wide problem splits a row block over more CTAs
 * (more partial results per row) than a square one — query them with the exact (Bx, By) of the calls, not with an upper bound. */
int64_t tt_simce_workspace_bytes(int64_t Bx, int64_t By, int32_t F);
int tt_simce_forward(const void* X, const void* Y, int64_t Bx, int64_t By, int32_t F, int32_t dtype, float inv_temp,
                     int64_t diag_offset, float* row_sumexp, float* col_sumexp, float* row_sum_s, float* col_sum_s,
                     float* diag, int32_t* row_argmax, void* workspace, int64_t workspace_bytes, tt_stream_t stream);
int tt_simce_finalize(const float* row_sumexp, const float* col_sumexp, const float* row_sum_s,
                      const float* col_sum_s, const float* diag, const int32_t* row_argmax, int64_t Bx, int64_t By,
                      int64_t diag_offset, int64_t B_global, float inv_temp, float label_smoothing, float* inv_row,
                      float* inv_col, double* partial, tt_stream_t stream);
int tt_simce_scalars(const double* partial, int64_t B_global, float* out5, tt_stream_t stream);
/* Split forward for the train step (bf16 path only). tt_simce_forward_sums produces what the BACKWARD needs — row / column sums
 * of exp, their reciprocals inv_row / inv_col (== tt_simce_finalize's) and diag; tt_simce_forward_metrics finishes what only the
 * loss value and the metrics need (row_sum_s, col_sum_s or NULL, row_argmax) from the partials left in the workspace, e.g. on a
 * side stream next to the backward pass. Together they write exactly what tt_simce_forward writes. The workspace
 * (tt_simce_split_workspace_bytes) must stay untouched between the two calls. */
int64_t tt_simce_split_workspace_bytes(int64_t Bx, int64_t By, int32_t F);
int tt_simce_forward_sums(const void* X, const void* Y, int64_t Bx, int64_t By, int32_t F, int32_t dtype, float inv_temp,
                          int64_t diag_offset, float* row_sumexp, float* col_sumexp, float* inv_row, float* inv_col, float* diag,
                          void* workspace, int64_t workspace_bytes, tt_stream_t stream);
int tt_simce_forward_metrics(const void* X, const void* Y, int64_t Bx, int64_t By, int32_t F, int32_t dtype, float inv_temp,
                             float* row_sum_s, float* col_sum_s, int32_t* row_argmax, void* workspace, int64_t workspace_bytes,
                             tt_stream_t stream);

/* Three-pass train step of the bf16 path (tcgen05). The fused step needs, per logit, only E = exp(S/T): the row sums (P1), the
 * column sums together with the company gradient (P2, the transposed problem with its row scale applied afterwards) and the notice
 * gradient (P3) — one pass fewer than forward(2) + backward(2), and no multiply per logit: the X operand of every pass is
 * pre-scaled by c1 = log2(e)/T when it is written (tt_l2norm_forward_scaled), so the accumulator is the exponent.
 * Replaces, for the fused train step, src/towers/two_tower_train_task.py:99-134 + its autograd backward.
 *   Xs            bf16 [Bx,F]  c1 * unit rows (this rank's rows)        Y   bf16 [By,F]  unit rows (all ranks' rows)
 *   Ys            bf16 [By,F]  (gscale * k / sumexp_y[j]) * Y[j,:]       (k = exp(-1/T): the public sums are the shifted ones,
 *                              sum_j exp(S_ij/T - 1/T), as everywhere in this header)
 *   workspace     tt_simce3_workspace_bytes; the P1 partials stay valid for tt_simce3_metrics until the next tt_simce3_rows.
 * tt_simce3_rows         P1: partial row sums / chunk maxima -> workspace; diag[i] = S_ii / T
 * tt_simce3_rows_finish  row_sumexp, inv_row = 1 / row_sumexp; optional Xs_out = (gscale * k / row_sumexp) * X_unit (the Ys of P2)
 * tt_simce3_grad         P2 (inv_x == NULL: row scale deferred, row sums accumulated on the fly) or P3 (inv_x given)
 * tt_simce3_grad_finish  dX (fixed-order sum of the parts, target / label-smoothing corrections); deferred: also sumexp_out /
 *                        inv_out of these rows and optional Xs_out = (gscale * k / sumexp) * X_unit (the Ys of P3)
 * tt_simce3_scale_rows   Ys from exchanged sums (multi-GPU: the sums of other ranks' rows arrive by all-gather)
 * tt_simce3_metrics      sum_j S_ij / T and the arg-max column from the P1 partials (off the critical path) */
int64_t tt_simce3_workspace_bytes(int64_t Bx, int64_t By, int32_t F);
int tt_simce3_rows(const void* Xs, const void* Y, int64_t Bx, int64_t By, int32_t F, float inv_temp, int64_t diag_offset, float* diag,
                   void* workspace, int64_t workspace_bytes, tt_stream_t stream);
int tt_simce3_rows_finish(void* workspace, int64_t workspace_bytes, int64_t Bx, int64_t By, int32_t F, float inv_temp,
                          float* row_sumexp, float* inv_row, const void* X_unit, float gscale, void* Xs_out, tt_stream_t stream);
int tt_simce3_grad(const void* Xs, const void* Y, const void* Ys, int64_t Bx, int64_t By, int32_t F, float inv_temp,
                   int64_t diag_offset, const float* inv_x, float gscale, float label_smoothing, void* workspace,
                   int64_t workspace_bytes, tt_stream_t stream);
int tt_simce3_grad_finish(void* workspace, int64_t workspace_bytes, int64_t Bx, int64_t By, int32_t F, float inv_temp,
                          int64_t diag_offset, const void* Y, int32_t deferred, float gscale, float label_smoothing, int64_t B_global,
                          float* dX, float* sumexp_out, float* inv_out, const void* X_unit, void* Xs_out, tt_stream_t stream);
int tt_simce3_scale_rows(const void* Y_unit, const float* sumexp, float gscale, float inv_temp, int64_t By, int32_t F, void* Ys_out,
                         tt_stream_t stream);
int tt_simce3_metrics(const void* Xs, const void* Y, int64_t Bx, int64_t By, int32_t F, float* row_sum_s, int32_t* row_argmax,
                      void* workspace, int64_t workspace_bytes, tt_stream_t stream);

/* backward: dX[Bx,F] (f32) = dS · Y with
 *   dS_ij = gscale * (exp(S_ij - inv_temp) * (inv_x[i] + inv_y[j]) - 2*((1-ls)*[j == diag_offset+i] + ls/B_global))
 * gscale = dloss * 0.5 / B_global * inv_temp.  Call once with (X=notice, Y=company) for d notice and
 * once with (X=company, Y=notice, inv_x=inv_col, inv_y=inv_row) for d company. */
int tt_simce_backward(const void* X, const void* Y, int64_t Bx, int64_t By, int32_t F, int32_t dtype, float inv_temp,
                      int64_t diag_offset, const float* inv_x, const float* inv_y, float gscale,
                      float label_smoothing, int64_t B_global, float* dX, void* workspace, int64_t workspace_bytes,
                      tt_stream_t stream);
/* Test hook: make the tcgen05 (bf16) sim-CE kernels stage operands with the generic-proxy loader instead of TMA. */
int tt_debug_simce_force_manual_loads(int32_t on);
/* A/B knob of the three-pass logits kernels (results stay correct): in the gradient passes (P2 / P3) every fourth exp2 of a chunk
 * runs as a degree-4 polynomial on the FMA pipe (the MUFU pipe does 16 ex2/clk/SM; scripts/microbench/mufu2.cu). bit 0 = also in
 * the statistics pass P1 (measured slower: off by default), bit 1 = not in P2 / P3. Process-wide, not stream-ordered. */
int tt_debug_simce_knobs(int32_t bits);
/* Profiling: CTA 0 of the next metric-pass forward kernels writes clock64 stamps into an i64[3*2048] device buffer
 * ([mma | epilogue | loader][tile*4 + event]); pass NULL to switch it off. */
int tt_debug_simce_timeline(void* d_buf_4x2048_i64);
/* Lazily materialise S (eval / predict_batch / return_metrics "similarity_matrix" only). */
int tt_similarity_matrix(const void* X, const void* Y, int64_t Bx, int64_t By, int32_t F, int32_t dtype,
                         float inv_temp, float* S, int64_t lds, tt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Retrieval metrics without the B x B matrix.  Replaces
 * TwoTowerEvaluator.compute_recall_at_k / compute_mrr (src/evaluation/evaluator.py:20-71: two topk calls, a full
 * argsort and a Python loop over the rows of a materialised S) by one integer per row:
 *   rank[i] = #{j : S_ij > S_ip} + #{j < p : S_ij == S_ip},  p = diag_offset + i the positive's column
 * (the positive's 0-based position in a stable descending sort of row i), S = X·Yᵀ * inv_temp computed tile by tile with
 * the same fp32 main loop as tt_similarity_matrix (bit-identical logits), never stored.
 *   tt_eval_ranks              X f32 [Bx,F], Y f32 [By,F] -> diag f32 [Bx] (the positives' logits), rank i32 [Bx]
 *   tt_eval_ranks_from_matrix  the same ranks from a materialised S [Bx, lds] (the reference evaluator's entry points take S)
 *   tt_eval_rank_metrics       out f32 [nk+2] = {recall@ks[0..nk) = mean(rank < k), MRR = mean(1/(rank+1)), top-1 accuracy};
 *                              d_ks i32 [nk] on the device, nk <= 8; integer hit counts, fp64 reciprocal-rank sum, fixed order. */
int tt_eval_ranks(const void* X, const void* Y, int64_t Bx, int64_t By, int32_t F, int32_t dtype, float inv_temp,
                  int64_t diag_offset, float* diag, int32_t* rank, tt_stream_t stream);
int tt_eval_ranks_from_matrix(const float* S, int64_t Bx, int64_t By, int64_t lds, int64_t diag_offset, int32_t* rank,
                              tt_stream_t stream);
int tt_eval_rank_metrics(const int32_t* rank, int64_t B, const int32_t* d_ks, int32_t nk, float* out, tt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * On-GPU batch assembly from HBM-resident feature stores.  Replaces the CPU collate
 * (create_collate_fn, src/towers/pairs/unified_bid_data_loader.py:630-684: numpy fancy-indexing of
 * dataset.*_store['dense_projected'|'categorical'] by the batch's row indices), _build_batch_kjt (:827-841: [B,K] ids
 * flattened sample-major into the KJT values) and the H2D copy of the batch (scripts/train.py:261-273) for one tower:
 *   dense_out[b, 0:D]       = float(dense_store[idx[b], 0:D])          dense_store [n_rows, dense_ld] f32|bf16|f16
 *   values_out[slot(b,k)]   = int64(cat_store[idx[b], k])              cat_store   [n_rows, cat_ld]   i64|i32
 * slot(b,k) = b*K+k (TT_LAYOUT_SAMPLE_MAJOR, the reference's order) or k*B+b. Either output may be NULL. The outputs can be the
 * static input buffers of a captured train step, so only idx (8 bytes per sample and tower) crosses PCIe. The store must be
 * 16-byte aligned with dense_ld % 4 == 0; 128-bit stores are used when dense_out is 16-byte aligned with out_ld % 4 == 0.
 * An index outside [0, n_rows) yields zeros / id 0 and adds 1 to *d_err (device int32, may be NULL; the reference raises
 * IndexError at this point, the host-side wrapper turns a non-zero *d_err into the same exception). */
int tt_batch_gather(const void* dense_store, int32_t dense_dtype, int64_t dense_ld, int32_t D, const void* cat_store,
                    int32_t cat_dtype, int64_t cat_ld, int32_t K, int64_t n_rows, const int64_t* idx, int32_t B, float* dense_out,
                    int64_t out_ld, int64_t* values_out, int32_t layout, int32_t* d_err, tt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TWOTOWER_B200_H_ */
