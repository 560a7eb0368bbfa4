/*
 * ugeval.h — C ABI of libugeval.so, the B200-native (sm_100a) replacement for UnrealGo's batched
 * network-evaluation path.  Plain pointers and sizes only; no TensorFlow, torch or C++ types.
 *
 * Every entry point names the reference interface it replaces (paths relative to the reference tree).
 * Conventions: functions returning int give 0 on success, >0 for the reference's "returns false",
 * <0 where the reference would throw (std::runtime_error); ug_last_error() holds the message.
 * Exactly one thread may drive an evaluator handle at a time (same contract as the reference: only
 * UctSearch::NetworkEvalThread touches the evaluator, search/UctSearch.cpp:164-205); the leaf queue
 * (ug_queue_*) is the multi-producer entry.
 */
#ifndef UGEVAL_H
#define UGEVAL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UG_BOARD_SIZE 19          /* GO_MAX_SIZE,      config/BoardStaticConfig.h:11-15 */
#define UG_NUM_POINTS 361         /* GO_MAX_ONBOARD,   config/BoardStaticConfig.h:27    */
#define UG_NUM_MOVES 362          /* GO_MAX_MOVES,     config/BoardStaticConfig.h:28    */
#define UG_NUM_PLANES 17          /* NUM_MAPS,         funcapproximator/DlTFNetworkEvaluator.h:13 */
#define UG_HISTORY 8              /* (17-1)/2,         gouct/GoUctGlobalSearch.h:166    */
#define UG_MASK_WORDS 12          /* ceil(361/32) */

/* value_transform, funcapproximator/DlConfig.h:8-13 */
enum { UG_TR_IDENTITY = 0, UG_TR_FLIP_SIGN = 1, UG_TR_FLIP_PROB = 2 };
/* DataFormat, funcapproximator/DataFormat.h:22-25 */
enum { UG_DF_HWC = 0, UG_DF_CHW = 1 };
/* arithmetic of the tower.  BF16 / FP16: tcgen05 kind::f16 MMAs (same rate), fp32 accumulation, 16-bit operands
 * — bf16 keeps fp32's range with an 8-bit mantissa, fp16 has an 11-bit mantissa and saturates at 65504.
 * FP32: SIMT validation build (slow, for parity work only). */
enum { UG_PREC_BF16 = 0, UG_PREC_FP32 = 1, UG_PREC_FP16 = 2 };

/*
 * Packed search leaf: what a search thread hands to the GPU encoder instead of running
 * GoUctGlobalSearchState::CollectFeatures (gouct/GoUctGlobalSearch.h:160-206) on the CPU.
 * black[k]/white[k] = stones on the board after undoing k moves (k = 0 is the leaf); when the history is
 * shorter than 8 the oldest position repeats (GoBoard::GetLastMove null rule, go/GoBoard.h:659-667).
 * Bit i of word i>>5 is point index i = (row-1)*19 + (col-1)  (GoPointUtil::Point2Index, board/GoPoint.h:139-146).
 * to_play: 0 = black (SG_BLACK), 1 = white (SG_WHITE)  (board/GoBlackWhite.h:8-9).
 */
typedef struct ug_packed_pos {
  uint32_t black[UG_HISTORY][UG_MASK_WORDS];
  uint32_t white[UG_HISTORY][UG_MASK_WORDS];
  uint32_t to_play;
  uint32_t reserved[3];
} ug_packed_pos; /* 784 bytes */
/* For a caller that already holds CollectFeatures' output (char[17][19][19], gouct/GoUctGlobalSearch.h:160-206) — the
 * reference's own search state — and wants to hand the leaf to the queue: the same information, 784 bytes instead
 * of 6137.  Any non-zero byte counts as a stone; plane 16 (1 - to_play) gives the side to move.  Host only. */
void ug_pack_planes(const int8_t* planes_chw, ug_packed_pos* out);

typedef struct ug_net_info {
  int blocks;           /* residual blocks */
  int filters;          /* tower width C */
  int policy_channels;  /* 1x1 policy conv outputs */
  int value_channels;   /* 1x1 value conv outputs */
  int value_hidden;     /* value FC width */
  int input_planes;     /* 17 */
  double flops_per_position; /* algorithmic FLOPs (SURVEY 8(d) formula evaluated on this net) */
  int conv_ctas;        /* 1 or 2: cta_group of the tower kernel in use */
  int sm_count;
  int conv_impl;        /* 1 = TMA-im2col tower kernel, 2 = resident-tile kernel (see ug_eval_set_conv_impl) */
  int small_batch_max;  /* batches up to this size run the small-batch tower kernel (see ug_eval_set_small_batch) */
  int precision;        /* UG_PREC_* in use: an fp16 evaluator falls back to bf16 for a checkpoint outside fp16's range */
  /* range calibration of the loaded checkpoint (ug_eval_load_checkpoint runs a few built-in positions through the fp32
   * validation tower): largest |layer output|, largest |batch-norm-folded weight|, and 65504 / max of the two */
  float calib_max_activation;
  float calib_max_weight;
  float fp16_headroom;
  /* fp16 auto-ranging: tensors stored times 2^-shift (0 = not rescaled); the residual stream's shift and the largest
   * shift of a block-internal tensor */
  int fp16_stream_shift;
  int fp16_max_block_shift;
  /* fp16 mode: how often a tower epilogue had to clamp an output to +-65504 since the evaluator was created (counted
   * per thread and tile, not per element); anything but 0 means results are not trustworthy */
  uint64_t saturation_events;
} ug_net_info;

typedef struct ug_eval ug_eval;
typedef struct ug_queue ug_queue;
typedef struct ug_dlcfg ug_dlcfg;

/* ------------------------------------------------------------------ evaluator (DlTFNetworkEvaluator<T>) */

/* DlTFNetworkEvaluator(graphPath, df) ctor, funcapproximator/DlTFNetworkEvaluator.cc:27-53: never loads;
 * defaults: checkpoint "bootstrap-ckpt", input "feature_input", outputs the two resnet/tower_0 names.
 * device = CUDA ordinal; precision = UG_PREC_*; max_batch = largest n a run call may pass (buffers are sized once). */
ug_eval* ug_eval_create(int device, int precision, int max_batch);
/* ~DlTFNetworkEvaluator, .cc:67-69 (safe when nothing was loaded). */
void ug_eval_destroy(ug_eval* h);
/* message of the last failing call on h (h may be NULL for ug_eval_create failures) */
const char* ug_last_error(const ug_eval* h);

/* SetNetworkInput / SetNetworkOutput, .cc:56-64: node names without ":0"; outputs[0]=policy, [1]=value. */
int ug_eval_set_io(ug_eval* h, const char* input, const char* const* outputs, int n_outputs);
/* LoadGraph, .cc:77-83 (+LoadMetaGraph .cc:115-167): ".meta" -> MetaGraphDef, ".pb" -> GraphDef, else 1
 * ("false").  Unreadable / unparsable file or an op pattern the loader does not recognise -> <0 (throws). */
int ug_eval_load_graph(ug_eval* h, const char* path);
/* UpdateCheckPoint, .cc:230-236,170-200: TF V2 bundle prefix (<p>.index + <p>.data-00000-of-00001); "" reuses
 * the last prefix; >0 on failure (never throws); may be called repeatedly (hot swap). */
int ug_eval_load_checkpoint(ug_eval* h, const char* prefix);
/* MetaGraphLoaded, .cc:239-241 */
int ug_eval_ready(const ug_eval* h);
/* 1 when graph and weights are both resident */
int ug_eval_has_weights(const ug_eval* h);
/* value_transform applied by run calls (the reference reads it from the DlConfig singleton, .cc:300) */
int ug_eval_set_value_transform(ug_eval* h, int vt);
/* description of the loaded network and its range calibration; waits for the evaluator's own stream (it reads the
 * device-side saturation counter): not for a hot loop */
int ug_eval_info(const ug_eval* h, ug_net_info* out);
/* use cta_group::1 or ::2 tower kernels (default 2); testing/benchmark knob */
int ug_eval_set_conv_ctas(ug_eval* h, int ctas);
/* Largest batch size <= n that fills whole waves of the persistent tower kernel (a wave = 74 CTA pairs x 256 pixel
 * rows = 52.47 positions on a 148-SM part): 213 positions cost 5 waves, exactly as much as 262; 209 cost 4.  The
 * leaf queue uses it when it flushes a partial batch.  Returns n itself below one wave. */
int ug_eval_efficient_batch(const ug_eval* h, int n);
/* the same rule for a given SM count (host-only: no device needed) */
int ug_efficient_batch_for_sms(int sm_count, int n);
/* tower kernel generation: 1 = TMA-im2col operand staging (conv3x3_tc.cu), 2 = shared-memory-resident input tile
 * with shifted-descriptor taps and TMA-staged epilogue (conv3x3_v2.cu, default); testing/benchmark knob */
int ug_eval_set_conv_impl(ug_eval* h, int impl);
/* What an fp16 evaluator does with a checkpoint whose calibrated range leaves less than 8x headroom below 65504:
 * 0 = auto (default): store the hot tensors times 2^-k (exact; folded into the packed weights and biases, accuracy
 *     unchanged) and, only if a folded weight itself is out of range, evaluate with bf16 operands instead (a note goes
 *     to stderr; ug_net_info.precision / fp16_stream_shift tell); 1 = refuse it (ug_eval_load_checkpoint returns >0);
 * 2 = load it anyway; 3 = bf16 operands without rescaling.  UGEVAL_FP16_RANGE=auto|refuse|ignore|bf16. */
int ug_eval_set_fp16_range_policy(ug_eval* h, int policy);
/* Batches of at most max_batch positions run the tower on the small-batch kernel (conv3x3_sb.cu: 128-row x nt-column
 * tiles on many SMs) — the batch sizes the reference's engine issues (1 .. MAX_BATCHES = 16,
 * funcapproximator/DlTFNetworkEvaluator.h:15; search/UctSearch.cpp:183-186).  max_batch 0 disables; nt = output
 * channels per tile (16/32/64/128, 0 = chosen from the grid size), plus 256 to force single-CTA tiles or 512 to force
 * CTA-pair tiles (cta_group::2, 256 rows x nt; nt >= 32) — otherwise chosen from the batch.  Default 32 / 0
 * (UGEVAL_SB_MAX, UGEVAL_SB_NT, UGEVAL_SB_PAIR=0|1); a batch also has to fit one wave of the SMs at 128 channels per
 * CTA (26 positions for 256 filters on 148 SMs). */
int ug_eval_set_small_batch(ug_eval* h, int max_batch, int nt);
/* carry the residual stream as a hi+lo pair of 16-bit tensors (1) or as one 16-bit tensor (0); numerics knob */
int ug_eval_set_residual_split(ug_eval* h, int on);

/* Evaluate(char feature[][17][19][19], double policies[][362], double value[], n), .cc:287-318.
 * HOST buffers, synchronous.  On failure outputs are left untouched (as .cc:315-317) and <0 is returned. */
int ug_eval_run_planes(ug_eval* h, const int8_t* chw, int n, double* policy, double* value);
/* Evaluate(char feature[][19][19][17], ...) overload, .cc:321-352. */
int ug_eval_run_planes_hwc(ug_eval* h, const int8_t* hwc, int n, double* policy, double* value);

/* Optional: page-lock a caller buffer that run_planes calls will use (feature, policy or value array — the reference's
 * EvalBuffer is one fixed allocation, search/UctSearch.h:126-130) so the evaluator copies straight from/to it instead
 * of going through its own pinned staging; outputs are then widened to double on the GPU.  0 ok, >0 refused
 * (the call still works through staging).  Unregister before freeing the memory; destroy unregisters the rest. */
int ug_eval_register_host(ug_eval* h, void* p, size_t bytes);
int ug_eval_unregister_host(ug_eval* h, void* p);

/* Native fast path: packed leaves in, fp32 out, HOST buffers, synchronous (H2D + kernels + D2H).
 * legal (optional, n x 12 words, bit i = move index i legal, 361 = pass): priors are renormalised over the
 * legal set as UctSearch::UpdatePrior does (search/UctSearch.cpp:1257-1270), illegal entries are 0. */
int ug_eval_run_packed(ug_eval* h, const ug_packed_pos* pos, int n, const uint32_t* legal, float* policy,
                       float* value);
/* Same with DEVICE buffers, asynchronous on `stream` (a cudaStream_t; NULL = the evaluator's own stream). */
int ug_eval_run_packed_device(ug_eval* h, const ug_packed_pos* d_pos, int n, const uint32_t* d_legal,
                              float* d_policy, float* d_value, void* stream);
/* block until everything queued on the evaluator's own stream has finished */
int ug_eval_sync(ug_eval* h);

/* ------------------------------------------------------------------ encoder hooks (CollectFeatures seam) */
/* UctThreadState::CollectFeatures(char[][19][19], 17), search/UctSearch.h:95: packed leaves -> the reference's
 * int8 planes [n][17][19][19], computed on the GPU (HOST buffers). */
int ug_encode_planes(ug_eval* h, const ug_packed_pos* pos, int n, int8_t* chw_out);
/* The tower's actual input: [n][19][19][17] bytes read back from the NHWC bf16 tensor the encoder wrote
 * (what TransformFeature, .cc:366-389, hands to TensorFlow as bool[n,19,19,17]). */
int ug_encode_nhwc(ug_eval* h, const ug_packed_pos* pos, int n, int8_t* hwc_out);
/* same, from reference planes (the K2 re-layout kernel used by ug_eval_run_planes) */
int ug_transform_planes(ug_eval* h, const int8_t* chw, int n, int8_t* hwc_out);

/* ------------------------------------------------------------------ debugging / parity */
/* After a run call: copy the fp32 image of intermediate tensor `layer` to host as [n][361][channels].
 * layer 0 = stem output, 1..2*blocks = tower convs in order; returns channels or <0. */
int ug_eval_debug_layer(ug_eval* h, int layer, int n, float* out);
/* Stop the tower after `layer` on subsequent run calls (-1 = run everything), so ug_eval_debug_layer can read it
 * before later blocks reuse the buffer.  Outputs of truncated runs are meaningless. */
int ug_eval_set_debug_stop(ug_eval* h, int layer);
/* Device-buffer kernel timings of the last ug_eval_bench call. */
typedef struct ug_bench_result {
  float ms_total;    /* CUDA-event time of `iters` full passes (encoder+tower+heads), device-resident inputs */
  float ms_encoder;  /* per-pass averages of the three kernel groups, measured in separate event brackets */
  float ms_tower;
  float ms_heads;
  int launches_per_pass;
} ug_bench_result;
int ug_eval_bench(ug_eval* h, const ug_packed_pos* host_pos, int n, int iters, ug_bench_result* out);

/* Raw kernel entry for unit tests: one 3x3 conv layer on device buffers (bf16 NHWC in/out, packed bf16
 * weights [cout][9*cin], fp32 bias, optional bf16 residual).  ck in {32,64}, ctas in {1,2}. */
int ug_k_conv3x3_bf16(int device, const void* d_in, const void* d_w, const float* d_bias, const void* d_res,
                      void* d_out, int n, int cin, int cout, int relu, int ck, int ctas);

/* Raw entry of the second-generation conv kernel (shared-memory-resident input tile, shifted-descriptor taps,
 * TMA-staged epilogue): 16-bit NHWC in/out on device buffers of n_alloc*361 rows, optional hi/lo residual and
 * lo output; runs once, then `iters` more times between CUDA events (ms_out = average launch time).
 * d_prof (optional, device, [74][64] int64): per-cluster stall-cycle counters of the last launch. */
int ug_k_conv3x3_v2(int device, const void* d_in, const void* d_w, const float* d_bias, const void* d_res_hi,
                    const void* d_res_lo, void* d_out_hi, void* d_out_lo, int n, int n_alloc, int cin, int cout,
                    int relu, int f16, int desc_mode, int iters, float* ms_out, long long* d_prof);

/* Raw entry of the small-batch conv kernel (same buffers as ug_k_conv3x3_v2; nt as ug_eval_set_small_batch).
 * d_prof (optional, device, [iters][148][16] int64): per-CTA time stamps of every timed launch (grids of <= 148 CTAs). */
int ug_k_conv3x3_sb(int device, const void* d_in, const void* d_w, const float* d_bias, const void* d_res_hi,
                    const void* d_res_lo, void* d_out_hi, void* d_out_lo, int n, int n_alloc, int cin, int cout,
                    int relu, int f16, int nt, int pdl, int iters, float* ms_out, long long* d_prof);

/* planning probe: clusters of `cluster_size` CTAs of the tower kernel that fit on the device at once (a part that
 * strands SM pairs for clusters of 4 returns fewer than SMs/4) */
int ug_debug_max_active_clusters(int device, int cluster_size);

/* ------------------------------------------------------------------ leaf-batching queue (Q1) */
/* Replaces EvalBuffer/EvalMsg + NetworkEvalThread's spin loop (search/UctSearch.h:59-72,126-130;
 * search/UctSearch.cpp:164-205): lock-free multi-producer ring, the consumer thread gathers up to max_batch
 * leaves, double-buffered pinned staging on two CUDA streams.  A partial batch is flushed timeout_us after its first
 * leaf when the GPU is idle; while a batch is still running the consumer keeps collecting until just before the GPU is
 * expected to run out of work (from the measured time per kernel wave).  timeout_us = 0: flush at once (one producer
 * that waits for every evaluation: latency mode, the consumer polls instead of sleeping). */
ug_queue* ug_queue_create(ug_eval* h, int max_batch, int timeout_us, int capacity);
void ug_queue_destroy(ug_queue* q);
/* called by search threads; copies *pos (and legal, optional 12 words); returns ticket >= 0, or <0 */
int64_t ug_queue_submit(ug_queue* q, const ug_packed_pos* pos, const uint32_t* legal);
/* blocks until the ticket's own result is ready; policy[362], *value */
int ug_queue_wait(ug_queue* q, int64_t ticket, float* policy, float* value);
/* non-blocking variant: 1 = result copied (ticket consumed), 0 = not ready yet, <0 = error */
int ug_queue_poll(ug_queue* q, int64_t ticket, float* policy, float* value);
/* counters: batches run, leaves evaluated, flushes caused by the timeout */
int ug_queue_stats(ug_queue* q, uint64_t* batches, uint64_t* leaves, uint64_t* flush_timeouts);
/* Compact priors (UctSearch::UpdatePrior search/UctSearch.cpp:1257-1270): with on != 0 a leaf that
 * was submitted with a legal mask gets back only its #legal renormalised priors, in ascending move-index order (the
 * order GenerateLegalMoves emits children, pass last), in policy[0 .. #legal): the head kernel packs them and the
 * device-to-host copy shrinks from 362 floats per leaf to the widest row of the batch.  Set before submitting. */
int ug_queue_set_compact_priors(ug_queue* q, int on);
/* out[0..count): batches, leaves, timeout flushes, full batches, batches cut back to whole kernel waves, failed batches,
 * deepest submission queue seen at a batch start, times the gather thread went to sleep, microseconds the evaluator
 * stream spent on batches (CUDA events around every batch: device busy time), batches launched onto an idle stream and
 * the microseconds it had been idle for (host clock) */
int ug_queue_stats_ex(ug_queue* q, uint64_t* out, int count);
/* pause the consumer, swap the checkpoint, resume (hot swap between searches, search/UctSearch.cpp:178-182) */
int ug_queue_swap_checkpoint(ug_queue* q, const char* prefix);

/* ------------------------------------------------------------------ self-play driver (the path's caller) */
/* UctDeepPlayer::SelfPlayOneGame / DoSearch / UctSearch::DeepUctSearchTree (search/UctDeepPlayer.cpp:61-121,
 * :288-349; search/UctSearch.cpp:803-854) for `games` independent games multiplexed over `threads` host threads
 * and one leaf queue: every playout expands and evaluates one leaf (1 network evaluation per playout). */
typedef struct ug_selfplay_config {
  int games;               /* concurrent games on this evaluator (sharding unit across GPUs) */
  int threads;             /* host threads driving them */
  int playouts_per_move;   /* MAX_SEARCH_ITERATIONS is 4000 in the reference (search/UctDeepPlayer.cpp:15) */
  int max_moves;           /* per game; 0 = 722 (m_maxGameLength, search/UctDeepPlayer.cpp:30) */
  int max_batch;           /* leaf-queue batch cap (<= evaluator max_batch) */
  int timeout_us;          /* leaf-queue flush timeout; 0 = 200 us, < 0 = none (a single game with one leaf in flight:
                            * nothing to batch, every microsecond of the window is latency per playout) */
  int reuse_tree;          /* dlcfg reuse_searchtree */
  int random_opening_moves;/* harness option: first k moves uniformly random (0 = reference behaviour) */
  float puct;              /* 0 = 2.5 (search/UctSearch.cpp:370) */
  uint64_t seed;
  /* UctDeepPlayer::WriteTFRecord (search/UctDeepPlayer.cpp:362-388): when record_dir is non-NULL every finished game
   * is written to <record_dir>/selfplay_<game_index_base + game>.tfrecord — walking back from the last non-pass
   * move: the 6137 plane bytes of the position after the move, the root visit counts of the search that chose it
   * (UctSearch::DeepUCTSelectBestChild, search/UctSearch.cpp:782-801: raw counts), and +-1 from the point of view of
   * the player who made the move, by Tromp-Taylor score (GoBoardUtil::TrompTaylorScore, go/GoBoardUtil.h:626-680). */
  const char* record_dir;
  float komi;              /* 0 = 7.5 (DEFAULT_KOMI, go/GoRules.h:9) */
  int game_index_base;     /* global index of this evaluator's first game (rank * games when sharded) */
  /* 0/1: one leaf per game in flight — the shipped FORCE_SINGLE_THREAD search (CMakeLists.txt:15).  k > 1: up to k
   * descents of one tree in flight with the reference's multi-thread virtual-loss bookkeeping (a counter on every node
   * of a descent until its backup, search/UctSearch.cpp:814-850; Select adds the parent's count minus one to the
   * visit total, :1066-1070; a child's virtual losses count as zero-valued visits in its mean — the statistics branch
   * of GetMeanValue, :665-679, not the USE_DIRECT_VISITCOUNT one of the shipped build, :661-664, which ignores them),
   * so that a single game can fill a batch.  Results then depend on the order evaluations return. */
  int leaves_per_game;
  /* harness option (0 = reference behaviour): every game starts from a position reached by this many uniformly random
   * legal moves (generated like the search's own moves, no pass unless forced) played before the first search, so
   * that a bounded benchmark sample searches middle-game and endgame positions — captures on the board, fewer legal
   * moves, longer capture/ko work per move — instead of near-empty boards.  Not recorded: incompatible with record_dir. */
  int prelude_moves;
} ug_selfplay_config;
typedef struct ug_selfplay_stats {
  uint64_t playouts, evaluations, batches, flush_timeouts, moves, games_finished, terminal_leaves;
  double seconds, playouts_per_s, mean_batch, mean_tree_nodes;
  /* where the host side stands: share of the search threads' time spent searching (the rest: asleep, waiting for
   * evaluations), search-thread microseconds per playout, and the leaf queue's batch bookkeeping */
  double search_busy_fraction, host_us_per_playout;
  uint64_t full_batches, wave_cut_batches, max_queue_depth, gather_sleeps;
  double gpu_busy_fraction;   /* share of the run the evaluator stream was working on a batch */
  uint64_t idle_launches;     /* batches launched when nothing was running or queued */
  double idle_gap_seconds;    /* total time the stream had been idle at those launches */
} ug_selfplay_stats;
/* moves_out (optional): [games][(max_moves or 722) + prelude_moves] policy indices (361 = pass), the prelude's moves
 * first; n_moves_out: [games] */
int ug_selfplay_run(ug_eval* h, const ug_selfplay_config* cfg, ug_selfplay_stats* stats, int16_t* moves_out,
                    int* n_moves_out);

/* test hooks: the driver's incremental board (goboard.h), to be checked against the oracle's restated GoBoard */
void* ug_board_new(void);
void ug_board_free(void* b);
int ug_board_legal(void* b, int idx);
void ug_board_play(void* b, int idx);
void ug_board_pack(void* b, ug_packed_pos* out);
int ug_board_generate(void* b, uint16_t* moves, uint32_t* legal);
int ug_board_to_play(void* b);
float ug_board_score(void* b, float komi);     /* Tromp-Taylor score used for the recorded reward (go/GoBoardUtil.h:626-680) */
void ug_board_planes(void* b, int8_t* chw_out); /* host CollectFeatures of the current position: char[17][19][19] */

/* ------------------------------------------------------------------ self-play records (DlTFRecordWriter) */
/* DlTFRecordWriter(filename), funcapproximator/DlTFRecordWriter.cc:18-22: TFRecord file of tf.train.Example
 * messages, uncompressed.  NULL if the file cannot be created. */
typedef struct ug_tfrecord ug_tfrecord;
ug_tfrecord* ug_tfrecord_open(const char* path);
/* WriteExample(feature17, f_len, policy, p_size, reward), .cc:59-76: keys "feature" (bytes), "policy" (floats),
 * "reward" (1 float). */
int ug_tfrecord_write_example(ug_tfrecord* w, const void* feature17, size_t f_len, const float* policy,
                              size_t p_size, float reward);
/* WriteExample with caller-chosen keys, .cc:78-100 */
int ug_tfrecord_write_example_named(ug_tfrecord* w, const char* feature_name, const void* feature, size_t f_len,
                                    const char* policy_name, const float* policy, size_t p_size,
                                    const char* reward_name, float reward);
int ug_tfrecord_flush(ug_tfrecord* w);      /* Flush(), .cc:34-37 */
int64_t ug_tfrecord_close(ug_tfrecord* w);  /* ~DlTFRecordWriter, .cc:24-32; returns the number of records written */

/* ------------------------------------------------------------------ checkpoint info files (DlCheckPoint) */
/* GetCheckPointInfo, funcapproximator/DlCheckPoint.cc:162-172: three lines — name, sha1, url.  0 ok; 1 when the file is
 * missing or has fewer than three lines (outputs untouched, as the reference). */
int ug_ckptinfo_read(const char* path, char* name, char* sha1, char* url, int cap);
/* WriteCheckpointInfo, .cc:188-203 (no newline after the url); 1 when the file exists and override_if_exists is 0 */
int ug_ckptinfo_write(const char* path, const char* name, const char* sha1, const char* url, int override_if_exists);
/* CheckDefaultCheckPoint, .cc:130-137: <prefix>.index and <prefix>.data-00000-of-00001 both exist */
int ug_checkpoint_files_exist(const char* prefix);

/* ------------------------------------------------------------------ dlcfg (DlConfig) */
/* DlConfig(filename) + parse(), funcapproximator/DlConfig.cc:20-41. A missing file yields an empty map. */
ug_dlcfg* ug_dlcfg_open(const char* path);
void ug_dlcfg_close(ug_dlcfg* c);
/* get(key, default), DlConfig.cc:164-176; returns bytes written (excluding NUL), truncated to cap-1 */
int ug_dlcfg_get(const ug_dlcfg* c, const char* key, const char* dflt, char* out, int cap);
int ug_dlcfg_value_transform(const ug_dlcfg* c);   /* DlConfig.cc:228-246 */
int ug_dlcfg_reuse_search_tree(const ug_dlcfg* c); /* DlConfig.cc:221-226 */
/* nn_outputs split at ':' (DlConfig.cc:214-219); returns count, writes up to max names of cap bytes each */
int ug_dlcfg_network_outputs(const ug_dlcfg* c, char* out, int cap, int max);

/* Host-only (no CUDA call): parse graph_path (.meta/.pb), build the network plan from input/policy/value node
 * names (NULL = the reference defaults) and, if ckpt_prefix is non-empty, verify every variable the plan needs.
 * Writes a JSON description (or the error message) to out.  0 ok, 1 bad suffix, 2 checkpoint problem, <0 graph
 * unreadable/unsupported.  Same loader as LoadGraph/UpdateCheckPoint. */
int ug_inspect_model(const char* graph_path, const char* ckpt_prefix, const char* input, const char* policy_out,
                     const char* value_out, char* out, int cap);

/* Host-only: one DT_FLOAT variable of a TF V2 checkpoint bundle through the same reader UpdateCheckPoint uses
 * (.index SSTable walk, <prefix>.data-0000k-of-0000N shard files, masked crc32c per tensor).  Returns the element count
 * and fills dims[0..*ndim) (up to 8); -1 = unreadable (reason in err), -2 = cap too small. */
long long ug_bundle_read_float(const char* ckpt_prefix, const char* key, float* out, long long cap, long long* dims,
                               int* ndim, char* err, int err_cap);

/* library build info */
const char* ug_version(void);
/* CRC-32C (Castagnoli), the checksum of TF bundle entries; exported for the checkpoint tooling */
uint32_t ug_crc32c(const void* p, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* UGEVAL_H */
