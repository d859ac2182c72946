/*
 * ctcps.h -- C ABI of the sm_100a CTC prefix scorer (libctcps_b200.so).
 *
 * This is the drop-in boundary for the ONE hot path of BUTSpeechFIT/huggingface_asr that this
 * repository accelerates: src/decoding/ctc_scorer.py (log-softmax -> prefix-score forward
 * recursion -> state select -> joint-score combine).  The reference is pure Python/torch and has
 * no FFI of its own; each entry point below names the reference lines it replaces, and
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - fp32 log domain, int64 ids, row-major; logzero = -1e10 (ctc_scorer.py:29);
 *   - return value: 0 = ok, < 0 = argument error (CTCPS_E_*), > 0 = cudaError_t;
 *     ctcps_error_string() decodes both;
 *   - the caller owns every buffer; the library keeps no global state besides the last error text.
 *
 * Shapes: B utterances, W hypotheses per utterance, BW = B*W, T frames, V vocabulary,
 * S = scoring_num (0 = full vocabulary), snum = S ? S : V.
 *   ldx : row stride (floats) of x_logp,  multiple of 4, >= V   (16-byte rows for TMA / float4)
 *   ldr : innermost stride (floats) of r, multiple of 4, >= snum
 */
#ifndef CTCPS_H_
#define CTCPS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTCPS_LOGZERO (-10000000000.0f)

enum {
    CTCPS_OK = 0,
    CTCPS_E_BADARG = -1,    /* null pointer, non-positive size, id out of range */
    CTCPS_E_ALIGN = -2,     /* ldx / ldr / pointer alignment violated */
    CTCPS_E_WORKSPACE = -3, /* workspace too small */
    CTCPS_E_NODRIVER = -4,  /* cuTensorMapEncodeTiled not obtainable from the driver */
    CTCPS_E_TOOBIG = -5     /* sizes exceed what the kernels index with 32-bit lanes */
};

int ctcps_version(void);
const char *ctcps_error_string(int code);

/* Round a vocabulary / candidate count up to the stride the kernels run fastest with (multiple of 64 floats =
 * 256-byte rows; any multiple of 4 is accepted by ctcps_score). */
int ctcps_padded_ld(int n);

/* Lazy state selection (ctcps_select_lazy, ctcps_select_lazy_candidates): 1 (default) = a warp per surviving column,
 * time-parallel: the recursion is an affine map in the (logsumexp, +) semiring, lanes compose the maps of their blocks of
 * frames, a warp scan hands every lane its incoming state (k_select_lazy_pscan; 2*ceil(T/32) + 6 dependent logsumexp levels
 * instead of T); 0 (env CTCPS_SELECT_PSCAN=0) = one thread per surviving column walks the T frames (k_select_lazy_scan:
 * bit-identical to the column the materialising kernel writes).  The two are fp32 evaluation orders of the same recursion;
 * against an fp64 run the time-parallel one is no further than the sequential one (tests/test_gpu_select_pscan.py), joint
 * scores agree to ~1e-6.  Returns the previous mode; any other argument only queries.  Process-wide. */
int ctcps_set_select_pscan(int mode);

/* L2 look-ahead of the lazy scoring kernel (ctcps_score_lazy / _topk): how many 16 KB chunks (8 frames x 512 tokens) beyond
 * its shared-memory ring every CTA prefetches into L2 with TMA.  Default 0 = none (env CTCPS_PSI_PREFETCH): measured, the
 * look-ahead only costs (0 / 2 / 4 / 8 chunks: 0.371 / 0.376 / 0.390 / 0.489 ms at C2); kept as an A/B switch.  Returns the
 * previous value; an argument outside [0, 64] only queries.  Process-wide; results do not depend on it. */
int ctcps_set_psi_prefetch(int chunks);
/* Native decode loop (ctcps_decode_step, full vocabulary): 1 (default; env CTCPS_SKIP_DONE) = the scoring kernel does not stream
 * the tiles of utterances whose beam search has finished (HF's loop scores them until the whole batch is done and discards the
 * result); the hypotheses returned are bit-identical.  0 = score every row every step.  Returns the previous setting; any
 * other argument only queries. */
int ctcps_set_skip_done(int on);
/* Lazy scoring kernel: 1 (default; env CTCPS_FRAME_WINDOW) = stream only the 8-frame chunks in which the previous state puts a
 * nonzero weight on some hypothesis of the group -- exp(r_sum[t-1] - offset) underflows to exactly 0 far behind and far ahead
 * of where a prefix ends in the audio, and such chunks add exactly 0 to every prefix score; results are bit-identical.
 * 0 = stream every frame from the prefix length to T.  Returns the previous setting; any other argument only queries. */
int ctcps_set_frame_window(int on);
/* Instrumentation (bench.py's roofline): a device uint64 every lazy scoring launch adds the number of chunks it streams to
 * (NULL = off); ctcps_stream_chunk_bytes(W) = bytes of one chunk (8 frames x 512 tokens of posteriors + the group's weights). */
int ctcps_set_stream_counter(void *device_u64);
int ctcps_stream_chunk_bytes(int W);

/* Widest hypothesis group one thread of the lazy scoring kernel accumulates (2..20, default 20, env CTCPS_PSI_MAX_GROUP):
 * beams wider than the group are split into several groups that each stream the x tile.  Changes the layout of the scoring
 * workspace: set it before the first call of a decode, not in between.  Returns the previous value; other arguments query. */
int ctcps_set_psi_max_group(int hyps);

/* Bytes of scratch ctcps_score needs for these sizes. */
int ctcps_workspace_bytes(int B, int T, int V, int W, int S, size_t *out_bytes);

/*
 * K-a.  Replaces F.log_softmax (ctc_scorer.py:279) + the length-padding loop (:39-42) + the
 * blank-column broadcast of self.x[1] (:44-46).
 *   in  logits (B,T,V) rows of stride ld_in;  lens (B) int64
 *   out x_logp (B,T,*) rows of stride ldx  (may alias logits when ld_in == ldx: in-place, as the
 *       reference pads its argument in place);  blank_lp (B,T) = x_logp[b,t,blank]
 *   apply_log_softmax = 0 when the input already holds log-posteriors (CTCPrefixScoreTH ctor).
 */
int ctcps_init(const float *logits, int ld_in, const int64_t *lens, int B, int T, int V, int blank,
               int apply_log_softmax, float *x_logp, int ldx, float *blank_lp, void *stream);

/* Row-wise log-softmax; replaces LogSoftmaxProcessor.__call__ (ctc_scorer.py:357-365). */
int ctcps_log_softmax(const float *in, int ld_in, float *out, int ld_out, int rows, int V, void *stream);

/*
 * Initial state, replaces the `state is None` branch (ctc_scorer.py:74-85):
 *   r0[t,0,h] = logzero, r0[t,1,h] = sum_{tau<=t} blank_lp[b,tau]   (sequential fp32 running sum,
 *   the order torch.cumsum uses), r0 is (T,2,BW).
 * t_begin > 0 continues a running sum for extend_state (ctc_scorer.py:231-256): rows < t_begin
 * of r0 are left untouched and the sum restarts from r0[t_begin-1,1,h].
 */
int ctcps_initial_state(const float *blank_lp, int B, int T, int W, int t_begin, float *r0, void *stream);

/*
 * K-b.  Replaces CTCPrefixScoreTH.__call__ (ctc_scorer.py:58-178), fused with the processor arithmetic
 * around it (ctc_scorer.py:325,332).  ctcps_score is the call without attention weights (frames
 * max(ol, 1) .. T, :133-136); ctcps_score_window takes the frame window [start, end) the reference
 * derives from att_w and margin on the host (:127-132): the recursion and the log_psi sum run over
 * start <= t < end, r stays logzero outside (1 <= start, 1 <= end <= T; start == 1 when ol == 0).
 *   in  x_logp (B,T,ldx), blank_lp (B,T)
 *       r_prev (T,2,BW)            selected state (or ctcps_initial_state output)
 *       s_prev                     NULL = scalar 0.0 (:83); else element (h,v) at
 *                                  s_prev[h*s_row_stride + v*s_col_stride]  ((BW,) vector: 1, 0)
 *       last_ids (BW) int64        y[h][-1] (:69);  ol = len(y[0]) - 1 (:68)
 *       scoring_ids (BW,S) int64   NULL / S == 0 = full vocabulary (:98-102)
 *       att_scores (BW,V)          NULL, or the attention log-probs: column `blank` is set to
 *                                  logzero IN PLACE (:325) and joint = (1-w)*att + w*ctc (:332)
 *   out r (T,2,BW,ldr)             forward variables, ALL T frames written (:106-113,148-151)
 *       log_psi (BW,V)  token_scores (BW,V)  (:154-176)
 *       joint (BW,V)               only with att_scores
 *       scoring_idmap (BW,V) int64 only with scoring_ids (:91-95)
 *   one_minus_w, w                 (float)(1 - ctc_weight) and (float)ctc_weight, rounded by the
 *                                  caller in double like the reference's Python scalars
 * `start > end` (ol > T, :138-145) yields r = logzero everywhere and log_psi = token_scores = logzero.
 */
int ctcps_score(const float *x_logp, int ldx, const float *blank_lp, const float *r_prev, const float *s_prev,
                int64_t s_row_stride, int64_t s_col_stride, const int64_t *last_ids, int ol, int B, int W, int T,
                int V, int blank, const int64_t *scoring_ids, int S, int64_t *scoring_idmap, float *att_scores,
                float one_minus_w, float w, float *r, int ldr, float *log_psi, float *token_scores, float *joint,
                void *workspace, size_t workspace_bytes, void *stream);
int ctcps_score_window(const float *x_logp, int ldx, const float *blank_lp, const float *r_prev, const float *s_prev,
                       int64_t s_row_stride, int64_t s_col_stride, const int64_t *last_ids, int ol, int start, int end,
                       int B, int W, int T, int V, int blank, const int64_t *scoring_ids, int S, int64_t *scoring_idmap,
                       float *att_scores, float one_minus_w, float w, float *r, int ldr, float *log_psi,
                       float *token_scores, float *joint, void *workspace, size_t workspace_bytes, void *stream);

/*
 * K-c.  Replaces CTCPrefixScoreTH.index_select_state (ctc_scorer.py:180-207).
 *   best_ids (B,W) int64 = hyp*V + tok inside the utterance; with scoring_idmap the lane is
 *   idmap[hyp,tok] (-1 -> 0, :196-202).
 *   out r_new (T,2,BW), s_new (BW)  (the reference's (BW,V) s_new is this vector broadcast, :194)
 */
int ctcps_select(const float *r, int ldr, const float *log_psi, const int64_t *best_ids, const int64_t *scoring_idmap,
                 int B, int W, int T, int V, int S, float *r_new, float *s_new, void *stream);

/*
 * Lazy-state ("survivor recompute") variants of K-b / K-c.  log_psi, token_scores and joint do not depend on the
 * new forward variables r, and index_select_state keeps only BW of the BW*V columns of r; so the (T,2,BW,V) state
 * need not exist.  ctcps_score_lazy computes the same outputs as ctcps_score (full vocabulary) without writing r;
 * ctcps_select_lazy re-runs the recursion of THAT step (same r_prev, last_ids, ol) for the selected
 * (hyp, token) columns and returns what ctcps_select would have gathered from r.  Same arithmetic per lane as the
 * materialising kernels.  HBM bytes per step drop from 8*T*BW*V (write) to 4*T*B*V (read).
 */
int ctcps_score_lazy(const float *x_logp, int ldx, const float *blank_lp, const float *r_prev, const float *s_prev,
                     int64_t s_row_stride, int64_t s_col_stride, const int64_t *last_ids, int ol, int B, int W, int T,
                     int V, int blank, float *att_scores, float one_minus_w, float w, float *log_psi,
                     float *token_scores, float *joint, void *workspace, size_t workspace_bytes, int workspace_prepared,
                     void *stream);
/* ctcps_score_lazy with the utterance lengths `xlens` (B) that ctcps_init padded with (NULL = none): frames past an utterance's
 * length hold logzero for every token but blank (:39-42), exp(x) is exactly 0 there, so their chunks are not streamed --
 * results are bit-identical, a ragged batch costs its real frames instead of B x T. */
int ctcps_score_lazy_lens(const float *x_logp, int ldx, const float *blank_lp, const int64_t *xlens, const float *r_prev,
                          const float *s_prev, int64_t s_row_stride, int64_t s_col_stride, const int64_t *last_ids, int ol, int B,
                          int W, int T, int V, int blank, float *att_scores, float one_minus_w, float w, float *log_psi,
                          float *token_scores, float *joint, void *workspace, size_t workspace_bytes, int workspace_prepared,
                          void *stream);

/* next_workspace (nullable): the workspace of the NEXT ctcps_score_lazy call.  When given, the scan also writes the
 * per-hypothesis stream that call needs (for r_prev = r_new, s_prev = s_new, last ids = the selected tokens,
 * ol + 1), and the call may pass workspace_prepared = 1 to skip its own preparation kernel. */
int ctcps_select_lazy(const float *x_logp, int ldx, const float *blank_lp, const float *r_prev, const int64_t *last_ids,
                      int ol, const float *log_psi, const int64_t *best_ids, int B, int W, int T, int V, float *r_new,
                      float *s_new, void *next_workspace, size_t next_workspace_bytes, void *stream);

/*
 * N1 (the step right after the path): one fused beam-search step around the processor output.  Replaces, for one
 * decode step, what transformers 4.39.3 beam_search + BeamSearchScorer.process do between two processor calls (and
 * what huggingface_asr_b200/beam_search.py restates with torch ops): candidates = joint + running beam score, top 2W of
 * (B, W*V), eos candidates ranked inside the top W go to the finished pool with score / len_norm, the first W non-eos
 * candidates continue, done test (early_stopping=False), rows of done utterances continue with pad.
 *   joint (BW,V) processor output;  beam_scores (B,W) in/out;  ids_cur/ids_next (BW, ld_ids) int64, first L columns
 *   valid (bos first), ids_next gets L+1 columns;  pool_* (B,W[,ld_pool]) finished hypotheses (scores -inf = empty);
 *   done (B) bytes.  If done_ring (host-visible, e.g. pinned memory) is given, the last CTA stores
 *   (step_tag << 32 | number of done utterances) into done_ring[step_tag % ring].  workspace: ctcps_beam_step_workspace_bytes
 *   bytes, 16-byte aligned, ZEROED once before the first step (it holds arrival tickets that the kernel resets itself).
 */
int ctcps_beam_step_workspace_bytes(int B, int W, size_t *out_bytes);
int ctcps_beam_step(const float *joint, float *beam_scores, const int64_t *ids_cur, int64_t *ids_next, int64_t ld_ids, int L,
                    int B, int W, int V, int eos, int pad, float len_norm, float *pool_scores, int64_t *pool_lens,
                    int64_t *pool_seqs, int64_t ld_pool, unsigned char *done, void *workspace, size_t workspace_bytes,
                    int64_t *done_ring, int ring, int64_t step_tag, int64_t *best_ids_out, void *stream);
/* best_ids_out (nullable, (B,W) int64): source hypothesis * V + token of every continuing beam -- the ESPnet ids that
 * index_select_state expects (ctc_scorer.py:180-191).  HF hands its processors token ids only, so the reference's
 * processor selects every state from hypothesis 0; a loop that owns its beam step can pass these. */

/*
 * N2: pre-beam (partial scoring) decode step.  The reference scorer scores only `scoring_ids` when given
 * (ctc_scorer.py:90-97,117-121,155-162,196-202); the policy that picks them -- the top `pre_beam_size` tokens of the
 * decoder scores of every hypothesis -- is ESPnet's (espnet/nets/batch_beam_search.py, BatchBeamSearch.batch_beam /
 * beam_search.py pre_beam_score_key = "full"), the library the scorer was copied from (ctc_scorer.py:2).
 * Layout: x_vt (B,V,ldt) token-major log-posteriors, ldt = ctcps_padded_lt(T), frames >= T zero-filled.
 * scoring_ids of a hypothesis must be unique (ctcps_prebeam_topk's are).
 */
int ctcps_padded_lt(int T);
int ctcps_transpose_vt(const float *x_logp, int ldx, int B, int T, int V, float *x_vt, int ldt, void *stream);

/* scores[:, blank] = logzero in place (:325), then the S best (id, score) of every row, best first, ties by lower id. */
int ctcps_prebeam_topk(float *att_scores, int BW, int V, int blank, int S, int64_t *scoring_ids, float *cand_att, void *stream);

/* log_psi / token score / joint score (:154-176, :332) of the S candidates of every hypothesis, lazy state (no r):
 * outputs are (BW,S), aligned with scoring_ids; s_prev is a (BW) vector or NULL; cand_att (BW,S) the decoder scores of
 * the candidates (NULL: no joint).  Same workspace (and workspace_prepared meaning) as ctcps_score_lazy. */
int ctcps_score_candidates(const float *x_vt, int ldt, const float *r_prev, const float *s_prev, const int64_t *last_ids, int ol,
                           int B, int W, int T, int V, int blank, const int64_t *scoring_ids, int S, const float *cand_att,
                           float one_minus_w, float w, float *cand_log_psi, float *cand_token_scores, float *cand_joint,
                           void *workspace, size_t workspace_bytes, int workspace_prepared, void *stream);

/* The (BW,V) tensors the reference returns for a candidate step: log_psi = logzero outside the candidates (:156),
 * token_scores = log_psi - s_prev (:175-176), joint = (1-w)*att + w*token_scores (:332).  Outputs are nullable. */
int ctcps_candidates_to_dense(const float *att_scores, const float *s_prev, const int64_t *scoring_ids, const float *cand_log_psi,
                              const float *cand_token_scores, const float *cand_joint, int BW, int V, int S, float one_minus_w,
                              float w, int ol, int T, float *log_psi, float *token_scores, float *joint, void *stream);

/* index_select_state (:180-207 with scoring_idmap, :196-202) after a candidate step scored in lazy state: like
 * ctcps_select_lazy, s_new from cand_log_psi; a token that was not scored selects candidate 0 and s_new = logzero. */
int ctcps_select_lazy_candidates(const float *x_vt, int ldt, const float *blank_lp, const float *r_prev, const int64_t *last_ids,
                                 int ol, const int64_t *scoring_ids, int S, const float *cand_log_psi, const int64_t *best_ids, int B,
                                 int W, int T, int V, float *r_new, float *s_new, void *next_workspace,
                                 size_t next_workspace_bytes, void *stream);

/* ctcps_beam_step over the candidates only: cand_joint / cand_ids are (BW,S); ranking, ties and bookkeeping are those of
 * the dense step on a joint tensor that is -inf outside the candidates.  S >= 2. */
int ctcps_beam_step_candidates(const float *cand_joint, const int64_t *cand_ids, int S, float *beam_scores, const int64_t *ids_cur,
                               int64_t *ids_next, int64_t ld_ids, int L, int B, int W, int V, int eos, int pad, float len_norm,
                               float *pool_scores, int64_t *pool_lens, int64_t *pool_seqs, int64_t ld_pool, unsigned char *done,
                               void *workspace, size_t workspace_bytes, int64_t *done_ring, int ring, int64_t step_tag,
                               int64_t *best_ids_out, void *stream);

/*
 * Fused scoring + candidate ranking for a decode loop that owns its beam search (what ctcps_decode_step runs for the
 * full vocabulary).  ctcps_score_lazy_topk is ctcps_score_lazy (same arithmetic, bit for bit) whose epilogue still writes
 * log_psi (BW,V) -- the next state selection reads W of its entries per utterance (:193) -- but, instead of writing the
 * (BW,V) joint scores for a second kernel to rank, ranks the 512 tokens x hypothesis-group scores of a tile itself by
 * key = joint + beam_scores[h] and publishes the tile's K = 2W best, best first (ties: lower hyp*V+tok first):
 *   tile_lists [B][lists_per_utterance][K] x {float key; int32 hyp*V+tok}   (8 bytes each; shorter lists are closed with
 *   key = -inf, index = INT32_MAX);  ctcps_topk_lists_shape gives lists_per_utterance and K.
 * Needs V % 4 == 0, ol <= T, s_prev a (BW) vector or NULL.
 * ctcps_beam_step_lists is the beam step over those lists: same ranking, bookkeeping and outputs as ctcps_beam_step on the
 * dense joint tensor, plus last_ids_out (BW) = the new last token of every row.
 */
int ctcps_topk_lists_shape(int B, int W, int V, int *lists_per_utterance, int *K);
int ctcps_score_lazy_topk(const float *x_logp, int ldx, const float *r_prev, const float *s_prev, const int64_t *last_ids, int ol,
                          int B, int W, int T, int V, int blank, const float *att_scores, float one_minus_w, float w,
                          const float *beam_scores, float *log_psi, float *tile_lists, void *workspace, size_t workspace_bytes,
                          int workspace_prepared, void *stream);
/* the same with `done` (B) flags of the beam search (NULL = none): tiles of finished utterances are neither streamed nor ranked,
 * their lists and log_psi rows keep their previous contents; and with the utterance lengths (NULL = none), see ctcps_score_lazy_lens */
int ctcps_score_lazy_topk_active(const float *x_logp, int ldx, const float *r_prev, const float *s_prev, const int64_t *last_ids, int ol,
                                 int B, int W, int T, int V, int blank, const float *att_scores, float one_minus_w, float w,
                                 const float *beam_scores, const unsigned char *done, const int64_t *xlens, float *log_psi, float *tile_lists,
                                 void *workspace, size_t workspace_bytes, int workspace_prepared, void *stream);
int ctcps_beam_step_lists(const float *tile_lists, int lists_per_utterance, float *beam_scores, const int64_t *ids_cur, int64_t *ids_next,
                          int64_t ld_ids, int L, int B, int W, int V, int eos, int pad, float len_norm, float *pool_scores,
                          int64_t *pool_lens, int64_t *pool_seqs, int64_t ld_pool, unsigned char *done, void *workspace,
                          size_t workspace_bytes, int64_t *done_ring, int ring, int64_t step_tag, int64_t *best_ids_out,
                          int64_t *last_ids_out, void *stream);

/*
 * Native decode-step driver: ONE host call enqueues a whole joint-decoding step -- [top-S candidates,] prefix scoring
 * fused with the joint combine, the beam-search step, and (on a side stream, overlapping the caller's next decoder
 * forward pass) the lazy state selection that also prepares the next scoring call.  It is what
 * huggingface_asr_b200/beam_search.py::joint_beam_search_fused does with 5-10 ctypes / torch calls per step; at
 * ~150 us of GPU work per step those calls were the bottleneck.  Same kernels, same results.
 *
 * The caller fills the session once per generate() (all buffers are the caller's; [2] = ping-pong by step parity) and
 * calls ctcps_decode_step(session, decoder scores of this step, step) with step = 0, 1, 2, ...; before step 0:
 * ids[0][:, 0] = bos, last_ids[0][:] = bos, beam_scores = {0, -1e9, ...}, pool_scores = -inf, done = 0, beam_ws zeroed.
 * After step n the running hypotheses are ids[(n+1) & 1][:, :n+2].
 */
typedef struct ctcps_decode_session {
    int32_t B, W, T, V;
    int32_t S;             /* 0: full vocabulary, lazy state (ctcps_score_lazy); >= 2: pre-beam candidates */
    int32_t blank, eos, pad;
    int32_t use_beam_idx;  /* 1: select states with source hypothesis * V + token; 0: tokens only, like the reference (:326-329) */
    int32_t ldx, ldt, ring;
    float one_minus_w, w, length_penalty;
    const float *x_logp;   /* (B,T,ldx), S == 0 */
    const float *x_vt;     /* (B,V,ldt), S > 0 */
    const float *blank_lp; /* (B,T) */
    const float *r0;       /* (T,2,BW) ctcps_initial_state */
    float *r_sel[2];       /* (T,2,BW) selected state */
    float *s_sel[2];       /* (BW) */
    int64_t *last_ids[2];  /* (BW) last token of every row */
    int64_t *cand_ids[2];  /* (BW,S)   S > 0 */
    float *cand_att[2];    /* (BW,S)   S > 0 */
    float *cand_log_psi[2];/* (BW,S)   S > 0 */
    float *cand_joint;     /* (BW,S)   S > 0 */
    float *log_psi[2];     /* (BW,V)   S == 0 */
    float *joint;          /* (BW,V)   S == 0 */
    void *score_ws;        /* ctcps_workspace_bytes */
    size_t score_ws_bytes;
    float *beam_scores;    /* (B,W) */
    int64_t *ids[2];       /* (BW, ld_ids) */
    int64_t ld_ids;
    float *pool_scores;    /* (B,W) */
    int64_t *pool_lens;    /* (B,W) */
    int64_t *pool_seqs;    /* (B,W,ld_pool) */
    int64_t ld_pool;
    unsigned char *done;   /* (B) */
    void *beam_ws;         /* ctcps_beam_step_workspace_bytes, zeroed once */
    size_t beam_ws_bytes;
    int64_t *done_ring;    /* host-visible, `ring` entries, or NULL */
    int64_t *best_ids;     /* (B,W) */
    void *side_stream;     /* from ctcps_async_create; NULL: the selection runs on `stream` */
    void *ev_step, *ev_select;
    float *tile_lists;     /* S == 0: B * lists_per_utterance * K * 2 floats (ctcps_topk_lists_shape) enables the fused
                              scoring + top-2W step: the (BW,V) joint tensor is never written (`joint` may then be NULL
                              unless a prefix can outgrow T: ol > T falls back to the dense step); NULL: the dense step */
    int64_t tag_base;      /* added to `step` in the tag published to done_ring: a serial number of the decode in the high
                              bits (a multiple of `ring`), so that a late write of an earlier decode never matches */
    const int64_t *xlens;  /* (B) utterance lengths as given to ctcps_init, or NULL: the full-vocabulary scoring kernel does not
                              stream the frames past an utterance's length (they contribute exactly 0) */
} ctcps_decode_session;

/* sizeof(ctcps_decode_session) as this library was compiled: lets a binding check its mirror of the struct. */
size_t ctcps_decode_session_size(void);
int ctcps_async_create(void **side_stream, void **ev_step, void **ev_select);
int ctcps_async_destroy(void *side_stream, void *ev_step, void *ev_select);
/* Timing events for callers without a CUDA runtime binding (bench.py times the scoring call inside a step with them). */
int ctcps_event_create(void **event);
int ctcps_event_destroy(void *event);
int ctcps_event_elapsed_ms(void *begin, void *end, float *ms);
/* ev_score_begin / ev_score_end (nullable): recorded on `stream` around the scoring call of the step. */
int ctcps_decode_step(const ctcps_decode_session *s, float *att_scores, int step, void *ev_score_begin, void *ev_score_end,
                      void *stream);
/* After the last step: makes `stream` wait for the selection still running on the side stream, so that work enqueued on
 * `stream` afterwards (e.g. freeing and reusing the session's buffers) is ordered behind it.  No host synchronisation. */
int ctcps_decode_finish(const ctcps_decode_session *s, void *stream);

/*
 * Optional eos/space trick of the processor (ctc_scorer.py:333-349), in place on `next`:
 * rows with argmax(att) == eos and argmax(ctc) == space and next[eos] < next[space] < k*next[eos]
 * get next[eos] *= k.
 */
int ctcps_eos_space_trick(const float *att_scores, const float *ctc_scores, float *next, int BW, int V, int eos,
                          int space, float k, void *stream);

/*
 * N4 (the step before the path): the CTC head itself.  Replaces Wav2Vec2ForCTC.lm_head (src/reguler/e_branchformer.py:245-252:
 * logits = hidden W^T + b, an fp32 Linear) followed by F.log_softmax (src/decoding/ctc_scorer.py:279) and the length padding
 * (:39-46), i.e. it produces what ctcps_init produces, from the encoder's hidden states instead of its logits.
 *   hidden (B*T, d) fp32;  bias (V) or NULL;  w_hi / w_lo: the head's weight (V, d) prepared ONCE by
 *   ctcps_head_prepare_weight (each ctcps_head_weight_bytes(V, d) bytes): scaled by a power of two into fp16's range, split
 *   into two fp16 parts (11 + 11 significand bits) and stored tile by tile in the swizzled image the kernel's TMA bulk copies
 *   drop into shared memory (opaque to the caller; w_hi ends with the scale)
 *   x_logp (B,T,ldx), blank_lp (B,T): as ctcps_init;  apply_log_softmax = 0: x_logp receives the raw logits (no padding)
 * A hand-written tcgen05 kernel (TMA-fed 3xFP16 UMMA, two TMEM accumulators -- large term / small cross terms -- bias and the
 * softmax statistics in the TMEM -> register drain, TMA stores) + one streaming normalisation pass; fp32-grade accuracy (max
 * logit error ~1.4e-5 at d = 512, |logit| ~ 10).  d must be a multiple of 16.  workspace: ctcps_head_workspace_bytes(B*T, d),
 * 16-byte aligned.
 * ctcps_split_hi_lo is the plain elementwise split (hi = round-to-nearest TF32, lo = x - hi), same shapes.
 */
int ctcps_head_workspace_bytes(int64_t n, int d, size_t *out_bytes);
int ctcps_head_weight_bytes(int V, int d, size_t *out_bytes);
int ctcps_head_prepare_weight(const float *weight, int V, int d, float *w_hi, float *w_lo, void *stream);
int ctcps_split_hi_lo(const float *x, int64_t count, float *hi, float *lo, void *stream);
int ctcps_ctc_head(const float *hidden, const float *w_hi, const float *w_lo, const float *bias, const int64_t *lens, int B, int T, int d,
                   int V, int blank, int apply_log_softmax, float *x_logp, int ldx, float *blank_lp, void *workspace,
                   size_t workspace_bytes, void *stream);

/*
 * Round-1 form of the same head, kept for A/B (CTCPS_HEAD=cublas): operand split for a library GEMM (Wav2Vec2ForCTC.lm_head, src/reguler/
 * e_branchformer.py:245-252) at fp32 accuracy on the TF32 tensor cores.  x (n,d) -> out (n,3d):
 * weight_order = 0: [hi | lo | hi] (activations), 1: [lo | hi | hi] (weights); hi = tf32(x) rounded to nearest,
 * lo = x - hi.  out_h out_W^T = hi lo + lo hi + hi hi in one TF32 GEMM with fp32 accumulation (small terms first).
 */
int ctcps_split_tf32(const float *x, int64_t n, int d, int weight_order, float *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CTCPS_H_ */
