/*
 * ctc_prefix_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A scalar CPU restatement of the CTC prefix scorer of BUTSpeechFIT/huggingface_asr
 * (src/decoding/ctc_scorer.py, itself a copy of ESPnet's ctc_prefix_score.py).  It is the
 * checker the CUDA path is compared against.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may build, load or call it; the product
 * package (huggingface_asr_b200/) never does.
 *
 * Parity pinning: the reference ships NO tests or golden vectors for this path
 * (SURVEY.md section 4), so this oracle is pinned against outputs of the reference itself:
 * tests/golden/make_golden.py imports /root/reference/src/decoding/ctc_scorer.py unmodified,
 * runs it on seeded inputs and commits the vectors (the .npz files under tests/golden);
 * tests/test_oracle_vs_golden.py replays them through this file.
 *
 * Written from the algorithm's description, one scalar lane at a time (the reference is
 * whole-tensor torch code); every function cites the reference lines it restates.
 * Compile twice: -DREAL=float (ctcps_oracle32_*) and -DREAL=double (ctcps_oracle64_*, used to
 * adjudicate fp32 rounding disputes).  See oracle/Makefile.
 *
 * Layouts (row-major, last index fastest), identical to the reference's logical shapes:
 *   x        (B, T, V)      log-posteriors after padding            ctc_scorer.py:39-46
 *   r_prev   (T, 2, BW)     selected forward variables, [.,0]=non-blank, [.,1]=blank
 *   r        (T, 2, BW, S)  new forward variables, S = scoring_num or V
 *   log_psi  (BW, V), token_scores (BW, V), s_prev (BW, V) or NULL meaning the scalar 0.0
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef REAL
#define REAL float
#endif
#ifndef SUFFIX
#define SUFFIX 32
#endif
#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)
#define FN(name) CAT(CAT(ctcps_oracle, SUFFIX), CAT(_, name))

#define LOGZERO ((REAL)-10000000000.0) /* ctc_scorer.py:29 */

static inline REAL r_exp(REAL v) { return sizeof(REAL) == 4 ? (REAL)expf((float)v) : (REAL)exp((double)v); }
static inline REAL r_log(REAL v) { return sizeof(REAL) == 4 ? (REAL)logf((float)v) : (REAL)log((double)v); }

/* torch.logsumexp over two finite values: max + log(exp(a-max) + exp(b-max)).
 * Used for r_sum (ctc_scorer.py:115) and for both rows of the recursion (ctc_scorer.py:150-151). */
static inline REAL lse2(REAL a, REAL b) {
    REAL m = a > b ? a : b;
    return r_log(r_exp(a - m) + r_exp(b - m)) + m;
}

int FN(real_bytes)(void) { return (int)sizeof(REAL); }

/* OpenMP team size of THIS library's runtime (torchrun exports OMP_NUM_THREADS=1; the CPU baseline wants every core). */
#ifdef _OPENMP
#include <omp.h>
int FN(set_threads)(int n) {
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
}
#else
int FN(set_threads)(int n) { (void)n; return 1; }
#endif

/* F.log_softmax(encoder_logits, dim=-1)                                  ctc_scorer.py:279 */
void FN(log_softmax)(const REAL *logits, int64_t rows, int64_t V, REAL *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < rows; ++i) {
        const REAL *in = logits + i * V;
        REAL *o = out + i * V;
        REAL m = in[0];
        for (int64_t v = 1; v < V; ++v) m = in[v] > m ? in[v] : m;
        REAL s = 0;
        for (int64_t v = 0; v < V; ++v) s += r_exp(in[v] - m);
        REAL ls = r_log(s);
        for (int64_t v = 0; v < V; ++v) o[v] = in[v] - m - ls;
    }
}

/* In-place length padding: frames t >= len_b carry all mass on blank.    ctc_scorer.py:39-42 */
void FN(pad)(REAL *x, const int64_t *lens, int64_t B, int64_t T, int64_t V, int64_t blank) {
    for (int64_t b = 0; b < B; ++b) {
        int64_t l = lens[b];
        if (l < 0) l += T; /* python slice semantics of x[i, l:, :] for negative l */
        if (l < 0) l = 0;
        if (lens[b] < T) {
            for (int64_t t = l; t < T; ++t) {
                REAL *row = x + (b * T + t) * V;
                for (int64_t v = 0; v < V; ++v) row[v] = LOGZERO;
                row[blank] = 0;
            }
        }
    }
}

/* Initial state (state is None): r_prev[t,0,h] = logzero,
 * r_prev[t,1,h] = cumsum_t x[t,b,blank], same for the W hyps of b.       ctc_scorer.py:74-85 */
void FN(init_state)(const REAL *x, int64_t B, int64_t T, int64_t V, int64_t blank, int64_t W, REAL *r0) {
    const int64_t BW = B * W;
    for (int64_t b = 0; b < B; ++b) {
        REAL acc = 0;
        for (int64_t t = 0; t < T; ++t) {
            REAL xb = x[(b * T + t) * V + blank];
            acc = (t == 0) ? xb : acc + xb; /* torch.cumsum: sequential running sum */
            for (int64_t w = 0; w < W; ++w) {
                r0[(t * 2 + 0) * BW + b * W + w] = LOGZERO;
                r0[(t * 2 + 1) * BW + b * W + w] = acc;
            }
        }
    }
}

/* scoring_idmap[h, scoring_ids[h, s]] = s, others -1 (later s wins).     ctc_scorer.py:91-95 */
void FN(build_idmap)(const int64_t *scoring_ids, int64_t BW, int64_t S, int64_t V, int64_t *idmap) {
    for (int64_t i = 0; i < BW * V; ++i) idmap[i] = -1;
    for (int64_t h = 0; h < BW; ++h)
        for (int64_t s = 0; s < S; ++s) idmap[h * V + scoring_ids[h * S + s]] = s;
}

/*
 * CTCPrefixScoreTH.__call__ for margin == 0 / att_w is None.            ctc_scorer.py:58-178
 *   last_ids[h] = y[h][-1], ol = len(y[0]) - 1                           :68-69
 *   scoring_ids NULL (S == 0) => full vocabulary, snum = V               :98-102
 *   s_prev NULL => the python scalar 0.0 of the first step               :83
 * Outputs r (T,2,BW,snum), log_psi (BW,V), token_scores (BW,V); idmap (BW,V) when S > 0.
 * Returns 1 when the reference would take the "start > end" early return (:138-145; all
 * outputs logzero, r left as initialised), else 0.
 */
int FN(score_window)(const REAL *x, int64_t B, int64_t T, int64_t V, int64_t blank, const REAL *r_prev,
                     const REAL *s_prev, const int64_t *last_ids, int64_t ol, int64_t W,
                     const int64_t *scoring_ids, int64_t S, int64_t start, int64_t end, REAL *r, REAL *log_psi,
                     REAL *token_scores, int64_t *idmap) {
    /* start / end: the frame window of :127-136 (host scalars there too); without att_w, (max(ol,1), T) */
    const int64_t BW = B * W;
    const int64_t snum = S > 0 ? S : V;
    if (S > 0) FN(build_idmap)(scoring_ids, BW, S, V, idmap);

    if (start > end) { /* :138-145 */
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < T * 2 * BW * snum; ++i) r[i] = LOGZERO;
        if (ol == 0)
            for (int64_t h = 0; h < BW; ++h)
                for (int64_t s = 0; s < snum; ++s) {
                    int64_t v = S > 0 ? scoring_ids[h * S + s] : s;
                    r[h * snum + s] = x[((h / W) * T) * V + v];
                }
        for (int64_t i = 0; i < BW * V; ++i) log_psi[i] = token_scores[i] = LOGZERO;
        return 1;
    }

    if (S > 0) /* log_psi = full(logzero) before the scatter                :156 */
        for (int64_t i = 0; i < BW * V; ++i) log_psi[i] = LOGZERO;

#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t h = 0; h < BW; ++h) {
        const int64_t b = h / W;
        const REAL *xb_ = x + b * T * V;
        const int64_t c = last_ids[h];
        /* position of the last label inside this hyp's lane set (:117-124) */
        int64_t cpos = S > 0 ? idmap[h * V + c] : c;
        REAL *rsum = (REAL *)malloc(sizeof(REAL) * (size_t)T);
        for (int64_t t = 0; t < T; ++t) /* r_sum = logsumexp(r_prev, 1)  :115 */
            rsum[t] = lse2(r_prev[(t * 2 + 0) * BW + h], r_prev[(t * 2 + 1) * BW + h]);
        for (int64_t s = 0; s < snum; ++s) {
            const int64_t v = S > 0 ? scoring_ids[h * S + s] : s;
            const int is_last = (s == cpos);
#define PHI(t) (is_last ? r_prev[((t)*2 + 1) * BW + h] : rsum[(t)])
#define R(t, k) r[(((t)*2 + (k)) * BW + h) * snum + s]
            /* r = full(logzero); if ol == 0: r[0,0] = x_[0,0]           :106-113 */
            for (int64_t t = 0; t < start; ++t) { R(t, 0) = LOGZERO; R(t, 1) = LOGZERO; }
            for (int64_t t = end; t < T; ++t) { R(t, 0) = LOGZERO; R(t, 1) = LOGZERO; } /* frames past the window */
            if (ol == 0) R(0, 0) = xb_[0 * V + v];
            /* forward recursion                                          :148-151 */
            REAL rn = R(start - 1, 0), rb = R(start - 1, 1);
            for (int64_t t = start; t < end; ++t) {
                REAL nn = lse2(rn, PHI(t - 1)) + xb_[t * V + v];
                REAL nb = lse2(rn, rb) + xb_[t * V + blank];
                rn = nn; rb = nb;
                R(t, 0) = rn; R(t, 1) = rb;
            }
            /* log_psi = logsumexp(cat(log_phi_x[start:end], r[start-1,0]))   :154-167 */
            REAL m = R(start - 1, 0);
            for (int64_t t = start; t < end; ++t) {
                REAL term = PHI(t - 1) + xb_[t * V + v];
                m = term > m ? term : m;
            }
            REAL acc = 0;
            for (int64_t t = start; t < end; ++t) acc += r_exp(PHI(t - 1) + xb_[t * V + v] - m);
            acc += r_exp(R(start - 1, 0) - m);
            log_psi[h * V + v] = r_log(acc) + m; /* scatter for S > 0     :161-162 */
#undef PHI
#undef R
        }
        free(rsum);
    }
    /* blank exclusion, relative score, zero hack                         :173-176 */
    for (int64_t h = 0; h < BW; ++h) {
        log_psi[h * V + blank] = LOGZERO;
        for (int64_t v = 0; v < V; ++v) {
            REAL ts = log_psi[h * V + v] - (s_prev ? s_prev[h * V + v] : (REAL)0);
            token_scores[h * V + v] = (ts == 0) ? LOGZERO : ts;
        }
    }
    return 0;
}

/* The call without attention weights (:133-136): start = max(output_length, 1), end = T. */
int FN(score)(const REAL *x, int64_t B, int64_t T, int64_t V, int64_t blank, const REAL *r_prev,
              const REAL *s_prev, const int64_t *last_ids, int64_t ol, int64_t W,
              const int64_t *scoring_ids, int64_t S, REAL *r, REAL *log_psi, REAL *token_scores,
              int64_t *idmap) {
    return FN(score_window)(x, B, T, V, blank, r_prev, s_prev, last_ids, ol, W, scoring_ids, S, ol > 1 ? ol : 1, T, r,
                            log_psi, token_scores, idmap);
}

/*
 * CTCPrefixScoreTH.index_select_state.                                   ctc_scorer.py:180-207
 * best_ids (B,W) hold hyp*V + tok inside the utterance.  s_new is returned as the (BW,) vector
 * the reference broadcasts to (BW,V) (:194).  With idmap (S > 0) ids go through the
 * scoring_idmap, -1 -> 0 (:196-202).
 */
void FN(select)(const REAL *r, const REAL *log_psi, const int64_t *best_ids, const int64_t *idmap,
                int64_t B, int64_t W, int64_t T, int64_t V, int64_t S, REAL *r_new, REAL *s_new) {
    const int64_t BW = B * W;
    const int64_t snum = S > 0 ? S : V;
    for (int64_t j = 0; j < BW; ++j) {
        const int64_t b = j / W;
        int64_t vidx = best_ids[j] + b * W * V; /* :191 */
        s_new[j] = log_psi[vidx];               /* :193 */
        if (S > 0) {
            int64_t hyp_idx = best_ids[j] / V + b * W; /* :198 */
            int64_t label = best_ids[j] % V;           /* :199 */
            int64_t sidx = idmap[hyp_idx * V + label];
            if (sidx == -1) sidx = 0; /* :201 */
            vidx = sidx + hyp_idx * snum; /* :202 */
        }
        for (int64_t t = 0; t < T; ++t)
            for (int k = 0; k < 2; ++k) r_new[(t * 2 + k) * BW + j] = r[(t * 2 + k) * BW * snum + vidx]; /* :206 */
    }
}

/*
 * The arithmetic of CTCRescorerLogitsProcessor.__call__ around the scorer:
 * scores[:, pad] = logzero in place (:325), next = (1-w)*scores + w*ctc (:332), and the optional
 * eos/space trick (:333-349).  argmax = first maximal index (torch semantics).
 */
void FN(combine)(REAL *scores, const REAL *ctc, int64_t BW, int64_t V, int64_t pad, REAL w, int apply_trick,
                 int64_t eos, int64_t space, REAL trick_w, REAL *next) {
    for (int64_t h = 0; h < BW; ++h) {
        REAL *s = scores + h * V;
        const REAL *c = ctc + h * V;
        REAL *n = next + h * V;
        s[pad] = LOGZERO;
        int64_t am_s = 0, am_c = 0;
        for (int64_t v = 0; v < V; ++v) {
            n[v] = ((REAL)1 - w) * s[v] + w * c[v];
            if (s[v] > s[am_s]) am_s = v;
            if (c[v] > c[am_c]) am_c = v;
        }
        if (apply_trick && am_s == eos && am_c == space) {
            if (n[eos] < n[space] && trick_w * n[eos] > n[space]) n[eos] = n[eos] * trick_w;
        }
    }
}
