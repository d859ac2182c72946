// ctcps_prebeam.cuh -- kernels of the pre-beam ("partial scoring") decode step, SURVEY.md section 8(f) N2.
// Included by ctcps_kernels.cu inside its anonymous namespace (one translation unit, one library).
//
// The reference scorer scores a subset of the vocabulary when it is handed `scoring_ids` (ctc_scorer.py:90-97,
// 117-121, 155-162, 196-202); its processor never does.  ESPnet's beam search -- where the scorer comes from
// (ctc_scorer.py:2) -- picks those ids as the top `pre_beam_size` tokens of the decoder scores of every hypothesis.
// With S candidates instead of V tokens per hypothesis a decode step has BW*S lanes (C2: 82 k instead of 12.8 M), so
// the layout that serves it is token-major: x_vt (B, V, ldt), a token's time series contiguous, and the kernels are
//
//   k_transpose_vt    (B,T,ldx) -> (B,V,ldt), once per generate() (a K-a that writes the token-major layout directly --
//                     warp-per-row reductions + warp-private transpose tiles -- measured 1.30 ms against 0.60 + 0.74 ms for
//                     k_init + this kernel, and was not kept)
//   k_prebeam_topk    scores[:, pad] = logzero (:325) + top-S of every row of the decoder scores (radix select)
//   k_psi_cand        log_psi / token score / joint score of the S candidates of every hypothesis: a dot product over t of
//                     the per-hypothesis stream (the same `lin` workspace k_psi_full consumes) with exp(x_vt[b, v, :]),
//                     one warp-wide coalesced row read per candidate -- no recursion: in lazy-state mode the forward
//                     variables are only recomputed for the survivors
//   k_cand_to_dense   the reference-shaped (BW,V) outputs for callers that need them (HF's beam search)
//
// The recursion itself (state of the W survivors) is k_select_lazy_* of ctcps_kernels.cu on the token-major view.

// (B,T,ldx) -> (B,V,ldt); frames t >= T of the padded rows are zero-filled (never summed: their lin entries are zero).
__global__ void __launch_bounds__(256) k_transpose_vt(const float *__restrict__ x, int ldx, int T, int V, float *__restrict__ xt, int ldt) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int v0 = blockIdx.x * 32, t0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const float *src = x + (size_t)b * T * ldx;
    float *dst = xt + (size_t)b * V * ldt;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int t = t0 + ty + k * 8, v = v0 + tx;
        tile[ty + k * 8][tx] = (t < T && v < V) ? src[(size_t)t * ldx + v] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int v = v0 + ty + k * 8, t = t0 + tx;
        if (v < V && t < ldt) dst[(size_t)v * ldt + t] = tile[tx][ty + k * 8];
    }
}

// ------------------------------------------------------------------------------------------
// Top-S of every row of the decoder scores (one CTA per row): sets scores[row, blank] = logzero in place (:325) and returns
// the S best (id, score) pairs, best first, equal scores by lower id (torch's stable descending sort).
//
// Fast path (rows of up to TOPK_NT*4*TOPK_MAXU = 8192 floats, 16-byte aligned): the row is read ONCE into registers (all of a
// thread's 128-bit loads in flight together).  A cheap, provably safe threshold replaces any selection pass: every
// thread takes the max of its elements, every warp sorts its 32 thread-maxima (bitonic, shuffles only) and publishes its
// k-th largest with k = ceil(S / #warps); tau = the smallest of those.  At least #warps * k >= S distinct elements are
// >= tau, so the S best are all >= tau -- and only a few dozen elements are (the thread maxima are extreme values).
// Those are appended to a shared list and ranked exactly by (score, id) counting.  ~2 k warp instructions per row
// instead of ~15 k for the radix select (measured 90 us at C2, issue-bound) or ~200 serialised warp-wide insertions.
//
// General path (any V, any alignment, or more than TOPK_CAP elements >= tau, e.g. a constant row): radix select over
// order-preserving keys kept in shared memory: up to four 8-bit histogram passes narrow the key of the S-th best, the
// boundary bin is finished by exact rank counting (or, for > TOPK_TIES equal scores, by id order).
// ------------------------------------------------------------------------------------------
constexpr int TOPK_NT = 256;
constexpr int TOPK_MAXU = 8;     // float4 per thread in the fast path: rows up to 8192 floats
constexpr int TOPK_TIES = 128;
constexpr int TOPK_CAP = 512;

__device__ __forceinline__ unsigned topk_key(float x) {
    const unsigned u = __float_as_uint(x);
    return (u & 0x80000000u) ? u : ~u & 0x7fffffffu;
}
__device__ __forceinline__ float topk_value(unsigned k) { return __uint_as_float((k & 0x80000000u) ? k : ~k & 0x7fffffffu); }

struct TopkShared {
    unsigned hist[256];
    unsigned sh_prefix, sh_need, sh_cnt, n_sel, n_tie, n_cand;
    float warp_kth[TOPK_NT / 32];
    Cand sel[BEAM_MAXK];    // the winners, unordered
    Cand ties[TOPK_TIES];   // members of the boundary bin (radix path)
    Cand cand[TOPK_CAP];    // elements >= tau (fast path)
};

// writes the S winners of sh.sel, best first
__device__ __forceinline__ void topk_emit(const TopkShared &sh, int S, int row, int64_t *__restrict__ ids, float *__restrict__ cand_att) {
    const int tid = threadIdx.x;
    if (tid < S) {
        const Cand me = sh.sel[tid];
        int rank = 0;
        for (int o = 0; o < S; ++o) rank += cand_beats(sh.sel[o].s, sh.sel[o].i, me.s, me.i) ? 1 : 0;
        ids[(size_t)row * S + rank] = me.i;
        cand_att[(size_t)row * S + rank] = me.s;
    }
}

// radix select over the V keys of the row in shared memory -> sh.sel[0..S)
__device__ __forceinline__ void topk_radix(TopkShared &sh, const unsigned *keys, int V, int S) {
    const int tid = threadIdx.x, lane = tid & 31;
    unsigned prefix = 0;   // decided high bits of the S-th best key
    unsigned need = S;     // how many winners still have to come from keys that start with `prefix`
    int bits = 0;          // number of decided bits
    if (tid == 0) sh.n_sel = 0, sh.n_tie = 0;
    for (int pass = 0; pass < 4; ++pass) {
        sh.hist[tid] = 0;
        __syncthreads();
        const int shift = 24 - 8 * pass;
        for (int base = 0; base < V; base += TOPK_NT) {
            const int i = base + tid;
            bool valid = i < V;
            unsigned key = 0;
            if (valid) {
                key = keys[i];
                valid = bits == 0 || (key >> (32 - bits)) == prefix;
            }
            const unsigned m = __ballot_sync(0xffffffffu, valid);
            if (valid) {  // a warp whose lanes all hit one bin issues a single atomic, otherwise match.any aggregates equal bins
                const unsigned bin = (key >> shift) & 0xffu;
                const int leader = __ffs(m) - 1;
                const unsigned bin0 = __shfl_sync(m, bin, leader);
                if (__all_sync(m, bin == bin0)) {
                    if (lane == leader) atomicAdd(&sh.hist[bin], (unsigned)__popc(m));
                } else {
                    const unsigned peers = __match_any_sync(m, bin);
                    if (lane == __ffs(peers) - 1) atomicAdd(&sh.hist[bin], (unsigned)__popc(peers));
                }
            }
        }
        __syncthreads();
        if (tid < 32) {  // first bin (ascending key) where the cumulative count reaches `need`
            unsigned c[8], tot = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) c[k] = sh.hist[lane * 8 + k], tot += c[k];
            unsigned incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            const unsigned excl = incl - tot;
            const unsigned hit = __ballot_sync(0xffffffffu, incl >= need);
            if (lane == __ffs(hit) - 1) {
                unsigned below = excl, ck = c[0];
                int k = 0;
#pragma unroll
                for (int q = 0; q < 7; ++q)
                    if (k == q && below + c[q] < need) below += c[q], ck = c[q + 1], k = q + 1;
                sh.sh_prefix = (prefix << 8) | (unsigned)(lane * 8 + k);
                sh.sh_need = need - below;
                sh.sh_cnt = ck;
            }
        }
        __syncthreads();
        prefix = sh.sh_prefix, need = sh.sh_need, bits += 8;
        if (sh.sh_cnt <= TOPK_TIES) break;  // uniform: the boundary bin is small enough to finish exactly
    }
    // bits < 32: keys whose top `bits` bits are below `prefix` win outright, keys equal to it go to the boundary list.
    // bits == 32 with more than TOPK_TIES members: all of them carry the same score; the lowest ids win.
    const bool exact_ties = bits == 32 && sh.sh_cnt > TOPK_TIES;
    for (int i = tid; i < V; i += TOPK_NT) {
        const unsigned key = keys[i];
        const unsigned hi = key >> (32 - bits);
        if (hi < prefix) {
            const unsigned pos = atomicAdd(&sh.n_sel, 1u);
            sh.sel[pos].s = topk_value(key), sh.sel[pos].i = i;
        } else if (hi == prefix && !exact_ties) {
            const unsigned pos = atomicAdd(&sh.n_tie, 1u);
            sh.ties[pos].s = topk_value(key), sh.ties[pos].i = i;
        }
    }
    __syncthreads();
    const unsigned base_sel = sh.n_sel;  // = S - need
    if (!exact_ties) {
        const unsigned nt = sh.n_tie;
        for (unsigned q = tid; q < nt; q += TOPK_NT) {
            const Cand me = sh.ties[q];
            unsigned rank = 0;
            for (unsigned o = 0; o < nt; ++o) rank += cand_beats(sh.ties[o].s, sh.ties[o].i, me.s, me.i) ? 1u : 0u;
            if (rank < need) sh.sel[base_sel + rank] = me;
        }
    } else if (tid < 32) {  // one score, many ids: walk the row in id order and keep the first `need`
        unsigned got = 0;
        for (int base = 0; base < V && got < need; base += 32) {
            const int i = base + lane;
            const bool hit = i < V && keys[i] == prefix;
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            const unsigned before = __popc(m & ((1u << lane) - 1u));
            if (hit && got + before < need) sh.sel[base_sel + got + before].s = topk_value(prefix), sh.sel[base_sel + got + before].i = i;
            got += __popc(m);
        }
    }
    __syncthreads();
}

template <int TOPK_U>  // float4 per thread held in registers by the fast path (0: general path only)
__global__ void __launch_bounds__(TOPK_NT, TOPK_U <= 5 ? 4 : 3) k_prebeam_topk(float *att, int V, int blank, int S,
                                                                               int64_t *__restrict__ ids,
                                                                               float *__restrict__ cand_att) {
    extern __shared__ __align__(16) unsigned keys[];  // V keys of the row (general path)
    __shared__ TopkShared sh;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int row = blockIdx.x;
    float *a = att + (size_t)row * V;
    const int n4 = V >> 2;
    if constexpr (TOPK_U > 0) {  // the host picked this instantiation: V % 4 == 0, V / 4 <= TOPK_NT * TOPK_U, 16-byte aligned rows
        const float NEG = -INFINITY;
        float4 v[TOPK_U];
        const float4 *a4 = reinterpret_cast<const float4 *>(a);
#pragma unroll
        for (int u = 0; u < TOPK_U; ++u) {
            const int q = u * TOPK_NT + tid;
            v[u] = q < n4 ? a4[q] : make_float4(NEG, NEG, NEG, NEG);
        }
        if (tid == 0) sh.n_cand = 0;
        const int bq = blank >> 2, bj = blank & 3;  // scores[:, pad] = logzero before the candidates are drawn (:325)
        float mx = NEG;
#pragma unroll
        for (int u = 0; u < TOPK_U; ++u) {
            if (u * TOPK_NT + tid == bq) (&v[u].x)[bj] = LZ;
            mx = fmaxf(mx, fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)));
        }
        // descending bitonic sort of the warp's 32 thread maxima
#pragma unroll
        for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                const float other = __shfl_xor_sync(0xffffffffu, mx, j);
                const bool keep_max = ((lane & j) == 0) == ((lane & k) == 0);
                mx = keep_max ? fmaxf(mx, other) : fminf(mx, other);
            }
        const int kw = (S + TOPK_NT / 32 - 1) / (TOPK_NT / 32);  // <= 8 for S <= 64
        if (lane == kw - 1) sh.warp_kth[wid] = mx;
        __syncthreads();
        float tau = sh.warp_kth[0];
#pragma unroll
        for (int w = 1; w < TOPK_NT / 32; ++w) tau = fminf(tau, sh.warp_kth[w]);
#pragma unroll
        for (int u = 0; u < TOPK_U; ++u) {
            const float x[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (x[j] >= tau) {
                    const unsigned pos = atomicAdd(&sh.n_cand, 1u);
                    if (pos < TOPK_CAP) sh.cand[pos].s = x[j], sh.cand[pos].i = (u * TOPK_NT + tid) * 4 + j;
                }
        }
        __syncthreads();
        const unsigned nc = sh.n_cand;
        if (nc <= TOPK_CAP) {
            for (unsigned q = tid; q < nc; q += TOPK_NT) {
                const Cand me = sh.cand[q];
                int rank = 0;
                for (unsigned o = 0; o < nc; ++o) rank += cand_beats(sh.cand[o].s, sh.cand[o].i, me.s, me.i) ? 1 : 0;
                if (rank < S) {
                    ids[(size_t)row * S + rank] = me.i;
                    cand_att[(size_t)row * S + rank] = me.s;
                }
            }
            if (tid == 0 && blank >= 0) a[blank] = LZ;
            return;
        }
        // too many elements share the top scores (e.g. a constant row): hand the row to the radix select
#pragma unroll
        for (int u = 0; u < TOPK_U; ++u) {
            const int q = u * TOPK_NT + tid;
            if (q < n4) reinterpret_cast<uint4 *>(keys)[q] = make_uint4(topk_key(v[u].x), topk_key(v[u].y), topk_key(v[u].z), topk_key(v[u].w));
        }
    } else {
        (void)n4;
        for (int i = tid; i < V; i += TOPK_NT) keys[i] = topk_key(i == blank ? LZ : a[i]);
    }
    __syncthreads();
    topk_radix(sh, keys, V, S);
    topk_emit(sh, S, row, ids, cand_att);
    if (tid == 0 && blank >= 0) a[blank] = LZ;
}

struct CandArgs {
    const float *xt;        // (B,V,ldt) token-major log-posteriors
    int ldt;
    const float *lin;       // (B*G, Tpad, HWP): exp(r_sum[t-1] - Gm), zero outside the summed frames (workspace of the lazy mode)
    const float *Gmax;      // (BW)
    const float *psic;      // (BW) linear-domain sum for the last label's column
    const float *s_prev;    // (BW) or null
    const int64_t *last_ids;
    const int64_t *ids;     // (BW,S)
    const float *cand_att;  // (BW,S) or null
    float omw, w;
    float *cand_log_psi, *cand_ts, *cand_joint;  // (BW,S); cand_ts / cand_joint may be null
    int B, W, T, V, S, blank, ol, G, HW, HWP, Tpad;
};

// CTA = (utterance b, hypothesis group g), one warp per hypothesis.  The group's lin block (Tpad x HWP, hypothesis
// innermost in the workspace) is staged transposed into shared memory with coalesced loads; then the warp walks its S
// candidates: the 32 lanes read a candidate's time series as consecutive 128-bit words (512 contiguous bytes per warp
// request -- a thread-per-candidate mapping read 16 bytes from 15-32 different rows per request and ran at a third of the
// speed, DRAM-pattern-bound), multiply with exp() against the lin column and reduce with shuffles.  The loads of several
// candidates are in flight together.  Lane s finishes candidate s (epilogue shared with the full-vocabulary kernels).
constexpr int CAND_CU = 4;      // candidates whose row reads are in flight together
constexpr int CAND_QMAX = 3;    // 512-byte row segments per candidate in flight (covers T <= 384 in one trip)
constexpr int CAND_SPLIT = 2;   // CTAs per (utterance, hypothesis group): 512 CTAs of 5 warps at C2 instead of 256 of 10

__global__ void __launch_bounds__(320) k_psi_cand(const CandArgs a) {
    extern __shared__ __align__(16) float lin_s[];  // [hypotheses of this CTA][ldt]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int part = blockIdx.x % CAND_SPLIT;
    const int bg = blockIdx.x / CAND_SPLIT;
    const int b = bg / a.G, g = bg - b * a.G;
    const int nhyp_g = min(a.HW, a.W - g * a.HW);            // hypotheses of the group
    const int hper = (a.HW + CAND_SPLIT - 1) / CAND_SPLIT;   // hypotheses per CTA (= warps per CTA)
    const int hh0 = part * hper;
    const int nhyp = max(0, min(hper, nhyp_g - hh0));
    const int ldt = a.ldt;
    {   // stage: global (t, hh) -> shared [hh - hh0][t], 128-bit loads, all of a thread's loads in flight together
        const float4 *src = reinterpret_cast<const float4 *>(a.lin + ((size_t)(b * a.G + g) * a.Tpad) * a.HWP);
        const int n4 = (ldt * a.HWP) >> 2;  // HWP % 4 == 0
        for (int i0 = 0; i0 < n4; i0 += blockDim.x * 4) {
            float4 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = i0 + k * blockDim.x + threadIdx.x;
                if (i < n4) v[k] = src[i];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = i0 + k * blockDim.x + threadIdx.x;
                if (i < n4) {
                    const int e0 = i * 4, t = e0 / a.HWP, hq = e0 - t * a.HWP;  // 4 consecutive hypotheses of frame t
                    const float x[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int hh = hq + j - hh0;
                        if (hh >= 0 && hh < nhyp) lin_s[hh * ldt + t] = x[j];
                    }
                }
            }
        }
    }
    __syncthreads();
    if (wid >= nhyp) return;
    const int h = b * a.W + g * a.HW + hh0 + wid;
    const float4 *l4 = reinterpret_cast<const float4 *>(lin_s + wid * ldt);
    const int start = a.ol > 1 ? a.ol : 1;
    const int nq = ldt >> 2;                            // float4 per row
    const int q_lo = ((a.ol == 0 ? 0 : start) >> 2);    // lin is zero before `start`
    const long long last = a.last_ids[h];
    const float gm = a.Gmax[h], pcl = a.psic[h];
    const float sp_row = a.s_prev != nullptr ? a.s_prev[h] : 0.f;
    const int64_t *ids = a.ids + (size_t)h * a.S;
    const float *xb = a.xt + (size_t)b * a.V * ldt;
    EpiArgs e;
    e.Gmax = a.Gmax, e.s_prev = a.s_prev, e.s_rs = 1, e.s_cs = 0, e.att = nullptr, e.omw = a.omw, e.w = a.w;
    e.log_psi = nullptr, e.token_scores = nullptr, e.joint = nullptr, e.V = a.V, e.blank = a.blank, e.ol = a.ol;
    const float4 lz4 = make_float4(LZ, LZ, LZ, LZ);

    for (int s0 = 0; s0 < a.S; s0 += 32) {   // lane s - s0 finishes candidate s
        float my_sum = 0.f, my_x0 = LZ;
        const int s1 = min(a.S, s0 + 32);
        long long my_id = s0 + lane < s1 ? ids[s0 + lane] : -1;  // one coalesced read; batches take their ids by shuffle
        if (my_id >= a.V) my_id = -1;
        for (int sb = s0; sb < s1; sb += CAND_CU) {
            float acc[CAND_CU], x0[CAND_CU];
            const float4 *row[CAND_CU];
#pragma unroll
            for (int u = 0; u < CAND_CU; ++u) {
                acc[u] = 0.f, x0[u] = LZ;
                const long long v = __shfl_sync(0xffffffffu, my_id, (sb + u - s0) & 31);  // -1 past the last candidate
                row[u] = (sb + u < s1 && v >= 0) ? reinterpret_cast<const float4 *>(xb + (size_t)v * ldt) : nullptr;
            }
            for (int qb = q_lo; qb < nq; qb += 32 * CAND_QMAX) {
                float4 l[CAND_QMAX], xv[CAND_CU][CAND_QMAX];
#pragma unroll
                for (int k = 0; k < CAND_QMAX; ++k) {
                    const int q = qb + k * 32 + lane;
                    l[k] = q < nq ? l4[q] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int u = 0; u < CAND_CU; ++u) xv[u][k] = (q < nq && row[u] != nullptr) ? __ldg(row[u] + q) : lz4;
                }
#pragma unroll
                for (int k = 0; k < CAND_QMAX; ++k) {
                    const int q = qb + k * 32 + lane;
#pragma unroll
                    for (int u = 0; u < CAND_CU; ++u) {
                        if (q == 0) x0[u] = xv[u][k].x;
                        acc[u] = fmaf(l[k].x, ex2_approx(xv[u][k].x * LOG2E), acc[u]);
                        acc[u] = fmaf(l[k].y, ex2_approx(xv[u][k].y * LOG2E), acc[u]);
                        acc[u] = fmaf(l[k].z, ex2_approx(xv[u][k].z * LOG2E), acc[u]);
                        acc[u] = fmaf(l[k].w, ex2_approx(xv[u][k].w * LOG2E), acc[u]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < CAND_CU; ++u) {
                const float tot = warp_sum(acc[u]);
                const float first = __shfl_sync(0xffffffffu, x0[u], 0);  // x[b, 0, v]: lane 0 holds q = 0 when q_lo == 0
                if (lane == sb + u - s0) my_sum = tot, my_x0 = first;
            }
        }
        const int s = s0 + lane;
        if (s < s1) {
            const long long vv = my_id;
            const bool ok = vv >= 0;
            float S_lin = my_sum;
            if (ok && last == vv) S_lin = pcl;  // phi = r_prev blank there (:117-124)
            const float av_in = a.cand_att != nullptr ? a.cand_att[(size_t)h * a.S + s] : 0.f;
            float lp, ts, jt, av;
            epi_lane(e, h, (int)vv, S_lin, gm, a.ol == 0 ? my_x0 : LZ, sp_row, av_in, lp, ts, jt, av);
            if (!ok) lp = LZ, ts = LZ, jt = LZ;
            const size_t o = (size_t)h * a.S + s;
            a.cand_log_psi[o] = lp;
            if (a.cand_ts != nullptr) a.cand_ts[o] = ts;
            if (a.cand_joint != nullptr) a.cand_joint[o] = jt;
        }
    }
}

// `start > end` early return of the reference (:138-145) for the candidate path: everything logzero.
__global__ void k_cand_all_logzero(const float *__restrict__ cand_att, float omw, float w, size_t n, float *cand_log_psi, float *cand_ts,
                                   float *cand_joint) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cand_log_psi[i] = LZ;
        if (cand_ts != nullptr) cand_ts[i] = LZ;
        if (cand_joint != nullptr) cand_joint[i] = __fadd_rn(__fmul_rn(omw, cand_att != nullptr ? cand_att[i] : 0.f), __fmul_rn(w, LZ));
    }
}

// The reference-shaped dense outputs of a candidate step (:156, :161-162, :173-176, :332): CTA per hypothesis row.
// Tokens that were not scored have log_psi = logzero, token score logzero - s_prev; candidates are scattered on top.
__global__ void __launch_bounds__(256) k_cand_to_dense(const float *__restrict__ att, const float *__restrict__ s_prev,
                                                       const int64_t *__restrict__ ids, const float *__restrict__ cand_log_psi,
                                                       const float *__restrict__ cand_ts, const float *__restrict__ cand_joint,
                                                       int V, int S, float omw, float w, int all_logzero, float *log_psi,
                                                       float *token_scores, float *joint) {
    const int h = blockIdx.x;
    const float sp = s_prev != nullptr ? s_prev[h] : 0.f;
    float ts0 = all_logzero ? LZ : LZ - sp;
    if (ts0 == 0.f) ts0 = LZ;
    const size_t o = (size_t)h * V;
    const float wt = __fmul_rn(w, ts0);
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
        if (log_psi != nullptr) log_psi[o + v] = LZ;
        if (token_scores != nullptr) token_scores[o + v] = ts0;
        if (joint != nullptr) joint[o + v] = __fadd_rn(__fmul_rn(omw, att[o + v]), wt);
    }
    __syncthreads();
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const long long v = ids[(size_t)h * S + s];
        if (v < 0 || v >= V) continue;
        const size_t c = (size_t)h * S + s;
        if (log_psi != nullptr) log_psi[o + v] = cand_log_psi[c];
        if (token_scores != nullptr) token_scores[o + v] = cand_ts[c];
        if (joint != nullptr) joint[o + v] = cand_joint[c];
    }
}

// ------------------------------------------------------------------------------------------
// N4 (the step before the path): the CTC head's GEMM at fp32 accuracy on the TF32 tensor cores.  logits = h W^T is
// computed as ONE TF32 GEMM over operands split into TF32-exact parts and stacked along K:
//   h' = [h_hi | h_lo | h_hi] (n, 3d),  W' = [W_lo | W_hi | W_hi] (V, 3d),  h' W'^T = h_hi W_lo + h_lo W_hi + h_hi W_hi
// with x_hi = tf32(x) (round to nearest) and x_lo = x - x_hi (exact in fp32); the dropped h_lo W_lo term is 2^-22
// relative, fp32 accumulation.  The two small cross terms come FIRST along K: the tensor cores' accumulator rounds toward
// zero at every k-step, a bias proportional to the running sum, so the large hi*hi part should pass through as few
// k-steps as possible.  This kernel does the split; the GEMM itself is a plain library call (cuBLAS).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_split_tf32(const float *__restrict__ x, long long n, int d, int weight_order, float *__restrict__ out) {
    const long long total = n * (long long)(d >> 2);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / (d >> 2);
        const int c = (int)(i - r * (d >> 2)) * 4;
        const float4 v = *reinterpret_cast<const float4 *>(x + r * d + c);
        const float in[4] = {v.x, v.y, v.z, v.w};
        float hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            unsigned t;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(in[j]));
            hi[j] = __uint_as_float(t);
            lo[j] = in[j] - hi[j];
        }
        const float4 h4 = make_float4(hi[0], hi[1], hi[2], hi[3]), l4 = make_float4(lo[0], lo[1], lo[2], lo[3]);
        float *o = out + r * 3 * d + c;
        *reinterpret_cast<float4 *>(o) = weight_order ? l4 : h4;
        *reinterpret_cast<float4 *>(o + d) = weight_order ? h4 : l4;
        *reinterpret_cast<float4 *>(o + 2 * d) = h4;
    }
}
