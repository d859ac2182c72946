// ctcps_prebeam.cuh -- kernels of the pre-beam ("partial scoring") decode step, SURVEY.md section 8(f) N2.
// Included by ctcps_kernels.cu inside its anonymous namespace (one translation unit, one library).
//
// The reference scorer scores a subset of the vocabulary when it is handed `scoring_ids` (ctc_scorer.py:90-97,
// 117-121, 155-162, 196-202); its processor never does.  ESPnet's beam search -- where the scorer comes from
// (ctc_scorer.py:2) -- picks those ids as the top `pre_beam_size` tokens of the decoder scores of every hypothesis.
// With S candidates instead of V tokens per hypothesis a decode step has BW*S lanes (C2: 82 k instead of 12.8 M), so
// the layout that serves it is token-major: x_vt (B, V, ldt), a token's time series contiguous, and the kernels are
//
//   k_transpose_vt    (B,T,ldx) -> (B,V,ldt), once per generate()
//   k_prebeam_topk    scores[:, pad] = logzero (:325) + top-S of every row of the decoder scores (warp-register lists)
//   k_psi_cand        log_psi / token score / joint score of the S candidates of every hypothesis: a dot product over t of
//                     the per-hypothesis stream (the same `lin` workspace k_psi_full consumes) with exp(x_vt[b, v, :]) --
//                     no recursion: in lazy-state mode the forward variables are only recomputed for the survivors
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

// One CTA (4 warps) per row of the decoder scores: sets scores[row, blank] = logzero in place (:325) and returns the S
// best (score, id) pairs of the row, best first, equal scores by lower id.  Each warp keeps the sorted top-S of its quarter
// of the row in registers (warp_list_insert), the four lists are merged by rank counting in shared memory.
template <int KL>
__global__ void __launch_bounds__(BEAM_NT) k_prebeam_topk(float *att, int V, int blank, int S, int64_t *__restrict__ ids,
                                                          float *__restrict__ cand_att) {
    __shared__ Cand wl[BEAM_NW * BEAM_MAXK];
    __shared__ Cand top[BEAM_MAXK];
    __shared__ float kth[BEAM_NW];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int row = blockIdx.x;
    const float NEG = -INFINITY;
    float ls[KL];
    int li[KL];
#pragma unroll
    for (int j = 0; j < KL; ++j) ls[j] = NEG, li[j] = 0x7fffffff;
    float thr = NEG;
    const int thr_lane = (S - 1) & 31, thr_list = (S - 1) >> 5;
    float *a = att + (size_t)row * V;
    const int per_warp = (((V + BEAM_NW - 1) / BEAM_NW + 31) / 32) * 32;
    const int s0 = min(V, wid * per_warp), e0 = min(V, s0 + per_warp);
    constexpr int U = 8;
    for (int vb0 = s0; vb0 < e0; vb0 += 32 * U) {
        float cu[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = vb0 + u * 32 + lane;
            // -inf (a token masked by another processor) still ranks, below everything finite, so that S ids always exist
            cu[u] = i < e0 ? (i == blank ? LZ : fmaxf(a[i], -3.0e38f)) : NEG;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float c = cu[u];
            unsigned m = __ballot_sync(0xffffffffu, c > thr);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const float cs = __shfl_sync(0xffffffffu, c, src);
                if (cs > thr) {  // ids are visited in increasing order: on ties the lower id stays
                    warp_list_insert<KL>(ls, li, cs, vb0 + u * 32 + src, lane);
                    thr = __shfl_sync(0xffffffffu, (KL == 2 && thr_list) ? ls[KL - 1] : ls[0], thr_lane);
                }
            }
        }
    }
    Cand *mine = wl + wid * BEAM_MAXK;
#pragma unroll
    for (int j = 0; j < KL; ++j) mine[j * 32 + lane].s = ls[j], mine[j * 32 + lane].i = li[j];
    if (KL == 1) mine[32 + lane].s = NEG, mine[32 + lane].i = 0x7fffffff;
    if (lane == 0) kth[wid] = thr;
    for (int k = tid; k < BEAM_MAXK; k += BEAM_NT) top[k].s = NEG, top[k].i = 0x7fffffff;
    __syncthreads();
    float bound = kth[0];
    for (int q = 1; q < BEAM_NW; ++q) bound = fmaxf(bound, kth[q]);
    rank_select(wl, BEAM_NW * BEAM_MAXK, S, bound, top);
    __syncthreads();
    if (tid < S) {
        const int id = top[tid].i;
        ids[(size_t)row * S + tid] = id;
        cand_att[(size_t)row * S + tid] = id == blank ? LZ : a[id];
    }
    __syncthreads();
    if (tid == 0) a[blank] = LZ;
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

// Warp = (hypothesis h, chunk of 32 candidates).  The warp first copies the hypothesis' lin column into shared memory
// (it is strided by HWP in the workspace), then every lane streams its token's time series with 128-bit loads:
// acc = sum_t lin[t] * exp(x[t]), ONE accumulator walked in frame order -- the same operations in the same order as a
// lane of k_psi_full, so a candidate's scores are bit-identical to the full-vocabulary step's.
__global__ void __launch_bounds__(128) k_psi_cand(const CandArgs a) {
    extern __shared__ __align__(16) float lin_s[];  // [warps per CTA][ldt]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nchunk = (a.S + 31) >> 5;
    const int item = blockIdx.x * (blockDim.x >> 5) + wid;
    if (item >= a.B * a.W * nchunk) return;
    const int h = item / nchunk, chunk = item - h * nchunk;
    const int b = h / a.W, w = h - b * a.W;
    const int g = w / a.HW, hh = w - g * a.HW;
    float *ls = lin_s + (size_t)wid * a.ldt;
    const float *lsrc = a.lin + ((size_t)(b * a.G + g) * a.Tpad) * a.HWP + hh;
    for (int t = lane; t < a.ldt; t += 32) ls[t] = lsrc[(size_t)t * a.HWP];  // ldt <= Tpad; entries >= T are zero
    __syncwarp();
    const int s = chunk * 32 + lane;
    if (s >= a.S) return;
    const long long v = a.ids[(size_t)h * a.S + s];
    const int start = a.ol > 1 ? a.ol : 1;
    float acc = 0.f, x0 = LZ;
    if (v >= 0 && v < a.V) {
        const float4 *xr = reinterpret_cast<const float4 *>(a.xt + ((size_t)b * a.V + (size_t)v) * a.ldt);
        const float4 *l4 = reinterpret_cast<const float4 *>(ls);
        if (a.ol == 0) x0 = a.xt[((size_t)b * a.V + (size_t)v) * a.ldt];
        const int q0 = (a.ol == 0 ? 0 : start) >> 2, q1 = a.ldt >> 2;
#pragma unroll 4
        for (int q = q0; q < q1; ++q) {
            const float4 xv = __ldg(xr + q);
            const float4 l = l4[q];
            acc = fmaf(l.x, ex2_approx(xv.x * LOG2E), acc);
            acc = fmaf(l.y, ex2_approx(xv.y * LOG2E), acc);
            acc = fmaf(l.z, ex2_approx(xv.z * LOG2E), acc);
            acc = fmaf(l.w, ex2_approx(xv.w * LOG2E), acc);
        }
        if (a.last_ids[h] == v) acc = a.psic[h];  // phi = r_prev blank there (:117-124)
    }
    EpiArgs e;
    e.Gmax = a.Gmax, e.s_prev = a.s_prev, e.s_rs = 1, e.s_cs = 0, e.att = nullptr, e.omw = a.omw, e.w = a.w;
    e.log_psi = nullptr, e.token_scores = nullptr, e.joint = nullptr, e.V = a.V, e.blank = a.blank, e.ol = a.ol;
    const float sp_row = a.s_prev != nullptr ? a.s_prev[h] : 0.f;
    const float av_in = a.cand_att != nullptr ? a.cand_att[(size_t)h * a.S + s] : 0.f;
    float lp, ts, jt, av;
    epi_lane(e, h, (int)v, acc, a.Gmax[h], x0, sp_row, av_in, lp, ts, jt, av);
    if (!(v >= 0 && v < a.V)) lp = LZ, ts = LZ, jt = LZ;
    const size_t o = (size_t)h * a.S + s;
    a.cand_log_psi[o] = lp;
    if (a.cand_ts != nullptr) a.cand_ts[o] = ts;
    if (a.cand_joint != nullptr) a.cand_joint[o] = jt;
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
