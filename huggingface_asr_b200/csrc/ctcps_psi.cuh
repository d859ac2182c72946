// ctcps_psi.cuh -- the lazy-state scoring kernel (K-b without the state write) and its fused per-tile top-2W epilogue.
// Included by ctcps_kernels.cu inside its anonymous namespace (one translation unit, one library).
//
// Lazy-state variant of K-b ("survivor recompute"): log_psi -- and with it the token and joint
// scores -- depends only on phi = f(r_prev) and x, NOT on the new forward variables r (ctc_scorer.py:154,164-167).  The
// only consumer of r is index_select_state, which keeps W columns per utterance out of W*V.  So in lazy mode the
// (T,2,BW,V) state is never written: k_psi_full computes the scores (reads x once: 4*T*B*V bytes instead of writing
// 8*T*BW*V), and k_select_lazy_* re-runs the recursion for the BW surviving (hyp, token) columns only.
//
//   psi[h, v] = sum_t lin[t, h] * exp(x[t, b, v]),   lin[t, h] = exp(r_sum[t-1, h] - G_h)  (zero outside the summed frames)
//
// is a skinny (W x T) x (T x V) product per utterance with exp() applied on the fly to the streamed operand; it is bound
// by reading x once from HBM, so it stays on the FP32 pipes (the 1e-4 log-space tolerance needs the fp32 mantissa).
//
// Round-2 structure (what changed against round 1 and why, profiles/r2d_psi_ab.md):
//   * packed FMAs pair two HYPOTHESES of one token (acc = (h0, h1) += (lin_h0, lin_h1) * (p, p)): the lin pairs come out
//     of shared memory already packed (LDS.128 = two pairs) and only the 4 token probabilities are duplicated per frame,
//     instead of one register move per hypothesis (10 at W = 10, 20 at W = 20);
//   * up to 20 hypotheses per thread: a W = 20 tile is no longer processed by two hypothesis groups that each re-read the
//     x tile from L2 and recompute exp(x) (round 1: C4 was MUFU / issue-bound at 0.46 of the HBM peak);
//   * thread 0 still issues the TMA loads (a refill by whichever warp releases a stage last, through a shared-memory ticket,
//     was measured: the ticket's return value stalls every warp once per chunk -- slower on the small shapes), but it now
//     walks the chunk sequence with incremental cursors (round 1 paid three integer divisions per chunk) and keeps a second
//     cursor further ahead that prefetches chunks into L2 (UTMAPF): ncu showed 24 % of the warp samples spinning on the
//     `full` barrier -- the ring (197 KB per SM, two thirds of it in flight at best) is too short a queue for HBM;
//   * TOPK mode (native decode loop): the epilogue ranks the tile's 512 x HW joint scores (+ running beam scores) and
//     publishes the tile's best 2W candidates; the (BW,V) joint tensor is never written or re-read (log_psi still is
//     written: the state selection reads W of its entries per utterance), and the separate per-row top-2W kernel of
//     round 1 disappears.  The beam step merges nvt x G short sorted lists.

struct PsiTopk {
    const float *beam_scores;  // (BW) running beam scores, added to the joint scores for ranking
    float2 *lists;             // [B][nvt*G][K]: (key, dense index hyp*V+tok as int bits), best first
    int K;                     // 2W
    const unsigned char *done; // (B) or null: utterances whose beam search has finished -- their tiles are not streamed at all
};

struct PsiArgs {
    const float *lin;    // (B*G, Tpad, HWP) exp(r_sum[t-1] - Gm), zero outside the summed range
    const float *Gmax;   // (BW)
    const float *psic;   // (BW) linear-domain sum for the column of the last label (phi = r_prev blank there)
    const float *s_prev;
    long long s_rs, s_cs;
    const int64_t *last_ids;
    float *att;
    float omw, w;
    float *log_psi, *token_scores, *joint;
    int B, W, T, V, blank, ol, G, Tpad, nvt;
    int prefetch;        // chunks of L2 look-ahead beyond the shared-memory ring (0 = none)
    const int64_t *xlens; // (B) utterance lengths or null: frames past the length hold exp(x) == 0 and are not streamed
    const int2 *frange;   // (B*G) or null: first / last 8-frame chunk in which any hypothesis of the group has lin != 0 (k_lin_range)
    unsigned long long *counter;  // or null: every launch adds the number of chunks it streams (instrumentation of bench.py)
    int evict_first;      // 1: the x stream is loaded with L2 evict-first priority
    PsiTopk tk;
};

constexpr int PSI_NT = 128;      // threads of the lazy scoring kernel (512-token tiles)
constexpr int PSI_MAXT = 256;      // tiles of a CTA whose chunk range is looked up once at kernel start (the rest: on the fly)
constexpr int PSI_TOPK_CAP = 256;  // elements of a tile at or above the threshold that the exact ranking takes

template <int HWP, int NT, int NSTAGE>
struct PsiSmem {
    static constexpr int VTILE = NT * 4;
    static constexpr int NBOX = VTILE / BOXC;
    alignas(128) float xs[NSTAGE][NBOX][TT][BOXC];
    alignas(16) float lin[NSTAGE][TT][HWP];
    alignas(8) uint64_t full[NSTAGE];
    alignas(8) uint64_t empty[NSTAGE];
    int cursor[2][8];  // thread 0's two look-ahead cursors (kept here, not in registers: every thread would pay for them)
    short tfirst[PSI_MAXT], tcnt[PSI_MAXT];  // first chunk / number of chunks of this CTA's first PSI_MAXT tiles (tile_chunks, once per launch)
};

// one warp per (padded) hypothesis: lin stream, Gmax and the last-label column sum
__global__ void __launch_bounds__(128) k_prep_psi(const float *__restrict__ r_prev, const XView x,
                                                  const int64_t *__restrict__ last_ids, int B, int W, int T, int V, int HW,
                                                  int HWP, int G, int start, int Tpad, float *__restrict__ lin,
                                                  float *__restrict__ Gmax, float *__restrict__ psic) {
    const int lane = threadIdx.x & 31;
    const int hp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (hp >= B * G * HWP) return;
    const int BW = B * W;
    const int b = hp / (G * HWP);
    const int rem = hp - b * (G * HWP);
    const int g = rem / HWP, hh = rem - g * HWP;
    const int w = g * HW + hh;
    const bool valid = hh < HW && w < W;
    const int h = b * W + w;
    float gm = -INFINITY;
    if (valid)
        for (int t = lane; t < T; t += 32)
            if (t >= start - 1 && t <= T - 2)
                gm = fmaxf(gm, lse2_precise(r_prev[((size_t)t * 2 + 0) * BW + h], r_prev[((size_t)t * 2 + 1) * BW + h]));
    gm = warp_max(gm);
    if (!(gm > -INFINITY)) gm = 0.f;
    float *base = lin + ((size_t)(b * G + g) * Tpad) * HWP + hh;
    long long c = valid ? last_ids[h] : -1;
    if (c < 0 || c >= V) c = -1;
    float sc = 0.f;
    for (int te = lane; te < Tpad; te += 32) {
        float e = 0.f;
        const int f = te - 1;
        if (valid && f >= start - 1 && f <= T - 2) {
            const float a = r_prev[((size_t)f * 2 + 0) * BW + h], cb = r_prev[((size_t)f * 2 + 1) * BW + h];
            e = expf(lse2_precise(a, cb) - gm);
            if (c >= 0) sc += expf(cb - gm) * expf(x.at(b, te, c));
        }
        base[(size_t)te * HWP] = e;
    }
    sc = warp_sum(sc);
    if (valid && lane == 0) {
        Gmax[h] = gm;
        psic[h] = sc;
    }
}

// The lin stream is sparse in time: lin[t, h] = exp(r_sum[t-1, h] - offset) underflows to exactly 0 wherever the prefix is more
// than e^-87 (e^-103 with subnormals) less probable than at its best frame -- far behind and far ahead of where the prefix
// ends in the audio.  A chunk in which lin is 0 for every hypothesis of the group adds lin * p = 0 to every sum, bit for bit
// (p = exp(x) <= 1 is finite), so the scoring kernel need not stream it.  One CTA per (utterance, hypothesis group) finds
// the first and last chunk with a nonzero entry; {1, 0} = none.
__global__ void __launch_bounds__(128) k_lin_range(const float *__restrict__ lin, int Tpad, int HWP, int2 *__restrict__ frange) {
    __shared__ int s_lo[4], s_hi[4];
    const float *base = lin + (size_t)blockIdx.x * Tpad * HWP;
    int lo = 0x7fffffff, hi = -1;
    const int n4 = HWP >> 2;  // HWP % 4 == 0: a frame's row is n4 float4
    for (int te = threadIdx.x; te < Tpad; te += 128) {
        const float4 *row = reinterpret_cast<const float4 *>(base + (size_t)te * HWP);
        bool nz = false;
        for (int q = 0; q < n4; ++q) {
            const float4 v = row[q];
            nz = nz || v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f;
        }
        if (nz) lo = min(lo, te), hi = max(hi, te);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) s_lo[threadIdx.x >> 5] = lo, s_hi[threadIdx.x >> 5] = hi;
    __syncthreads();
    if (threadIdx.x == 0) {
        lo = min(min(s_lo[0], s_lo[1]), min(s_lo[2], s_lo[3]));
        hi = max(max(s_hi[0], s_hi[1]), max(s_hi[2], s_hi[3]));
        frange[blockIdx.x] = hi >= 0 ? make_int2(lo / TT, hi / TT) : make_int2(1, 0);
    }
}

// ------------------------------------------------------------------------------------------
// TOPK epilogue: the tile's K best (joint score + running beam score) out of 512 tokens x HW hypotheses, ranked exactly
// by (score descending, dense index ascending) -- the order of the beam step (and of torch.topk on distinct scores).
//
//   pass 1  per hypothesis pair (rolled loop over the rows of the sums, two rows per iteration): log_psi -- streamed to the
//           dense (BW,V) log_psi tensor, which is all the next state selection reads (:193) -- then token score, joint
//           score and key = joint + beam score; the key replaces the sum in its register (the rows rotate through
//           positions 0 and 1, so after HW / 2 iterations they are back in place); every thread keeps the max of its keys;
//   tau     = the K-th largest of the NT thread maxima: every warp sorts its 32 maxima (bitonic, shuffles only) and
//           publishes them; a thread ranks its value among the other warps' sorted lists by binary search.  At least K
//           elements of the tile are >= tau, so the K best are, and only a few dozen elements in all;
//   pass 2  compares the 4 * HW keys a thread holds with tau (straight-line code on registers) and appends the few
//           survivors to a shared list, which is then ranked exactly by counting.
// Round 2's first version kept log_psi in the registers and recomputed the keys in pass 2 from a second read of the decoder
// scores (so that nothing dense was written at all): 1.8 k extra instructions per thread and tile, +31 us per C2 launch
// (profiles/r2c_psi_ncu.md) -- more than the 8 us the 51 MB log_psi write costs.
// A tile where more than PSI_TOPK_CAP elements reach tau (fewer than K threads hold a valid score, or a constant row)
// falls back to K rounds of block-wide argmax over the keys.  Invalid positions (v >= V, padded hypotheses) carry -inf keys
// with index INT_MAX and never enter; a list with fewer than K valid entries is closed with such sentinels.
// ------------------------------------------------------------------------------------------
struct PsiTopkSmem {
    float sorted[PSI_NT];             // per warp: its 32 thread maxima, descending
    float2 cand[PSI_TOPK_CAP];
    float2 red[PSI_NT / 32];
    float2 winner;
    float tau;
    unsigned int n_cand;
};

template <int HW, int NWARP, bool TOPK>
struct PsiTopkScratch {};
template <int HW, int NWARP>
struct PsiTopkScratch<HW, NWARP, true> {
    PsiTopkSmem t;
    float scal[NWARP * 3 * HW];  // per warp: Gmax, s_prev, beam score of the tile's hypotheses
};
template <int HW, int HWP, int NT, int NSTAGE, bool TOPK>
struct PsiSmemAll {
    PsiSmem<HWP, NT, NSTAGE> pipe;
    PsiTopkScratch<HW, NT / 32, TOPK> topk;
};

__device__ __forceinline__ bool key_beats(float s1, int i1, float s2, int i2) { return s1 > s2 || (s1 == s2 && i1 < i2); }

// second half of epi_lane: log_psi -> token score -> joint score (:175-176, :325, :332)
__device__ __forceinline__ float joint_of(const EpiArgs &e, int v, float lp, float sp, float av) {
    float ts = lp - sp;
    if (ts == 0.f) ts = LZ;
    if (v == e.blank) av = LZ;
    return __fadd_rn(__fmul_rn(e.omw, av), __fmul_rn(e.w, ts));
}
// first half: linear-domain sum -> log_psi (:164-173)
__device__ __forceinline__ float log_psi_of(const EpiArgs &e, int v, float S, float gm, float x0) {
    float lp = gm + logf(S);
    if (!(lp > LZ)) lp = LZ;
    if (e.ol == 0) lp = lse2_precise(lp, x0);
    if (v == e.blank) lp = LZ;
    return lp;
}

template <int HW>
__device__ __forceinline__ void epilogue_topk(const EpiArgs &e, const PsiTopk &tk, PsiTopkSmem &ts, float (&S)[HW][4], const float (&x0)[4],
                                              int hrow0 /* global row of the tile's first hypothesis */,
                                              int w0 /* its index inside the utterance */, int nhyp, int v0,
                                              float2 *__restrict__ list_out, float *scal /* this warp's [3][HW] shared scratch */) {
    static_assert(HW % 2 == 0, "rows are processed in pairs");
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int V = e.V, K = tk.K;
    const bool col_ok = v0 < V;  // V % 4 == 0 on this path: the thread's 4 tokens are valid together
    const float NEG = -INFINITY;
    // per-hypothesis scalars of the tile, staged by each warp for itself (no CTA barrier; read back as LDS broadcasts)
    if (lane < HW) {
        const int h = hrow0 + (lane < nhyp ? lane : 0);
        scal[0 * HW + lane] = e.Gmax[h];
        scal[1 * HW + lane] = e.s_prev != nullptr ? e.s_prev[(long long)h * e.s_rs] : 0.f;
        scal[2 * HW + lane] = tk.beam_scores[h];
    }
    if (tid == 0) ts.n_cand = 0;
    __syncwarp();

    constexpr int AHEAD = (HW < 4 || HW > 12) ? 2 : 4;  // decoder-score rows in flight (even); wide groups have no registers to spare
    float4 attv[AHEAD];
    auto att_row = [&](int hh) {
        return (col_ok && hh < nhyp) ? __ldg(reinterpret_cast<const float4 *>(e.att + (size_t)(hrow0 + hh) * V + v0)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    // ---- pass 1: sums -> log_psi (streamed out) -> keys (kept in S), thread max of the keys ------------------------
    float mx = NEG;
#pragma unroll
    for (int i = 0; i < AHEAD; ++i) attv[i] = att_row(i);
#pragma unroll 1
    for (int hh = 0; hh < HW; hh += 2) {
        float key[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float gm = scal[0 * HW + hh + r], sp = scal[1 * HW + hh + r], bm = scal[2 * HW + hh + r];
            const float av[4] = {attv[r].x, attv[r].y, attv[r].z, attv[r].w};
            const bool ok = col_ok && hh + r < nhyp;
            float lp[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                lp[j] = log_psi_of(e, v0 + j, S[r][j], gm, x0[j]);
                key[r][j] = ok ? joint_of(e, v0 + j, lp[j], sp, av[j]) + bm : NEG;
                mx = fmaxf(mx, key[r][j]);
            }
            if (ok) __stcs(reinterpret_cast<float4 *>(e.log_psi + (size_t)(hrow0 + hh + r) * V + v0), make_float4(lp[0], lp[1], lp[2], lp[3]));
        }
#pragma unroll
        for (int i = 0; i + 2 < AHEAD; ++i) attv[i] = attv[i + 2];
        if (AHEAD >= 2) {
            attv[AHEAD - 2] = att_row(hh + AHEAD);
            attv[AHEAD - 1] = att_row(hh + AHEAD + 1);
        }
#pragma unroll
        for (int i = 0; i + 2 < HW; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) S[i][j] = S[i + 2][j];
#pragma unroll
        for (int j = 0; j < 4; ++j) S[HW - 2][j] = key[0][j], S[HW - 1][j] = key[1][j];
    }
    // ---- tau: the K-th largest thread maximum --------------------------------------------------------------------
    {
        float v = mx;  // descending bitonic sort of the warp's 32 maxima: lane p ends up with the p-th largest
#pragma unroll
        for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                const float other = __shfl_xor_sync(0xffffffffu, v, j);
                const bool keep_max = ((lane & j) == 0) == ((lane & k) == 0);
                v = keep_max ? fmaxf(v, other) : fminf(v, other);
            }
        ts.sorted[tid] = v;
        __syncthreads();
        // rank of (warp wid, position lane) in the order: value descending, then warp, then position
        int rank = lane;
#pragma unroll
        for (int ow = 0; ow < PSI_NT / 32; ++ow) {
            if (ow == wid) continue;
            const float *lst = ts.sorted + ow * 32;
            // number of entries of lst that come before v: entries > v, plus entries == v when ow < wid
            int lo = 0, hi = 32;
#pragma unroll
            for (int it = 0; it < 6; ++it) {
                if (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    const float m = lst[mid];
                    const bool before = m > v || (m == v && ow < wid);
                    if (before) lo = mid + 1;
                    else hi = mid;
                }
            }
            rank += lo;
        }
        if (rank == K - 1) ts.tau = v;  // ranks are a permutation of 0..NT-1: exactly one writer
        __syncthreads();
    }
    const float tau = ts.tau;
    // ---- pass 2: the keys >= tau into the shared list ---------------------------------------------------------------
#pragma unroll
    for (int hh = 0; hh < HW; ++hh)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (S[hh][j] >= tau && col_ok && hh < nhyp) {
                const unsigned pos = atomicAdd(&ts.n_cand, 1u);
                if (pos < PSI_TOPK_CAP) ts.cand[pos] = make_float2(S[hh][j], __int_as_float((w0 + hh) * V + v0 + j));
            }
    __syncthreads();
    const unsigned nc = ts.n_cand;
    const float2 sentinel = make_float2(NEG, __int_as_float(0x7fffffff));
    if (nc <= PSI_TOPK_CAP) {
        for (unsigned q = tid; q < nc; q += PSI_NT) {
            const float2 me = ts.cand[q];
            int rank = 0;
            for (unsigned o = 0; o < nc; ++o) {
                const float2 c = ts.cand[o];
                rank += key_beats(c.x, __float_as_int(c.y), me.x, __float_as_int(me.y)) ? 1 : 0;
            }
            if (rank < K) list_out[rank] = me;
        }
        for (int r = (int)nc + tid; r < K; r += PSI_NT) list_out[r] = sentinel;
        __syncthreads();  // the list and the counters are reused by the next tile
        return;
    }
    // ---- fallback: K rounds of block-wide argmax over the keys that come after the previous winner ------------------
    float pk = INFINITY;
    int pi = -1;
    for (int r = 0; r < K; ++r) {
        float bk = NEG;
        int bi = 0x7fffffff;
#pragma unroll
        for (int hh = 0; hh < HW; ++hh)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = (w0 + hh) * V + v0 + j;
                const float key = S[hh][j];
                if (col_ok && hh < nhyp && key_beats(pk, pi, key, idx) && (bi == 0x7fffffff || key_beats(key, idx, bk, bi))) bk = key, bi = idx;
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ok = __shfl_xor_sync(0xffffffffu, bk, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi != 0x7fffffff && (bi == 0x7fffffff || key_beats(ok, oi, bk, bi))) bk = ok, bi = oi;
        }
        if (lane == 0) ts.red[wid] = make_float2(bk, __int_as_float(bi));
        __syncthreads();
        if (tid == 0) {
            float2 best = ts.red[0];
            for (int q = 1; q < PSI_NT / 32; ++q) {
                const float2 c = ts.red[q];
                const int ci = __float_as_int(c.y), bi2 = __float_as_int(best.y);
                if (ci != 0x7fffffff && (bi2 == 0x7fffffff || key_beats(c.x, ci, best.x, bi2))) best = c;
            }
            if (__float_as_int(best.y) == 0x7fffffff) best = sentinel;
            ts.winner = best;
            list_out[r] = best;
        }
        __syncthreads();
        const float2 wv = ts.winner;
        pk = wv.x, pi = __float_as_int(wv.y);
        if (pi == 0x7fffffff) {  // nothing left: close the list
            for (int q = r + 1 + tid; q < K; q += PSI_NT) list_out[q] = sentinel;
            break;
        }
    }
    __syncthreads();
}

// Persistent kernel: grid = #SMs x resident CTAs, CTA i walks tiles i, i+grid, ... (tile = (utterance, 512-token tile,
// hypothesis group)).  The chunks of all its tiles form one flat sequence kept NSTAGE stages ahead of the consumers with
// TMA; stages are handed back through `empty` mbarriers (one arrival per warp), so the warps of a CTA never meet at a
// CTA-wide barrier inside the stream.  The lin stream is zero outside the summed frame range, which makes the inner loop
// branch-free: 8 frames x (1 LDS.128 of x, 4 FMUL + 4 ex2, HWP/4 LDS.128 of lin, 2*HW FFMA2 with a scalar operand).
template <int HW, int HWP, int NT, int MINB, int NSTAGE, bool TOPK>
__global__ void __launch_bounds__(NT, MINB) k_psi_full(const __grid_constant__ CUtensorMap tmx, const PsiArgs a) {
    static_assert(HW % 2 == 0 && HW <= HWP && HWP % 4 == 0, "hypotheses are processed in pairs");
    static_assert(!TOPK || NT == PSI_NT, "the top-k scratch is sized for PSI_NT threads");
    using Smem = PsiSmem<HWP, NT, NSTAGE>;
    using SmemAll = PsiSmemAll<HW, HWP, NT, NSTAGE, TOPK>;
    constexpr int VTILE = Smem::VTILE;
    constexpr int NBOX = Smem::NBOX;
    constexpr int NWARP = NT / 32;
    constexpr int HP = HW / 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SmemAll &sma = *reinterpret_cast<SmemAll *>(smem_raw);
    Smem &sm = sma.pipe;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    (void)wid;
    const int T = a.T, V = a.V, W = a.W;
    const int start = a.ol > 1 ? a.ol : 1;
    const int c0 = (a.ol == 0 ? 0 : start) / TT;
    const int cN = (T - 1) / TT;
    const int nchunk = cN - c0 + 1;
    const int ntiles = a.B * a.nvt * a.G;
    const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    constexpr uint32_t STAGE_BYTES = TT * VTILE * 4 + TT * HWP * 4;

    auto decode_tile = [&](int tile, int &b, int &vt, int &g) {
        g = tile % a.G;
        tile /= a.G;
        vt = tile % a.nvt;
        b = tile / a.nvt;
    };
    // TOPK mode (a decode loop that owns its beam search): the tiles of an utterance whose search has finished are skipped --
    // producer and consumers agree on that from the same `done` flags, which do not change while the kernel runs.  HF's
    // loop scores such rows until the whole batch is done and throws the result away; under the synthetic decoder the last
    // two steps of a decode (every utterance finished, all-tie scores: the epilogue's slow path) cost 1.7x a normal step.
    auto tile_skipped = [&](int b) -> bool {
        if constexpr (TOPK) return a.tk.done != nullptr && a.tk.done[b] != 0;
        return false;
    };
    // Chunks of a tile that are streamed: count -1 = the tile is skipped altogether (no epilogue either), else `count` chunks from
    // `first`: from c0 (or the group's first chunk with a nonzero lin) up to the chunk that holds the utterance's last frame (or the
    // group's last chunk with a nonzero lin).  Past the length x is logzero for every token but blank (:39-42), exp(x) is
    // exactly 0 and lin * 0 + acc == acc bit for bit; the blank column's score is overwritten with logzero anyway (:173).
    auto tile_chunks = [&](int b, int g, int &first) -> int {
        first = c0;
        if (tile_skipped(b)) return -1;
        int last = cN;
        if (a.xlens != nullptr) {
            const long long lraw = a.xlens[b];
            long long l = lraw < 0 ? lraw + T : lraw;  // the length as K-a applied it
            if (l < 0) l = 0;
            if (lraw >= T) l = T;
            const int ll = l > 0 ? (int)((l - 1) / TT) : -1;
            last = ll < last ? ll : last;
        }
        if (a.frange != nullptr) {
            const int2 r = a.frange[b * a.G + g];
            last = r.y < last ? r.y : last;
            if (a.ol != 0) first = r.x > first ? r.x : first;  // the first step reads x[0] out of chunk 0 (:112-113)
            else if (last < 0) last = 0;
        }
        const int n = last - first + 1;
        return n > 0 ? n : 0;
    };
    // The chunk range of a tile hangs on three words of global memory (done flag, length, lin range): looked up for all tiles
    // of the CTA at once, by a thread each, and kept in shared memory -- one round trip per launch.  (Looked up where they were
    // needed they were a chain of dependent loads in front of every tile, for the producer and again for the consumers: 8 % of
    // the warp samples before the first TMA issue plus ~1.5 us per tile, ncu r3m.)
    const bool clipped = (TOPK && a.tk.done != nullptr) || a.xlens != nullptr || a.frange != nullptr;
    const bool table = clipped && cN < 32767;  // chunk indices fit the 16-bit table
    if (table) {
        for (int ti = tid; ti < my_tiles && ti < PSI_MAXT; ti += NT) {
            int b, vt, g, first;
            decode_tile((int)blockIdx.x + ti * (int)gridDim.x, b, vt, g);
            const int n = tile_chunks(b, g, first);
            sm.tfirst[ti] = (short)first, sm.tcnt[ti] = (short)n;
        }
        __syncthreads();
    }
    auto tile_info = [&](int ti, int b, int g, int &first) -> int {
        if (!clipped) {
            first = c0;
            return nchunk;
        }
        if (table && ti < PSI_MAXT) {
            first = sm.tfirst[ti];
            return sm.tcnt[ti];
        }
        return tile_chunks(b, g, first);
    };
    int nitems = my_tiles * nchunk;  // used by thread 0 only
    if (clipped && tid == 0) {
        nitems = 0;
        for (int ti = 0; ti < my_tiles; ++ti) {
            int b, vt, g, first;
            if (!table || ti >= PSI_MAXT) decode_tile((int)blockIdx.x + ti * (int)gridDim.x, b, vt, g);
            else b = vt = g = 0;
            const int n = tile_info(ti, b, g, first);
            nitems += n > 0 ? n : 0;
        }
    }
    if (a.counter != nullptr && tid == 0) atomicAdd(a.counter, (unsigned long long)nitems);
    // Thread 0 walks the CTA's flat chunk sequence twice ahead of the consumers, with two cursors that advance
    // incrementally (a tile is decoded once, when a cursor enters it -- no division per chunk):
    //   `is`  the next item to load into the shared-memory ring (NSTAGE items ahead of the one being consumed);
    //   `pf`  the next item to prefetch into L2 with TMA (a.prefetch further items ahead).  The ring bounds what can
    //         be in flight towards shared memory (4 CTAs x 3 stages x 16 KB per SM, a third of it being read at any time);
    //         the L2 prefetches carry the rest of the HBM queue depth, and the ring's own loads then mostly hit L2.
    struct Cursor {
        int k, ti, ci, b, vt, g, nch, cs;
    };
    auto cursor_load = [&](int which) {
        const int *p = sm.cursor[which];
        return Cursor{p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]};
    };
    auto cursor_store = [&](int which, const Cursor &c) {
        int *p = sm.cursor[which];
        p[0] = c.k, p[1] = c.ti, p[2] = c.ci, p[3] = c.b, p[4] = c.vt, p[5] = c.g, p[6] = c.nch, p[7] = c.cs;
    };
    auto cursor_enter = [&](Cursor &c) {  // decode the cursor's tile, stepping over tiles that stream nothing
        while (c.ti < my_tiles) {
            decode_tile((int)blockIdx.x + c.ti * (int)gridDim.x, c.b, c.vt, c.g);
            c.nch = tile_info(c.ti, c.b, c.g, c.cs);
            if (c.nch > 0) break;
            ++c.ti;
        }
    };
    auto cursor_next = [&](Cursor &c) {
        ++c.k;
        if (++c.ci == c.nch) {
            c.ci = 0;
            ++c.ti;
            cursor_enter(c);
        }
    };
    const uint64_t l2_stream_policy = l2_policy_evict_first();
    auto issue = [&](const Cursor &c) {  // chunk c.cs + c.ci of tile (c.b, c.vt, c.g) into stage c.k % NSTAGE
        const int s = c.k % NSTAGE;
        const int ch = c.cs + c.ci;
        mbar_expect_tx(&sm.full[s], STAGE_BYTES);
        if (a.evict_first) {
#pragma unroll
            for (int bx = 0; bx < NBOX; ++bx)
                tma_load_2d_hint(&sm.xs[s][bx][0][0], &tmx, c.vt * VTILE + bx * BOXC, c.b * T + ch * TT, &sm.full[s], l2_stream_policy);
        } else {
#pragma unroll
            for (int bx = 0; bx < NBOX; ++bx)
                tma_load_2d(&sm.xs[s][bx][0][0], &tmx, c.vt * VTILE + bx * BOXC, c.b * T + ch * TT, &sm.full[s]);
        }
        bulk_load_1d(&sm.lin[s][0][0], a.lin + ((size_t)(c.b * a.G + c.g) * a.Tpad + (size_t)ch * TT) * HWP, TT * HWP * 4, &sm.full[s]);
    };
    auto prefetch = [&](const Cursor &c) {
        const int ch = c.cs + c.ci;
#pragma unroll
        for (int bx = 0; bx < NBOX; ++bx) tma_prefetch_2d(&tmx, c.vt * VTILE + bx * BOXC, c.b * T + ch * TT);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) mbar_init(&sm.full[s], 1), mbar_init(&sm.empty[s], NWARP);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        Cursor is = {0, 0, 0, 0, 0, 0, 0, 0}, pf = {0, 0, 0, 0, 0, 0, 0, 0};
        if (nitems > 0) {
            cursor_enter(is);
            while (is.k < nitems && is.k < NSTAGE) {  // prologue: fill the ring
                issue(is);
                cursor_next(is);
            }
            pf = is;
            while (pf.k < nitems && pf.k < NSTAGE + a.prefetch) {  // and start the L2 look-ahead
                prefetch(pf);
                cursor_next(pf);
            }
        }
        cursor_store(0, is);
        cursor_store(1, pf);
    }
    __syncthreads();

    const int bx = (tid * 4) / BOXC, col = (tid * 4) % BOXC;
    int k = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
        int b, vt, g;
        decode_tile((int)blockIdx.x + ti * (int)gridDim.x, b, vt, g);
        int cfirst;
        const int nch = tile_info(ti, b, g, cfirst);
        if (nch < 0) continue;  // its candidate list keeps the last step's contents: the beam step ignores a finished utterance
        unsigned long long acc2[HP][4];  // (hyp 2p, hyp 2p+1) of token j, packed for FFMA2
        float x0[4];
#pragma unroll
        for (int hp = 0; hp < HP; ++hp)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc2[hp][j] = 0ull;
#pragma unroll
        for (int j = 0; j < 4; ++j) x0[j] = LZ;

        for (int ci = 0; ci < nch; ++ci, ++k) {
            const int s = k % NSTAGE;
            mbar_wait(&sm.full[s], (uint32_t)((k / NSTAGE) & 1));
            if (a.ol == 0 && ci == 0) {  // r[0,0] = x_[0,0] enters log_psi as its own term (:158,165)
                const float4 xv4 = *reinterpret_cast<const float4 *>(&sm.xs[s][bx][0][col]);
                x0[0] = xv4.x, x0[1] = xv4.y, x0[2] = xv4.z, x0[3] = xv4.w;
            }
#pragma unroll
            for (int tt = 0; tt < TT; ++tt) {
                const float4 xv4 = *reinterpret_cast<const float4 *>(&sm.xs[s][bx][tt][col]);
                unsigned long long pp[4];
                {
                    const float p0 = ex2_approx(xv4.x * LOG2E), p1 = ex2_approx(xv4.y * LOG2E);
                    const float p2 = ex2_approx(xv4.z * LOG2E), p3 = ex2_approx(xv4.w * LOG2E);
                    pp[0] = pack2(p0, p0), pp[1] = pack2(p1, p1), pp[2] = pack2(p2, p2), pp[3] = pack2(p3, p3);
                }
                unsigned long long lp2[HWP / 2];
#pragma unroll
                for (int q = 0; q < HWP / 4; ++q) {
                    const ulonglong2 l2 = *reinterpret_cast<const ulonglong2 *>(&sm.lin[s][tt][q * 4]);
                    lp2[q * 2 + 0] = l2.x, lp2[q * 2 + 1] = l2.y;
                }
#pragma unroll
                for (int hp = 0; hp < HP; ++hp)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc2[hp][j] = ffma2(lp2[hp], pp[j], acc2[hp][j]);
            }
            // hand the stage back (one arrival per warp); thread 0 refills it once every warp has, and moves the L2 look-ahead on
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[s]);
            if (tid == 0) {
                if (k + NSTAGE < nitems) {  // the item that reuses this stage
                    Cursor is = cursor_load(0);
                    mbar_wait(&sm.empty[s], (uint32_t)((k / NSTAGE) & 1));
                    issue(is);
                    cursor_next(is);
                    cursor_store(0, is);
                }
                if (a.prefetch > 0 && k + NSTAGE + a.prefetch < nitems) {
                    Cursor pf = cursor_load(1);
                    prefetch(pf);
                    cursor_next(pf);
                    cursor_store(1, pf);
                }
            }
        }

        const int v0 = vt * VTILE + tid * 4;
        const int w0 = g * HW;
        const int h0 = b * W + w0;
        const int nhyp = min(HW, W - w0);
        float acc[HW][4];
#pragma unroll
        for (int hp = 0; hp < HP; ++hp)
#pragma unroll
            for (int j = 0; j < 4; ++j) unpack2(acc2[hp][j], acc[2 * hp][j], acc[2 * hp + 1][j]);
        // the column of each hypothesis' last label sums r_prev_blank instead of r_sum: take it from k_prep_psi
#pragma unroll
        for (int hh = 0; hh < HW; ++hh) {
            if (hh < nhyp) {
                const int cj = (int)(a.last_ids[h0 + hh] - (long long)v0);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (cj == j) acc[hh][j] = a.psic[h0 + hh];
            }
        }
        EpiArgs e;
        e.Gmax = a.Gmax, e.s_prev = a.s_prev, e.s_rs = a.s_rs, e.s_cs = a.s_cs, e.att = a.att, e.omw = a.omw, e.w = a.w;
        e.log_psi = a.log_psi, e.token_scores = a.token_scores, e.joint = a.joint, e.V = V, e.blank = a.blank, e.ol = a.ol;
        if constexpr (TOPK) {
            float2 *list_out = a.tk.lists + ((size_t)(b * a.nvt + vt) * a.G + g) * a.tk.K;
            epilogue_topk<HW>(e, a.tk, sma.topk.t, acc, x0, h0, w0, nhyp, v0, list_out, sma.topk.scal + wid * 3 * HW);
        } else {
            epilogue_tile<HW>(e, acc, x0, h0, nhyp, v0);
        }
    }
}
