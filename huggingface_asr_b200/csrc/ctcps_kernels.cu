// ctcps_kernels.cu -- sm_100a kernels + C ABI of the CTC prefix scorer (see include/ctcps.h).
//
// Replaces, for BUTSpeechFIT/huggingface_asr, src/decoding/ctc_scorer.py:
//   K-a  k_init            log-softmax + length padding + blank column          (:279, :39-46)
//        k_initial_state   blank-only forward variable of the empty prefix      (:74-85)
//   K-b  k_prep            per-hypothesis phi stream for the recursion          (:115-124)
//        k_score_full      forward recursion over T x (hyp, token) + log_psi +
//                          token scores + joint combine, full vocabulary        (:98-178, :325, :332)
//        k_score_partial   same on scoring_ids candidates                       (:90-97, :155-162)
//   K-c  k_select          index_select_state gather                            (:180-207)
//        k_trick           eos/space trick                                      (:333-349)
//
// Design of K-b (the hot kernel; DESIGN.md has the roofline arithmetic):
//   * one CTA = (utterance b, tile of VTILE tokens, group of HW hypotheses); a thread owns
//     4 consecutive tokens x HW hypotheses = 4*HW independent recursion chains in registers,
//     so latency is hidden by ILP and every log-posterior x[t,b,v] fetched from shared memory
//     is reused HW times;
//   * x[t0:t0+TT, b, v-tile] is staged by TMA (cp.async.bulk.tensor.2d, mbarrier complete_tx)
//     through a ring of NS stages; the token-independent per-hypothesis stream
//     {r_sum, r_prev_blank, exp(r_sum-G), exp(r_prev_blank-G)}[t-1] + blank log-prob[t] comes in
//     with the same barrier as one 1-D bulk copy;
//   * the state r (T,2,BW,V) -- 8 bytes per lane-step, the HBM roofline of the path -- is written
//     with 128-bit streaming stores, V innermost, 512 B contiguous per warp per (t, plane, hyp);
//   * log_psi is accumulated in the linear domain against a per-hypothesis offset G (one FFMA
//     per lane-step instead of a logaddexp), the two logaddexp of the recursion use MUFU ex2/lg2.
//   No tensor cores: nothing here is a dense contraction.  No CPU fallback.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ctcps.h"

namespace {

constexpr float LZ = CTCPS_LOGZERO;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

constexpr int TT = 8;        // frames per pipeline stage
constexpr int NS = 3;        // pipeline stages
constexpr int BOXC = 256;    // TMA box width (elements), the hardware maximum per dimension
constexpr int MAX_HW = 5;    // hypotheses per thread

thread_local char g_errbuf[256];

// ------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// torch.logsumexp over two values, max + log(1 + exp(min - max)), on the MUFU pipe.
__device__ __forceinline__ float lse2_fast(float a, float b) {
    const float m = fmaxf(a, b);
    const float d = -fabsf(a - b);
    const float e = ex2_approx(d * LOG2E);
    return fmaf(lg2_approx(1.0f + e), LN2, m);
}
// Same with full-precision libm-grade functions (used off the hot loop).
__device__ __forceinline__ float lse2_precise(float a, float b) {
    const float m = fmaxf(a, b);
    return logf(expf(a - m) + expf(b - m)) + m;
}

// packed 2 x fp32 (sm_100: FFMA2 issues two FMAs per instruction slot)
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &a, float &b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// The log-posteriors in either HBM layout: frame-major (B,T,ldx) for the TMA-fed full-vocabulary kernels, or token-major
// (B,V,ldt) ("vt": a token's time series is contiguous) for the pre-beam kernels that gather a few tokens per hypothesis.
struct XView {
    const float *p;
    long long sb, st, sv;  // element (b,t,v) at p[b*sb + t*st + v*sv]
    __device__ __forceinline__ float at(int b, int t, long long v) const { return p[(long long)b * sb + (long long)t * st + v * sv]; }
};

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
// the same with an L2 eviction-priority hint (createpolicy): a stream far larger than L2 that is read once per launch should
// not push out what the launch re-reads (decoder scores, the lin stream) or what the next kernel reads (log_psi)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_2d_hint(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
// TMA prefetch of a box into L2 (no shared-memory destination, no barrier): SASS UTMAPF.L2
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide reductions for <= 1024 threads; `red` is 32 floats of shared memory
__device__ __forceinline__ float block_max(float v, float *red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    v = lane < nw ? red[lane] : -INFINITY;
    return warp_max(v);
}
__device__ __forceinline__ float block_sum(float v, float *red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    v = lane < nw ? red[lane] : 0.f;
    return warp_sum(v);
}

// ------------------------------------------------------------------------------------------
// K-a: log-softmax + length padding + blank column.  One CTA per (b, t) row.  The row is read ONCE into registers
// (float4 when the row is 16-byte aligned; up to K_INIT_MAXV columns), reduced, and written once: 8*V bytes per row.
// ------------------------------------------------------------------------------------------
constexpr int K_INIT_NT = 256;
constexpr int K_INIT_R4 = 8;                                // float4 per thread held in registers
constexpr int K_INIT_MAXV = K_INIT_NT * K_INIT_R4 * 4;      // 8192 columns; wider rows take the 3-pass path

__global__ void __launch_bounds__(K_INIT_NT) k_init(const float *__restrict__ in, int ld_in, const int64_t *__restrict__ lens,
                                                    int T, int V, int blank, int apply, float *out, int ldx,
                                                    float *__restrict__ blank_lp) {
    __shared__ float red[32];
    const int row = blockIdx.x;
    const int b = row / T, t = row - b * T;
    const float *src = in + (size_t)row * ld_in;
    float *dst = out + (size_t)row * ldx;
    if (lens != nullptr) {
        const long long lraw = lens[b];
        long long l = lraw < 0 ? lraw + T : lraw;  // python slice x[i, l:, :]
        if (l < 0) l = 0;
        if (lraw < T && t >= l) {  // ctc_scorer.py:39-42
            for (int v = threadIdx.x; v < V; v += blockDim.x) dst[v] = (v == blank) ? 0.f : LZ;
            if (blank_lp != nullptr && threadIdx.x == 0) blank_lp[row] = 0.f;
            return;
        }
    }
    if (!apply) {
        for (int v = threadIdx.x; v < V; v += blockDim.x) {
            const float o = src[v];
            if (dst != src) dst[v] = o;
            if (v == blank && blank_lp != nullptr) blank_lp[row] = o;
        }
        return;
    }
    const bool vec = V <= K_INIT_MAXV && (V & 3) == 0 && (ld_in & 3) == 0 && (ldx & 3) == 0 &&
                     ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (vec) {
        const int n4 = V >> 2;
        float4 v4[K_INIT_R4];
        float m = -INFINITY;
#pragma unroll
        for (int q = 0; q < K_INIT_R4; ++q) {
            const int i = threadIdx.x + q * K_INIT_NT;
            if (i < n4) {
                v4[q] = __ldcs(reinterpret_cast<const float4 *>(src) + i);
                m = fmaxf(m, fmaxf(fmaxf(v4[q].x, v4[q].y), fmaxf(v4[q].z, v4[q].w)));
            }
        }
        m = block_max(m, red);
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < K_INIT_R4; ++q) {
            const int i = threadIdx.x + q * K_INIT_NT;
            if (i < n4) s += expf(v4[q].x - m) + expf(v4[q].y - m) + expf(v4[q].z - m) + expf(v4[q].w - m);
        }
        s = block_sum(s, red);
        const float ls = logf(s);
#pragma unroll
        for (int q = 0; q < K_INIT_R4; ++q) {
            const int i = threadIdx.x + q * K_INIT_NT;
            if (i < n4) {
                const float4 o = make_float4((v4[q].x - m) - ls, (v4[q].y - m) - ls, (v4[q].z - m) - ls, (v4[q].w - m) - ls);
                reinterpret_cast<float4 *>(dst)[i] = o;
                if (blank_lp != nullptr && (blank >> 2) == i) blank_lp[row] = (&o.x)[blank & 3];
            }
        }
        return;
    }
    float m = -INFINITY;
    for (int v = threadIdx.x; v < V; v += blockDim.x) m = fmaxf(m, src[v]);
    m = block_max(m, red);
    float s = 0.f;
    for (int v = threadIdx.x; v < V; v += blockDim.x) s += expf(src[v] - m);
    s = block_sum(s, red);
    const float ls = logf(s);
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
        const float o = (src[v] - m) - ls;
        dst[v] = o;
        if (v == blank && blank_lp != nullptr) blank_lp[row] = o;
    }
}

// Initial state / extend_state: thread per hypothesis, sequential fp32 running sum over t.
__global__ void k_initial_state(const float *__restrict__ blank_lp, int B, int T, int W, int t_begin, float *r0) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    const int BW = B * W;
    if (h >= BW) return;
    const int b = h / W;
    float acc = t_begin > 0 ? r0[((size_t)(t_begin - 1) * 2 + 1) * BW + h] : 0.f;
    for (int t = t_begin; t < T; ++t) {
        const float xb = blank_lp[(size_t)b * T + t];
        acc = (t == 0) ? xb : acc + xb;
        r0[((size_t)t * 2 + 0) * BW + h] = LZ;
        r0[((size_t)t * 2 + 1) * BW + h] = acc;
    }
}

// ------------------------------------------------------------------------------------------
// K-b prep: one warp per (padded) hypothesis builds the stream the recursion consumes.
//   aux[((b*G+g)*Tpad + t)*(HW+1) + hh] = {r_sum[t-1], r_prev[t-1,1], exp(r_sum[t-1]-Gm), exp(r_prev[t-1,1]-Gm)}
//   aux[...                      + HW] = {blank_lp[b,t], 0, 0, 0}
//   Gm[h] = max over the frames log_psi sums over, t-1 in [start-1, end-2]   (end = T unless an attention window is given)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_prep(const float *__restrict__ r_prev, const float *__restrict__ blank_lp, int B,
                                              int W, int T, int HW, int G, int start, int end, int Tpad,
                                              float4 *__restrict__ aux, float *__restrict__ Gmax) {
    const int lane = threadIdx.x & 31;
    const int hp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (hp >= B * G * HW) return;
    const int BW = B * W;
    const int b = hp / (G * HW);
    const int rem = hp - b * (G * HW);
    const int g = rem / HW, hh = rem - g * HW;
    const int w = g * HW + hh;
    const bool valid = w < W;
    const int h = b * W + w;
    float gm = -INFINITY;
    if (valid) {
        for (int t = lane; t < T; t += 32) {
            if (t >= start - 1 && t <= end - 2) {
                const float a = r_prev[((size_t)t * 2 + 0) * BW + h], c = r_prev[((size_t)t * 2 + 1) * BW + h];
                gm = fmaxf(gm, lse2_precise(a, c));
            }
        }
    }
    gm = warp_max(gm);
    if (!(gm > -INFINITY)) gm = 0.f;
    if (valid && lane == 0) Gmax[h] = gm;
    float4 *base = aux + ((size_t)(b * G + g) * Tpad) * (HW + 1);
    for (int te = lane; te < Tpad; te += 32) {
        float4 e = make_float4(LZ, LZ, 0.f, 0.f);
        const int f = te - 1;
        if (valid && f >= 0 && f < T) {
            const float a = r_prev[((size_t)f * 2 + 0) * BW + h], c = r_prev[((size_t)f * 2 + 1) * BW + h];
            const float rs = lse2_precise(a, c);
            e.x = rs;
            e.y = c;
            if (f >= start - 1 && f <= end - 2) {
                e.z = expf(rs - gm);
                e.w = expf(c - gm);
            }
        }
        base[(size_t)te * (HW + 1) + hh] = e;
        if (hh == 0) base[(size_t)te * (HW + 1) + HW] = make_float4(te < T ? blank_lp[(size_t)b * T + te] : 0.f, 0.f, 0.f, 0.f);
    }
}

// ------------------------------------------------------------------------------------------
// K-b main: full-vocabulary forward recursion.
// ------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
// Shared epilogue of the two full-vocabulary scoring kernels: linear-domain sums -> log_psi, token scores, joint
// scores (:164-176, :325, :332).  One call handles the 4 tokens x HW hyps of a thread; rows are accessed as float4
// when V % 4 == 0 (always 16-byte aligned then), the attention rows of all hyps are requested before any is used.
// ------------------------------------------------------------------------------------------
struct EpiArgs {
    const float *Gmax;
    const float *s_prev;
    long long s_rs, s_cs;
    float *att;
    float omw, w;
    float *log_psi, *token_scores, *joint;  // token_scores may be null (the processor does not need it)
    int V, blank, ol;
};

__device__ __forceinline__ void epi_lane(const EpiArgs &e, int h, int v, float S, float gm, float x0, float sp_row, float av,
                                         float &lp_out, float &ts_out, float &jt_out, float &av_out) {
    float lp = gm + logf(S);
    if (!(lp > LZ)) lp = LZ;
    if (e.ol == 0) lp = lse2_precise(lp, x0);
    if (v == e.blank) lp = LZ;
    const float sp = e.s_prev == nullptr ? 0.f : (e.s_cs == 0 ? sp_row : e.s_prev[(long long)h * e.s_rs + (long long)v * e.s_cs]);
    float ts = lp - sp;
    if (ts == 0.f) ts = LZ;
    if (v == e.blank) av = LZ;
    lp_out = lp;
    ts_out = ts;
    av_out = av;
    jt_out = __fadd_rn(__fmul_rn(e.omw, av), __fmul_rn(e.w, ts));
}

// The loop over the hypotheses is deliberately NOT unrolled: with HW = 10 the unrolled body (40 lanes of logf / expf /
// branches) was 13 k SASS instructions = 200 KB, far beyond the instruction caches, and every warp streamed it from L2
// once per tile -- ncu: 35 % of the stall samples of the lazy kernel at C1 were `no_instruction`, a fixed ~25 us per
// launch.  The sums live in registers, which cannot be indexed by a loop variable, so each iteration works on row 0 and
// then shifts the rows down by one (4*(HW-1) register moves).
template <int HW>
__device__ __forceinline__ void epilogue_tile(const EpiArgs &e, const float (&S_in)[HW][4], const float (&x0)[4], int h0, int nhyp,
                                              int v0) {
    const int V = e.V;
    if (v0 >= V) return;
    const bool vec = ((V & 3) == 0);  // then v0 + 3 < V and every row start is 16-byte aligned
    const bool has_att = e.att != nullptr;
    float S[HW][4];
#pragma unroll
    for (int hh = 0; hh < HW; ++hh)
#pragma unroll
        for (int j = 0; j < 4; ++j) S[hh][j] = S_in[hh][j];
    // attention rows of the next EPI_AHEAD hypotheses in flight (ncu at C1, a lone warp per scheduler: with one row
    // ahead an iteration of ~600 cycles waited ~1400 for its row -- 39 % of the kernel's samples were in this loop)
    constexpr int EPI_AHEAD = HW < 4 ? HW : 4;
    float4 attv[EPI_AHEAD];
#pragma unroll
    for (int i = 0; i < EPI_AHEAD; ++i) {
        attv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (vec && has_att && i < nhyp) attv[i] = *reinterpret_cast<const float4 *>(e.att + (size_t)(h0 + i) * V + v0);
    }
    // per-hypothesis scalars of the whole tile requested up front (one load latency instead of one per iteration: the
    // stores of an iteration may alias them as far as the compiler knows); they ride the same shift as the sums
    float gmv[HW], spv[HW];
    const bool row_sp = e.s_prev != nullptr && e.s_cs == 0;
#pragma unroll
    for (int hh = 0; hh < HW; ++hh) {
        const int h = h0 + (hh < nhyp ? hh : 0);
        gmv[hh] = e.Gmax[h];
        spv[hh] = row_sp ? e.s_prev[(long long)h * e.s_rs] : 0.f;
    }
#pragma unroll 1
    for (int hh = 0; hh < nhyp; ++hh) {
        const int h = h0 + hh;
        const float gm = gmv[0];
        const float sp_row = spv[0];
        const size_t o = (size_t)h * V + v0;
        if (vec) {
            const float4 cur = attv[0];
#pragma unroll
            for (int i = 0; i + 1 < EPI_AHEAD; ++i) attv[i] = attv[i + 1];
            if (has_att && hh + EPI_AHEAD < nhyp)  // the row EPI_AHEAD hypotheses ahead
                attv[EPI_AHEAD - 1] = *reinterpret_cast<const float4 *>(e.att + o + (size_t)EPI_AHEAD * V);
            const float av_in[4] = {cur.x, cur.y, cur.z, cur.w};
            float lp[4], ts[4], jt[4], av[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) epi_lane(e, h, v0 + j, S[0][j], gm, x0[j], sp_row, av_in[j], lp[j], ts[j], jt[j], av[j]);
            *reinterpret_cast<float4 *>(e.log_psi + o) = make_float4(lp[0], lp[1], lp[2], lp[3]);
            if (e.token_scores != nullptr) *reinterpret_cast<float4 *>(e.token_scores + o) = make_float4(ts[0], ts[1], ts[2], ts[3]);
            if (has_att) {
                *reinterpret_cast<float4 *>(e.joint + o) = make_float4(jt[0], jt[1], jt[2], jt[3]);
                if (e.blank >= v0 && e.blank < v0 + 4) e.att[(size_t)h * V + e.blank] = LZ;  // scores[:, pad] = logzero in place
            }
        } else {
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                const int v = v0 + j;
                if (v >= V) break;
                float lp, ts, jt, av;
                const float sj = j == 0 ? S[0][0] : (j == 1 ? S[0][1] : (j == 2 ? S[0][2] : S[0][3]));
                const float xj = j == 0 ? x0[0] : (j == 1 ? x0[1] : (j == 2 ? x0[2] : x0[3]));
                epi_lane(e, h, v, sj, gm, xj, sp_row, has_att ? e.att[o + j] : 0.f, lp, ts, jt, av);
                e.log_psi[o + j] = lp;
                if (e.token_scores != nullptr) e.token_scores[o + j] = ts;
                if (has_att) {
                    e.joint[o + j] = jt;
                    if (v == e.blank) e.att[o + j] = LZ;
                }
            }
        }
#pragma unroll
        for (int i = 0; i + 1 < HW; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) S[i][j] = S[i + 1][j];
            gmv[i] = gmv[i + 1], spv[i] = spv[i + 1];
        }
    }
}

struct ScoreArgs {
    const float4 *aux;
    const float *Gmax;
    const float *s_prev;
    long long s_rs, s_cs;
    const int64_t *last_ids;
    float *att;
    float omw, w;
    float *r;
    int ldr;
    float *log_psi, *token_scores, *joint;
    int B, W, T, V, blank, ol, G, Tpad, nvt;
    int start, end;  // frames the recursion runs over (:127-136): max(ol, 1) .. T unless an attention window is given
};

template <int HW, int NT>
struct ScoreSmem {
    static constexpr int VTILE = NT * 4;
    static constexpr int NBOX = VTILE / BOXC;
    alignas(128) float xs[NS][NBOX][TT][BOXC];
    alignas(16) float4 auxs[NS][TT][HW + 1];
    alignas(8) uint64_t full[NS];
    alignas(8) uint64_t empty[NS];
};

// One frame of the recursion for the 4 tokens of a thread and one hypothesis.
__device__ __forceinline__ void recur_hyp(float (&rn)[4], float (&rb)[4], float (&psi)[4], const float4 ph, const float (&xv)[4],
                                          const float (&p)[4], float xb, int cj) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const bool last = (cj == j);
        const float phi = last ? ph.y : ph.x;
        const float lin = last ? ph.w : ph.z;
        const float nn = lse2_fast(rn[j], phi) + xv[j];   // :150-151, non-blank row
        const float nb = lse2_fast(rn[j], rb[j]) + xb;    // :150-151, blank row
        rn[j] = nn;
        rb[j] = nb;
        psi[j] = fmaf(lin, p[j], psi[j]);                 // :154,164-167 in the linear domain
    }
}

// Persistent kernel (grid = #SMs x resident CTAs; CTA i walks tiles i, i+grid, ...; tile = (utterance, 512-token tile,
// hyp group)).  Thread 0 keeps the flat chunk sequence of its tiles NS stages ahead with TMA; stages come back through
// `empty` mbarriers (one arrival per warp), so the warps of a CTA never meet at a block-wide barrier inside the stream.
template <int HW, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_score_full(const __grid_constant__ CUtensorMap tmx, const ScoreArgs a) {
    using Smem = ScoreSmem<HW, NT>;
    constexpr int VTILE = Smem::VTILE;
    constexpr int NBOX = Smem::NBOX;
    constexpr int NWARP = NT / 32;
    static_assert(VTILE % BOXC == 0, "tile must be whole TMA boxes");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);

    const int tid = threadIdx.x;
    const int T = a.T, V = a.V, W = a.W, BW = a.B * a.W;
    const int start = a.start, end = a.end;
    const int c0 = (a.ol == 0 ? 0 : start) / TT;
    const int cN = (end - 1) / TT;
    const int nchunk = cN >= c0 ? cN - c0 + 1 : 0;  // an empty window (start == end) streams nothing
    const int ntiles = a.B * a.nvt * a.G;
    const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int nitems = my_tiles * nchunk;
    constexpr uint32_t STAGE_BYTES = TT * VTILE * 4 + TT * (HW + 1) * 16;
    const size_t plane = (size_t)BW * a.ldr;   // floats between the non-blank and blank planes
    const size_t frame = 2 * plane;            // floats per frame of r

    auto decode_tile = [&](int tile, int &b, int &vt, int &g) {
        g = tile % a.G;
        tile /= a.G;
        vt = tile % a.nvt;
        b = tile / a.nvt;
    };
    auto issue = [&](int k) {  // item k = (k / nchunk)-th tile of this CTA, chunk c0 + k % nchunk
        int b, vt, g;
        decode_tile((int)blockIdx.x + (k / nchunk) * (int)gridDim.x, b, vt, g);
        const int c = c0 + k % nchunk;
        const int s = k % NS;
        mbar_expect_tx(&sm.full[s], STAGE_BYTES);
#pragma unroll
        for (int bx = 0; bx < NBOX; ++bx) tma_load_2d(&sm.xs[s][bx][0][0], &tmx, vt * VTILE + bx * BOXC, b * T + c * TT, &sm.full[s]);
        bulk_load_1d(&sm.auxs[s][0][0], a.aux + ((size_t)(b * a.G + g) * a.Tpad + (size_t)c * TT) * (HW + 1), TT * (HW + 1) * 16,
                     &sm.full[s]);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) mbar_init(&sm.full[s], 1), mbar_init(&sm.empty[s], NWARP);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int k = 0; k < nitems && k < NS; ++k) issue(k);
    }
    __syncthreads();  // barrier init visible before anyone waits

    const int bx = (tid * 4) / BOXC, col = (tid * 4) % BOXC;
    int k = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
        int b, vt, g;
        decode_tile((int)blockIdx.x + ti * (int)gridDim.x, b, vt, g);
        const int v0 = vt * VTILE + tid * 4;
        const int h0 = b * W + g * HW;
        const int nhyp = min(HW, W - g * HW);  // valid hypotheses of this group
        const bool lane_ok = v0 < a.ldr;

        // which of my 4 tokens (if any) is the last label of hypothesis hh            (:122-124)
        int cj[HW];
#pragma unroll
        for (int hh = 0; hh < HW; ++hh) cj[hh] = hh < nhyp ? (int)(a.last_ids[h0 + hh] - (long long)v0) : -1;
        float rn[HW][4], rb[HW][4], psi[HW][4], x0[4];
#pragma unroll
        for (int hh = 0; hh < HW; ++hh)
#pragma unroll
            for (int j = 0; j < 4; ++j) rn[hh][j] = LZ, rb[hh][j] = LZ, psi[hh][j] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) x0[j] = LZ;
        float *rbase = a.r + (size_t)h0 * a.ldr + v0;

        // frames before `start` stay logzero (r = full(logzero), :106-111); frame 0 of the first step is set below
        if (lane_ok) {
            const float4 lz4 = make_float4(LZ, LZ, LZ, LZ);
            for (int t = (a.ol == 0 ? 1 : 0); t < start; ++t) {
                float *rp = rbase + (size_t)t * frame;
                for (int hh = 0; hh < nhyp; ++hh) {
                    __stcs(reinterpret_cast<float4 *>(rp + (size_t)hh * a.ldr), lz4);
                    __stcs(reinterpret_cast<float4 *>(rp + (size_t)hh * a.ldr + plane), lz4);
                }
            }
        }

        for (int ci = 0; ci < nchunk; ++ci, ++k) {
            const int c = c0 + ci;
            const int s = k % NS;
            mbar_wait(&sm.full[s], (uint32_t)((k / NS) & 1));
            const int tmax = min(TT, end - c * TT);
            for (int tt = 0; tt < tmax; ++tt) {
                const int t = c * TT + tt;
                const float4 xv4 = *reinterpret_cast<const float4 *>(&sm.xs[s][bx][tt][col]);
                const float xv[4] = {xv4.x, xv4.y, xv4.z, xv4.w};
                float *rp = rbase + (size_t)t * frame;
                if (t < start) {
                    if (t == 0 && a.ol == 0) {  // r[0,0] = x_[0,0]                      (:112-113)
#pragma unroll
                        for (int j = 0; j < 4; ++j) x0[j] = xv[j];
#pragma unroll
                        for (int hh = 0; hh < HW; ++hh)
#pragma unroll
                            for (int j = 0; j < 4; ++j) rn[hh][j] = xv[j];
                        if (lane_ok)
                            for (int hh = 0; hh < nhyp; ++hh) {
                                __stcs(reinterpret_cast<float4 *>(rp + (size_t)hh * a.ldr), xv4);
                                __stcs(reinterpret_cast<float4 *>(rp + (size_t)hh * a.ldr + plane), make_float4(LZ, LZ, LZ, LZ));
                            }
                    }
                    continue;
                }
                const float xb = sm.auxs[s][tt][HW].x;
                float p[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) p[j] = ex2_approx(xv[j] * LOG2E);
#pragma unroll
                for (int hh = 0; hh < HW; ++hh) {
                    const float4 ph = sm.auxs[s][tt][hh];
                    recur_hyp(rn[hh], rb[hh], psi[hh], ph, xv, p, xb, cj[hh]);
                    if (lane_ok && hh < nhyp) {
                        __stcs(reinterpret_cast<float4 *>(rp + (size_t)hh * a.ldr), make_float4(rn[hh][0], rn[hh][1], rn[hh][2], rn[hh][3]));
                        __stcs(reinterpret_cast<float4 *>(rp + (size_t)hh * a.ldr + plane),
                               make_float4(rb[hh][0], rb[hh][1], rb[hh][2], rb[hh][3]));
                    }
                }
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&sm.empty[s]);
            if (tid == 0 && k + NS < nitems) {  // refill the stage just released, once every warp is done with it
                mbar_wait(&sm.empty[s], (uint32_t)((k / NS) & 1));
                issue(k + NS);
            }
        }

        // frames past an attention window stay logzero as well (r = full(logzero), :106-111)
        if (lane_ok && end < T) {
            const float4 lz4 = make_float4(LZ, LZ, LZ, LZ);
            for (int t = end; t < T; ++t) {
                float *rp = rbase + (size_t)t * frame;
                for (int hh = 0; hh < nhyp; ++hh) {
                    __stcs(reinterpret_cast<float4 *>(rp + (size_t)hh * a.ldr), lz4);
                    __stcs(reinterpret_cast<float4 *>(rp + (size_t)hh * a.ldr + plane), lz4);
                }
            }
        }

        // epilogue: log_psi, token scores, joint scores                                 (:164-176, :325, :332)
        EpiArgs e;
        e.Gmax = a.Gmax, e.s_prev = a.s_prev, e.s_rs = a.s_rs, e.s_cs = a.s_cs, e.att = a.att, e.omw = a.omw, e.w = a.w;
        e.log_psi = a.log_psi, e.token_scores = a.token_scores, e.joint = a.joint, e.V = V, e.blank = a.blank, e.ol = a.ol;
        epilogue_tile<HW>(e, psi, x0, h0, nhyp, v0);
    }
}

#include "ctcps_psi.cuh"

// Lazy index_select_state: re-run the forward recursion of the PREVIOUS step for the surviving (hyp, token)
// column of every output hypothesis j -- exactly the lane k_score_full would have written to r[:, :, hyp, tok].
// Two launches: (a) fully parallel over (t, j): stage phi[t-1] and x[t, tok] into r_new[t, 0:2, j] (gathers,
// libm-grade logsumexp off the dependent chain); (b) one thread per j walks T in place with coalesced,
// prefetched reads -- only the two MUFU logaddexp remain on the serial chain.
struct LazySel {
    long long hs;   // source hypothesis (global row) whose column of r is taken
    long long tok;  // token of that column
    long long last; // the token the output hypothesis ends with (= tok unless it was not scored and lane 0 stands in)
    long long src;  // where s_new comes from: index into log_psi (BW,V) or, with candidates, cand_log_psi (BW,S); -1 = logzero
};
// cand_ids (BW,S) non-null = the step was scored on candidates only (ids unique per hypothesis): the lane is
// scoring_idmap[hyp, tok], a token that was not scored selects lane 0 (:196-202) and its prefix score is logzero (:156).
__device__ __forceinline__ LazySel lazy_source(const int64_t *__restrict__ best_ids, const int64_t *__restrict__ cand_ids, int S,
                                               int j, int W, int V) {
    const int b = j / W;
    const long long flat = best_ids[j] + (long long)b * W * V;  // :191
    LazySel q;
    q.hs = flat / V;
    q.tok = flat - q.hs * V;
    q.last = q.tok;
    q.src = flat;
    if (cand_ids != nullptr) {
        const int64_t *c = cand_ids + q.hs * S;
        int pos = -1;
        for (int s = 0; s < S; ++s)
            if (c[s] == q.tok) pos = s;
        q.src = pos < 0 ? -1 : q.hs * S + pos;
        if (pos < 0) q.tok = c[0];
    }
    return q;
}

// (a) thread = (output hypothesis j, chunk of LAZY_TC frames): resolves its source column once, then stages
// phi[t-1] (non-blank plane of r_new) and x[t, tok] (blank plane) for its frames.
constexpr int LAZY_TC = 8;
__global__ void __launch_bounds__(128) k_select_lazy_stage(const XView x, const float *__restrict__ r_prev,
                                                           const int64_t *__restrict__ last_ids, int ol,
                                                           const int64_t *__restrict__ best_ids,
                                                           const int64_t *__restrict__ cand_ids, int S, int B, int W, int T, int V,
                                                           float *__restrict__ r_new) {
    const int BW = B * W;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= BW) return;
    const int start = ol > 1 ? ol : 1;
    const LazySel q = lazy_source(best_ids, cand_ids, S, j, W, V);
    const bool last = last_ids[q.hs] == q.tok;
    const int b = j / W;
    const int t0 = blockIdx.y * LAZY_TC, t1 = min(T, t0 + LAZY_TC);
    for (int t = t0; t < t1; ++t) {
        float a = LZ, c = LZ;
        if (t >= start) {
            const float p0 = r_prev[((size_t)(t - 1) * 2 + 0) * BW + q.hs], p1 = r_prev[((size_t)(t - 1) * 2 + 1) * BW + q.hs];
            a = last ? p1 : lse2_precise(p0, p1);
            c = x.at(b, t, q.tok);
        } else if (t == 0 && ol == 0) {
            a = x.at(b, 0, q.tok);  // r[0,0] = x_[0,0] (:112-113); the blank plane stays logzero
        }
        r_new[((size_t)t * 2 + 0) * BW + j] = a;
        r_new[((size_t)t * 2 + 1) * BW + j] = c;
    }
}

// NEXT: with `lin` given, the thread also produces what k_prep_psi would compute from r_new for the NEXT score call
// (its hypothesis j, last label = the selected token, prefix length ol + 1): the lin stream, the offset and the
// last-label column sum.  The offset is s_new[j] = log psi(prefix j) instead of max_t r_sum[t]: psi(h) >= gamma_t(h)
// for every t (a path whose first t frames collapse to h has h as a prefix) and psi(h) <= T * max_t gamma_t(h), so
// exp(r_sum - s_new) lies in [~1/T, 1] at its largest -- as good an offset, and known before the scan starts.
template <bool NEXT>
__global__ void __launch_bounds__(64) k_select_lazy_scan(const XView x, const float *__restrict__ blank_lp, int ol,
                                                         const float *__restrict__ log_psi,  // (BW,V), or (BW,S) with cand_ids
                                                         const int64_t *__restrict__ best_ids,
                                                         const int64_t *__restrict__ cand_ids, int S, int B, int W, int T, int V,
                                                         float *r_new, float *s_new, float *__restrict__ lin,
                                                         float *__restrict__ Gmax, float *__restrict__ psic, int HW, int HWP, int G,
                                                         int Tpad) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int BW = B * W;
    if (j >= BW) return;
    const LazySel q = lazy_source(best_ids, cand_ids, S, j, W, V);
    const float sj = q.src < 0 ? LZ : log_psi[q.src];
    s_new[j] = sj;  // :193
    const int start = ol > 1 ? ol : 1;
    const float *xb = blank_lp + (size_t)(j / W) * T;
    float rn = (ol == 0) ? r_new[j] : LZ, rb = LZ;

    // NEXT: frames f of r_new that the next log_psi sums over are [ol, T-2]; lin entry te = f + 1
    float *lbase = nullptr;
    float pc = 0.f, mx = -INFINITY;  // mx: max of r_sum over the summed frames
    if (NEXT) {
        const int w = j % W;
        lbase = lin + ((size_t)((j / W) * G + w / HW) * Tpad) * HWP + (w % HW);
        for (int te = 0; te <= ol && te < Tpad; ++te) lbase[(size_t)te * HWP] = 0.f;
        for (int te = T; te < Tpad; ++te) lbase[(size_t)te * HWP] = 0.f;
    }
    // One warp per scheduler at best (BW threads in all), so the time of this kernel is (instructions per frame) x (issue
    // interval of a lone warp): running pointers instead of index arithmetic (the first version spent 3 of 4 instructions
    // on 64-bit address math: 383 cycles per frame), no per-frame predicates (first frame peeled, full batches, scalar
    // tail), and a software pipeline: the loads of batch k+1 are in flight while the serial chain of batch k runs.
    const size_t st2 = 2 * (size_t)BW;  // floats between two frames of r_new
    int t = start;
    float *pt = r_new + (size_t)t * st2 + j;     // r_new[t, 0, j]; the blank plane is BW floats further
    const float *pb = xb + t;
    float *pl = NEXT ? lbase + (size_t)t * HWP : nullptr;
    auto frame = [&](float ph_t, float xv_t, float bl_t, bool next) {
        const float ls = lse2_fast(rn, rb);  // = r_sum[t-1], a by-product of the blank row
        if (NEXT && next) {                  // frame f = t-1 of the next step's sums (f in [ol, T-2]), entry te = t
            mx = fmaxf(mx, ls);
            *pl = ex2_approx(fminf(ls - sj, 0.f) * LOG2E);
            // last-label column: r_prev_blank[f] * p[f+1, tok]
            pc = fmaf(ex2_approx(fminf(rb - sj, 0.f) * LOG2E), ex2_approx(xv_t * LOG2E), pc);
        }
        const float nn = lse2_fast(rn, ph_t) + xv_t;
        rn = nn;
        rb = ls + bl_t;
        pt[0] = rn;
        pt[BW] = rb;
        pt += st2;
        if (NEXT) pl += HWP;
    };
    if (ol >= 1 && t < T) {  // t - 1 = ol - 1 is not among the frames the next step sums over
        frame(pt[0], pt[BW], *pb, false);
        ++pb, ++t;
    }
    constexpr int U = 8;
    if (t + U <= T) {
        float ph[U], xv[U], bl[U];
        auto load = [&](const float *p, const float *xbp, float (&a)[U], float (&c)[U], float (&d)[U]) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                a[u] = p[0];
                c[u] = p[BW];
                d[u] = xbp[u];
                p += st2;
            }
        };
        load(pt, pb, ph, xv, bl);
        for (; t + U <= T; t += U, pb += U) {
            float ph2[U], xv2[U], bl2[U];
            const bool more = t + 2 * U <= T;
            if (more) load(pt + U * st2, pb + U, ph2, xv2, bl2);
#pragma unroll
            for (int u = 0; u < U; ++u) frame(ph[u], xv[u], bl[u], true);
            if (more) {
#pragma unroll
                for (int u = 0; u < U; ++u) ph[u] = ph2[u], xv[u] = xv2[u], bl[u] = bl2[u];
            }
        }
    }
    for (; t < T; ++t, ++pb) frame(pt[0], pt[BW], *pb, true);
    if (NEXT) {
        float gm = sj;
        // s_new is not a usable offset when it is not a prefix probability (finished beam: last label = pad, s_new = logzero)
        // or when every summed frame lies far below it (label aligned to the very last frames): redo this hypothesis'
        // stream against max_t r_sum, exactly like k_prep_psi.  Rare, and off the serial chain (plain loads).
        if (!(sj > -1e9f) || mx < sj - 60.f) {
            gm = mx > -INFINITY ? mx : 0.f;
            pc = 0.f;
            for (int f = ol; f <= T - 2; ++f) {
                const float a = r_new[((size_t)f * 2 + 0) * BW + j], c = r_new[((size_t)f * 2 + 1) * BW + j];
                lbase[(size_t)(f + 1) * HWP] = expf(lse2_precise(a, c) - gm);
                pc = fmaf(expf(c - gm), expf(x.at(j / W, f + 1, q.last)), pc);  // the NEXT step's last label is the real token
            }
        }
        Gmax[j] = gm;
        psic[j] = pc;
    }
}

// ------------------------------------------------------------------------------------------
// Time-parallel variant of k_select_lazy_scan (the default; CTCPS_SELECT_PSCAN=0 / ctcps_set_select_pscan(0) selects the sequential one).
// The recursion is an affine map in the (logsumexp, +) semiring,
//     rn' = lse(rn + xv, ph + xv)        rb' = lse(rn + bl, rb + bl)
// with five live coefficients (A_nn, A_bn, A_bb, c_n, c_b: rn never depends on rb), and affine maps compose
// associatively.  A WARP owns one output hypothesis: lane l composes the maps of its block of F = ceil((T-start)/32)
// frames (depth F), a 5-level warp scan of the 32 block maps gives every lane its incoming state, and the lane replays
// its block with the same `frame` arithmetic as the sequential kernel (depth F): 2F + 6 dependent logsumexp levels
// instead of T.  Everything stays in the log domain with the finite logzero, so there is no dynamic-range problem.
// Results differ from the sequential kernel only by the rounding of the block-entry states (numpy emulation against the
// reference's fp64 run: tools/pscan_prototype.py).  A CTA holds PS_H hypotheses; their staged columns of r_new (T,2,BW)
// go through shared memory so that global accesses stay coalesced across hypotheses; in shared memory a column is
// hypothesis-major with an odd hypothesis stride and an odd per-lane block stride (conflict-free on both sides).
// ------------------------------------------------------------------------------------------
constexpr int PS_H = 16;  // hypotheses (warps) per CTA

struct AffMap {
    float ann, abn, abb, cn, cb;
};
__device__ __forceinline__ AffMap aff_compose(const AffMap &m2, const AffMap &m1) {  // m2 after m1
    AffMap r;
    r.ann = m2.ann + m1.ann;
    r.abn = lse2_fast(m2.abn + m1.ann, m2.abb + m1.abn);
    r.abb = m2.abb + m1.abb;
    r.cn = lse2_fast(m2.ann + m1.cn, m2.cn);
    r.cb = lse2_fast(lse2_fast(m2.abn + m1.cn, m2.abb + m1.cb), m2.cb);
    return r;
}
__device__ __forceinline__ AffMap aff_shfl_up(const AffMap &m, int d) {
    AffMap r;
    r.ann = __shfl_up_sync(0xffffffffu, m.ann, d);
    r.abn = __shfl_up_sync(0xffffffffu, m.abn, d);
    r.abb = __shfl_up_sync(0xffffffffu, m.abb, d);
    r.cn = __shfl_up_sync(0xffffffffu, m.cn, d);
    r.cb = __shfl_up_sync(0xffffffffu, m.cb, d);
    return r;
}

template <bool NEXT>
__global__ void __launch_bounds__(PS_H * 32) k_select_lazy_pscan(const XView x, const float *__restrict__ blank_lp, int ol,
                                                                 const float *__restrict__ log_psi,
                                                                 const int64_t *__restrict__ best_ids,
                                                                 const int64_t *__restrict__ cand_ids, int S, int B, int W, int T,
                                                                 int V, float *r_new, float *s_new,
                                                                 float *__restrict__ lin, float *__restrict__ Gmax,
                                                                 float *__restrict__ psic, int HW, int HWP, int G, int Tpad, int F,
                                                                 int Fo, int HS) {
    extern __shared__ float ps_sm[];  // [PS_H][HS]: hypothesis jj, plane k, lane block l, offset i at jj*HS + k*32*Fo + l*Fo + i
    const int BW = B * W;
    const int j0 = blockIdx.x * PS_H;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int start = ol > 1 ? ol : 1;
    const int n = T - start;  // frames the recursion walks: t = start .. T-1 (n >= 1, checked by the host)
    const size_t st2 = 2 * (size_t)BW;

    // P0: staged columns (phi[t-1] in plane 0, x[t, tok] in plane 1) of this CTA's hypotheses, coalesced across hypotheses
    for (int idx = tid; idx < n * 2 * PS_H; idx += PS_H * 32) {
        const int jj = idx % PS_H, row = idx / PS_H;  // row = (t - start) * 2 + k
        const int k = row & 1, tr = row >> 1;
        const int l = tr / F, i = tr - l * F;
        float v = LZ;
        if (j0 + jj < BW) v = r_new[(size_t)(start * 2 + row) * BW + j0 + jj];
        ps_sm[jj * HS + k * 32 * Fo + l * Fo + i] = v;
    }
    __syncthreads();

    const int j = j0 + w;
    const bool active = j < BW;  // warp-uniform
    float sj = LZ, gm_mx = -INFINITY, pc = 0.f;
    float *lbase = nullptr;
    LazySel q;
    q.hs = 0, q.tok = 0, q.last = 0, q.src = -1;
    if (active) {
        q = lazy_source(best_ids, cand_ids, S, j, W, V);
        sj = q.src < 0 ? LZ : log_psi[q.src];
        if (lane == 0) s_new[j] = sj;  // :193
        const float *xb = blank_lp + (size_t)(j / W) * T;
        float *col = ps_sm + w * HS;
        const int lo = lane * F, hi = min(n, lo + F);  // this lane's frames, relative to `start`
        // P1: the map of this lane's block
        AffMap m;
        m.ann = 0.f, m.abn = LZ, m.abb = 0.f, m.cn = LZ, m.cb = LZ;
        for (int i = 0; lo + i < hi; ++i) {
            const float ph = col[lane * Fo + i], xv = col[32 * Fo + lane * Fo + i], bl = xb[start + lo + i];
            const float abn = bl + lse2_fast(m.ann, m.abn);
            const float cb = bl + lse2_fast(m.cn, m.cb);
            m.cn = xv + lse2_fast(m.cn, ph);
            m.ann = xv + m.ann;
            m.abb = bl + m.abb;
            m.abn = abn;
            m.cb = cb;
        }
        // P2: inclusive scan over the lanes, then one step up = the map of everything before this lane's block
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const AffMap up = aff_shfl_up(m, d);
            if (lane >= d) m = aff_compose(m, up);
        }
        AffMap e = aff_shfl_up(m, 1);
        if (lane == 0) e.ann = 0.f, e.abn = LZ, e.abb = 0.f, e.cn = LZ, e.cb = LZ;
        const float rn0 = (ol == 0) ? r_new[j] : LZ, rb0 = LZ;
        float rn = lse2_fast(e.ann + rn0, e.cn);
        float rb = lse2_fast(lse2_fast(e.abn + rn0, e.abb + rb0), e.cb);
        // P3: replay the block (same arithmetic per frame as k_select_lazy_scan) and the by-products of the NEXT scoring call
        if (NEXT) {
            const int wv = j % W;
            lbase = lin + ((size_t)((j / W) * G + wv / HW) * Tpad) * HWP + (wv % HW);
            for (int te = lane; te <= ol && te < Tpad; te += 32) lbase[(size_t)te * HWP] = 0.f;
            for (int te = T + lane; te < Tpad; te += 32) lbase[(size_t)te * HWP] = 0.f;
        }
        for (int i = 0; lo + i < hi; ++i) {
            const int t = start + lo + i;
            const float ph = col[lane * Fo + i], xv = col[32 * Fo + lane * Fo + i], bl = xb[t];
            const float ls = lse2_fast(rn, rb);  // = r_sum[t-1]
            if (NEXT && !(ol >= 1 && t == start)) {  // frame f = t-1 of the next step's sums (f in [ol, T-2]), entry te = t
                gm_mx = fmaxf(gm_mx, ls);
                lbase[(size_t)t * HWP] = ex2_approx(fminf(ls - sj, 0.f) * LOG2E);
                pc = fmaf(ex2_approx(fminf(rb - sj, 0.f) * LOG2E), ex2_approx(xv * LOG2E), pc);
            }
            rn = lse2_fast(rn, ph) + xv;
            rb = ls + bl;
            col[lane * Fo + i] = rn;
            col[32 * Fo + lane * Fo + i] = rb;
        }
    }
    __syncthreads();
    // P4: the new forward variables back to r_new, coalesced across hypotheses
    for (int idx = tid; idx < n * 2 * PS_H; idx += PS_H * 32) {
        const int jj = idx % PS_H, row = idx / PS_H;
        const int k = row & 1, tr = row >> 1;
        const int l = tr / F, i = tr - l * F;
        if (j0 + jj < BW) r_new[(size_t)(start * 2 + row) * BW + j0 + jj] = ps_sm[jj * HS + k * 32 * Fo + l * Fo + i];
    }
    if (NEXT && active) {
        const float mx = warp_max(gm_mx);
        pc = warp_sum(pc);
        float gm = sj;
        // same fallback as the sequential kernel: s_new is not a usable offset for a finished beam or when every summed frame
        // lies far below it -- redo the stream against max_t r_sum (rare; the lanes share the frames)
        if (!(sj > -1e9f) || mx < sj - 60.f) {
            gm = mx > -INFINITY ? mx : 0.f;
            pc = 0.f;
            const float *col = ps_sm + w * HS;
            for (int f = ol + lane; f <= T - 2; f += 32) {
                float a, c;
                if (f >= start) {
                    const int tr = f - start, l = tr / F, i = tr - l * F;
                    a = col[l * Fo + i], c = col[32 * Fo + l * Fo + i];
                } else {  // f = 0 with ol = 0: the staged frame 0 was not touched by the scan
                    a = r_new[(size_t)f * st2 + j], c = r_new[(size_t)f * st2 + BW + j];
                }
                lbase[(size_t)(f + 1) * HWP] = expf(lse2_precise(a, c) - gm);
                pc = fmaf(expf(c - gm), expf(x.at(j / W, f + 1, q.last)), pc);
            }
            pc = warp_sum(pc);
        }
        if (lane == 0) {
            Gmax[j] = gm;
            psic[j] = pc;
        }
    }
}

// ------------------------------------------------------------------------------------------
// K-b partial scoring: one thread per (hyp, candidate) lane.  ~V/S times less work than the full
// path; written for fidelity (libm-grade exp/log, running-max logsumexp), not for the roofline.
// ------------------------------------------------------------------------------------------
__global__ void k_fill_i64(int64_t *p, size_t n, int64_t v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_fill_f32(float *p, size_t n, float v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
// scoring_idmap[h, scoring_ids[h,s]] = s, later s wins                              (:91-95)
__global__ void k_build_idmap(const int64_t *__restrict__ ids, int BW, int S, int V, int64_t *idmap) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= BW) return;
    for (int s = 0; s < S; ++s) {
        const int64_t v = ids[(size_t)h * S + s];
        if (v >= 0 && v < V) idmap[(size_t)h * V + v] = s;
    }
}

__global__ void __launch_bounds__(128) k_score_partial(const float *__restrict__ x, int ldx, const float *__restrict__ blank_lp,
                                                       const float *__restrict__ r_prev, const int64_t *__restrict__ last_ids,
                                                       const int64_t *__restrict__ ids, const int64_t *__restrict__ idmap, int ol,
                                                       int start, int end, int B, int W, int T, int V, int S, float *r, int ldr,
                                                       float *log_psi) {
    const int lane = blockIdx.x * blockDim.x + threadIdx.x;
    const int BW = B * W;
    if (lane >= BW * S) return;
    const int h = lane / S, s = lane - h * S;
    const int b = h / W;
    const int64_t v = ids[lane];
    const int64_t c = last_ids[h];
    const bool last = (c >= 0 && c < V) ? (idmap[(size_t)h * V + c] == s) : false;  // :117-121
    const size_t plane = (size_t)BW * ldr, frame = 2 * plane;
    float *rp = r + (size_t)h * ldr + s;
    if (v < 0 || v >= V) {  // the reference raises an IndexError here; never read outside the posteriors: the lane is logzero
        for (int t = 0; t < T; ++t) {
            rp[(size_t)t * frame] = LZ;
            rp[(size_t)t * frame + plane] = LZ;
        }
        return;
    }
    const float *xr = x + (size_t)b * T * ldx + v;
    for (int t = 0; t < start && t < T; ++t) {
        rp[(size_t)t * frame] = LZ;
        rp[(size_t)t * frame + plane] = LZ;
    }
    float rn = LZ, rb = LZ;
    if (ol == 0) {
        rn = xr[0];
        rp[0] = rn;
    }
    for (int t = end; t < T; ++t) {  // frames past an attention window (:127-136)
        rp[(size_t)t * frame] = LZ;
        rp[(size_t)t * frame + plane] = LZ;
    }
    float m = rn, acc = 1.f;  // running-max logsumexp seeded with r[start-1,0]       (:158,165)
    for (int t = start; t < end; ++t) {
        const float p0 = r_prev[((size_t)(t - 1) * 2 + 0) * BW + h], p1 = r_prev[((size_t)(t - 1) * 2 + 1) * BW + h];
        const float phi = last ? p1 : lse2_precise(p0, p1);
        const float xv = xr[(size_t)t * ldx];
        const float nn = lse2_precise(rn, phi) + xv;
        const float nb = lse2_precise(rn, rb) + blank_lp[(size_t)b * T + t];
        rn = nn;
        rb = nb;
        rp[(size_t)t * frame] = rn;
        rp[(size_t)t * frame + plane] = rb;
        const float term = phi + xv;
        if (term > m) {
            acc = acc * expf(m - term) + 1.f;
            m = term;
        } else {
            acc += expf(term - m);
        }
    }
    if (v >= 0 && v < V && idmap[(size_t)h * V + v] == s) log_psi[(size_t)h * V + v] = logf(acc) + m;  // :161-162, later s wins
}

// blank exclusion, relative score, zero hack, joint combine for the paths that do not fuse it
__global__ void k_finalize(float *log_psi, const float *__restrict__ s_prev, long long s_rs, long long s_cs, float *att,
                           float omw, float w, int BW, int V, int blank, float *token_scores, float *joint, int all_logzero) {
    const size_t n = (size_t)BW * V;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int h = (int)(i / V), v = (int)(i - (size_t)h * V);
        float ts;
        if (all_logzero) {  // start > end early return                               (:138-145)
            log_psi[i] = LZ;
            ts = LZ;
        } else {
            float lp = log_psi[i];
            if (v == blank) lp = LZ, log_psi[i] = LZ;
            const float sp = s_prev != nullptr ? s_prev[(long long)h * s_rs + (long long)v * s_cs] : 0.f;
            ts = lp - sp;
            if (ts == 0.f) ts = LZ;
        }
        if (token_scores != nullptr) token_scores[i] = ts;
        if (att != nullptr) {
            float av = att[i];
            if (v == blank) av = LZ, att[i] = LZ;
            joint[i] = __fadd_rn(__fmul_rn(omw, av), __fmul_rn(w, ts));
        }
    }
}

// ------------------------------------------------------------------------------------------
// K-c: index_select_state gather.
// ------------------------------------------------------------------------------------------
__global__ void k_select(const float *__restrict__ r, int ldr, const float *__restrict__ log_psi,
                         const int64_t *__restrict__ best_ids, const int64_t *__restrict__ idmap, int B, int W, int T, int V,
                         int S, float *__restrict__ r_new, float *__restrict__ s_new) {
    const int BW = B * W;
    const size_t n = (size_t)T * 2 * BW;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(i % BW);
        const size_t tk = i / BW;
        const int b = j / W;
        const long long best = best_ids[j];
        const long long flat = best + (long long)b * W * V;  // :191
        long long hyp, lanei;
        if (idmap != nullptr) {  // :196-202
            hyp = best / V + (long long)b * W;
            long long label = best % V;
            long long si = idmap[hyp * V + label];
            lanei = si < 0 ? 0 : si;
        } else {
            hyp = flat / V;
            lanei = flat - hyp * V;
        }
        r_new[i] = r[(tk * BW + (size_t)hyp) * ldr + (size_t)lanei];  // :206
        if (tk == 0) s_new[j] = log_psi[flat];                            // :193
    }
}

// ------------------------------------------------------------------------------------------
// eos/space trick: one CTA per row, first-max argmax like torch.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_trick(const float *__restrict__ att, const float *__restrict__ ctc, float *next, int V,
                                               int eos, int space, float k) {
    __shared__ float sv[2][8];
    __shared__ int si[2][8];
    const int row = blockIdx.x;
    const float *pa = att + (size_t)row * V, *pc = ctc + (size_t)row * V;
    float ba = -INFINITY, bc = -INFINITY;
    int ia = 0x7fffffff, ic = 0x7fffffff;
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
        const float x = pa[v], y = pc[v];
        if (x > ba) ba = x, ia = v;
        if (y > bc) bc = y, ic = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float oa = __shfl_xor_sync(0xffffffffu, ba, o), oc = __shfl_xor_sync(0xffffffffu, bc, o);
        int ja = __shfl_xor_sync(0xffffffffu, ia, o), jc = __shfl_xor_sync(0xffffffffu, ic, o);
        if (oa > ba || (oa == ba && ja < ia)) ba = oa, ia = ja;
        if (oc > bc || (oc == bc && jc < ic)) bc = oc, ic = jc;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sv[0][wid] = ba, si[0][wid] = ia, sv[1][wid] = bc, si[1][wid] = ic;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < (int)(blockDim.x >> 5); ++q) {
            if (sv[0][q] > ba || (sv[0][q] == ba && si[0][q] < ia)) ba = sv[0][q], ia = si[0][q];
            if (sv[1][q] > bc || (sv[1][q] == bc && si[1][q] < ic)) bc = sv[1][q], ic = si[1][q];
        }
        if (ia == eos && ic == space) {
            float *n = next + (size_t)row * V;
            const float ne = n[eos], nsp = n[space];
            if (ne < nsp && k * ne > nsp) n[eos] = ne * k;
        }
    }
}

// ------------------------------------------------------------------------------------------
// N1: the beam-search step that follows the processor, fused into one launch.
// One CTA per utterance: top-2W of (joint + running beam score) over the W*V candidates, eos finalisation into
// the finished-hypothesis pool, choice of the W continuing beams, done test, and the reordered + extended
// input_ids rows.  Restates huggingface_asr_b200/beam_search.py::joint_beam_search (itself the contract of HF
// 4.39.3 beam_search / BeamSearchScorer.process with the processor) without ~60 small torch launches per step.
// Ordering: higher score first; equal scores: lower hyp*V+tok first.
// ------------------------------------------------------------------------------------------
constexpr int BEAM_NT = 128;
constexpr int BEAM_NW = BEAM_NT / 32;
constexpr int BEAM_MAXK = 64;   // 2W <= 64
constexpr int BEAM_MAXP = 8;    // CTAs per utterance

struct Cand {
    float s;
    int i;
};
__device__ __forceinline__ bool cand_beats(float s1, int i1, float s2, int i2) { return s1 > s2 || (s1 == s2 && i1 < i2); }

// Sorted top-(32*KL) list of a warp, held in registers: entry q of list j lives in lane q.  Inserts candidate (cs, ci),
// which the caller has already tested against the current K-th best; all 32 lanes call it together.
template <int KL>
__device__ __forceinline__ void warp_list_insert(float (&ls)[KL], int (&li)[KL], float cs, int ci, int lane) {
    // list 0: entries that stay ahead of the candidate form a prefix
    const int pos0 = __popc(__ballot_sync(0xffffffffu, cand_beats(ls[0], li[0], cs, ci)));
    const float fall_s = __shfl_sync(0xffffffffu, ls[0], 31);
    const int fall_i = __shfl_sync(0xffffffffu, li[0], 31);
    float up_s = __shfl_up_sync(0xffffffffu, ls[0], 1);
    int up_i = __shfl_up_sync(0xffffffffu, li[0], 1);
    if (lane == pos0) ls[0] = cs, li[0] = ci;
    else if (lane > pos0) ls[0] = up_s, li[0] = up_i;
    if (KL == 2) {
        if (pos0 < 32) {  // the entry that fell off list 0 enters list 1 at the front
            up_s = __shfl_up_sync(0xffffffffu, ls[KL - 1], 1);
            up_i = __shfl_up_sync(0xffffffffu, li[KL - 1], 1);
            if (lane == 0) ls[KL - 1] = fall_s, li[KL - 1] = fall_i;
            else ls[KL - 1] = up_s, li[KL - 1] = up_i;
        } else {
            const int pos1 = __popc(__ballot_sync(0xffffffffu, cand_beats(ls[KL - 1], li[KL - 1], cs, ci)));
            up_s = __shfl_up_sync(0xffffffffu, ls[KL - 1], 1);
            up_i = __shfl_up_sync(0xffffffffu, li[KL - 1], 1);
            if (lane == pos1) ls[KL - 1] = cs, li[KL - 1] = ci;
            else if (lane > pos1) ls[KL - 1] = up_s, li[KL - 1] = up_i;
        }
    }
}

// rank-select: out[rank] = cand for the K best of n candidates in shared memory (all-pairs counting, with a
// lower bound that skips candidates which cannot be among the K best)
__device__ __forceinline__ void rank_select(const Cand *cands, int n, int K, float bound, Cand *out) {
    for (int q = threadIdx.x; q < n; q += BEAM_NT) {
        const Cand me = cands[q];
        if (me.s < bound) continue;
        int rank = 0;
        for (int o = 0; o < n; ++o) {
            const Cand e = cands[o];
            rank += cand_beats(e.s, e.i, me.s, me.i) ? 1 : 0;
        }
        if (rank < K) out[rank] = me;
    }
}

// Phases 4-6 of a beam step, shared by k_beam_step and k_beam_merge: given the utterance's K = 2W best candidates (sorted,
// top[r].i = source hypothesis * V + token, sentinel entries have i = INT_MAX), finalise eos candidates ranked inside the
// top W into the finished pool, choose the W continuing beams, test `done`, write the reordered + extended rows, and let
// the last CTA of the grid publish (step, #done utterances) to host-visible memory.
// Warp 0 does the bookkeeping with ballots and prefix counts (round 1 ran it serially on one thread: ~1500 dependent
// instructions behind a CTA barrier); lane r owns candidates r and r + 32 and pool slot r.
struct BeamOut {
    float *beam_scores;
    int64_t *best_ids_out, *last_ids_out;
    const int64_t *ids_cur;
    int64_t *ids_next;
    long long ld_ids;
    float *pool_scores;
    int64_t *pool_lens, *pool_seqs;
    long long ld_pool;
    unsigned char *done;
    unsigned int *ticket;
    long long *done_ring;
    int ring;
    long long step_tag;
};

struct BeamShared {
    int job_src[32], job_dst[32], n_pool_jobs;
    int next_tok[32], next_src[32];
    float next_score[32];
    unsigned int last;
    int tot;
};

__device__ __forceinline__ void beam_bookkeep(const Cand *top, BeamShared &bs, const BeamOut &o, int b, int nB, int L, int W, int V, int eos,
                                              int pad, float len_norm) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int K = 2 * W;
    const float NEG = -INFINITY;
    if (tid < 32) {
        const float inv_norm = 1.0f / len_norm;  // torch divides a tensor by a scalar as a * (1 / scalar)
        const bool dn = o.done[b] != 0;
        float cs[2];
        int csrc[2], ctok[2];
        bool cok[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int r = lane + 32 * q;
            cok[q] = r < K && top[r].i != 0x7fffffff;
            const int i = cok[q] ? top[r].i : 0;
            cs[q] = cok[q] ? top[r].s : NEG;
            csrc[q] = i / V;
            ctok[q] = i - csrc[q] * V;
        }
        float ps = lane < W ? o.pool_scores[(size_t)b * W + lane] : INFINITY;  // pool slot `lane`
        int my_job_src = -1;
        // eos candidates ranked inside the top W are finalised, in rank order (W <= 32: they all live in q = 0)
        unsigned eos_mask = __ballot_sync(0xffffffffu, cok[0] && lane < W && ctok[0] == eos && !dn && cs[0] > NEG);
        while (eos_mask) {
            const int r = __ffs(eos_mask) - 1;
            eos_mask &= eos_mask - 1;
            const float fs = __shfl_sync(0xffffffffu, cs[0], r) * inv_norm;
            const int fsrc = __shfl_sync(0xffffffffu, csrc[0], r);
            float wv = ps;  // the worst slot, lowest index on ties
            int wi = lane;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, wv, d);
                const int oi = __shfl_xor_sync(0xffffffffu, wi, d);
                if (ov < wv || (ov == wv && oi < wi)) wv = ov, wi = oi;
            }
            if (fs > wv && lane == wi) ps = fs, my_job_src = fsrc;  // a later candidate may overwrite a slot filled in this same step
        }
        const bool has_job = lane < W && my_job_src >= 0;
        if (has_job) {
            o.pool_scores[(size_t)b * W + lane] = ps;
            o.pool_lens[(size_t)b * W + lane] = L - 1;
        }
        const unsigned job_mask = __ballot_sync(0xffffffffu, has_job);
        if (has_job) {
            const int jq = __popc(job_mask & ((1u << lane) - 1u));
            bs.job_src[jq] = my_job_src, bs.job_dst[jq] = lane;
        }
        if (lane == 0) bs.n_pool_jobs = __popc(job_mask);
        // the first W non-eos candidates continue: candidate (lane, q) goes to position = number of continuing candidates before it
        const unsigned m0 = __ballot_sync(0xffffffffu, cok[0] && ctok[0] != eos);
        const unsigned m1 = __ballot_sync(0xffffffffu, cok[1] && ctok[1] != eos);
        const int n0 = __popc(m0), ncont = n0 + __popc(m1);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const unsigned m = q == 0 ? m0 : m1;
            const int pos = (q == 0 ? 0 : n0) + __popc(m & ((1u << lane) - 1u));
            if (((m >> lane) & 1u) && pos < W)
                bs.next_score[pos] = cs[q], bs.next_tok[pos] = ctok[q], bs.next_src[pos] = csrc[q];
        }
        __syncwarp();
        float ns = NEG;
        int ntok = pad, nsrc = 0;
        if (lane < W && lane < ncont) ns = bs.next_score[lane], ntok = bs.next_tok[lane], nsrc = bs.next_src[lane];
        // done test of the utterance (BeamHypotheses.is_done, early_stopping = False)
        const bool full = __all_sync(0xffffffffu, ps > NEG);
        float worst = ps;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) worst = fminf(worst, __shfl_xor_sync(0xffffffffu, worst, d));
        const float top0 = top[0].i != 0x7fffffff ? top[0].s : NEG;
        const bool dn_new = dn || (full && worst >= top0 * inv_norm);
        __syncwarp();
        if (lane < W) {
            if (dn_new) ns = 0.f, ntok = pad, nsrc = 0;
            o.beam_scores[b * W + lane] = ns;
            // what index_select_state wants (ESPnet ids: source hypothesis * V + token, :180-191)
            if (o.best_ids_out != nullptr) o.best_ids_out[b * W + lane] = (long long)nsrc * V + ntok;
            if (o.last_ids_out != nullptr) o.last_ids_out[b * W + lane] = ntok;
            bs.next_tok[lane] = ntok, bs.next_src[lane] = nsrc;
        }
        if (lane == 0) o.done[b] = dn_new ? 1 : 0;
    }
    __syncthreads();

    // ---- copies: finished prefixes into the pool, reordered + extended rows into ids_next ----------
    // One warp per row, the row's loads issued together before its stores (round 1 walked the W rows one after the other with
    // the whole CTA: W dependent global round trips, 10-20 us of the step at W = 10-20).
    const int nt = blockDim.x, nw = nt >> 5, wid = tid >> 5;
    const int64_t *__restrict__ ids_cur = o.ids_cur;
    auto copy_row = [&](const int64_t *__restrict__ src, int64_t *__restrict__ dst, int n) {
        for (int k0 = 0; k0 < n; k0 += 128) {
            int64_t v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = k0 + u * 32 + lane;
                v[u] = k < n ? src[k] : 0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = k0 + u * 32 + lane;
                if (k < n) dst[k] = v[u];
            }
        }
    };
    const int njobs = bs.n_pool_jobs;
    for (int q = wid; q < njobs + W; q += nw) {
        if (q < njobs) {
            copy_row(ids_cur + ((size_t)b * W + bs.job_src[q]) * o.ld_ids + 1 /* drop bos */,
                     o.pool_seqs + ((size_t)b * W + bs.job_dst[q]) * o.ld_pool, L - 1);
        } else {
            const int w = q - njobs;
            int64_t *dst = o.ids_next + ((size_t)b * W + w) * o.ld_ids;
            copy_row(ids_cur + ((size_t)b * W + bs.next_src[w]) * o.ld_ids, dst, L);
            if (lane == 0) dst[L] = bs.next_tok[w];
        }
    }

    // ---- the last utterance to finish its step publishes (step, #done utterances) to host-visible memory
    if (o.done_ring != nullptr) {
        __threadfence();
        __syncthreads();
        if (tid == 0) bs.last = (atomicAdd(o.ticket, 1u) == (unsigned)(nB - 1)) ? 1u : 0u, bs.tot = 0;
        __syncthreads();
        if (bs.last) {
            int cnt = 0;
            for (int k = tid; k < nB; k += nt) cnt += ((volatile unsigned char *)o.done)[k] ? 1 : 0;
            atomicAdd(&bs.tot, cnt);
            __syncthreads();
            if (tid == 0) {
                *o.ticket = 0;
                o.done_ring[o.step_tag % o.ring] = (o.step_tag << 32) | (long long)bs.tot;
                __threadfence_system();
            }
        }
    }
}

// Beam step over the per-tile candidate lists of the fused scoring kernel (k_psi_full<TOPK>): one CTA per utterance merges
// its nlists sorted lists of K = 2W (key = joint + running beam score, dense index) into the K best and does the
// bookkeeping.  Every list is sorted, so its K-th key bounds the merge from below: entries under the largest such bound
// cannot be among the K best.
constexpr int MERGE_MAX = 2048;  // candidates of an utterance held in shared memory (indices fit 16 bits)
__global__ void __launch_bounds__(BEAM_NT) k_beam_merge(const float2 *__restrict__ lists, int nlists, int L, int W, int V, int eos, int pad,
                                                        float len_norm, BeamOut o) {
    __shared__ float2 cands[MERGE_MAX];
    __shared__ Cand top[BEAM_MAXK];
    __shared__ float bound_s[BEAM_NT / 32];
    __shared__ BeamShared bs;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int K = 2 * W, n = nlists * K;
    const float NEG = -INFINITY;
    const float2 *src = lists + (size_t)b * n;
    // a finished utterance has nothing to rank (its lists may be stale: the scoring kernel skips it): every entry a sentinel
    const bool finished = o.done[b] != 0;
    float bnd = NEG;
    for (int q = tid; q < n; q += BEAM_NT) {
        const float2 c = finished ? make_float2(NEG, __int_as_float(0x7fffffff)) : src[q];
        cands[q] = c;
        if ((q % K) == K - 1 && __float_as_int(c.y) != 0x7fffffff) bnd = fmaxf(bnd, c.x);
    }
    for (int k = tid; k < BEAM_MAXK; k += BEAM_NT) top[k].s = NEG, top[k].i = 0x7fffffff;
    bnd = warp_max(bnd);
    if (lane == 0) bound_s[wid] = bnd;
    __syncthreads();
    float bound = bound_s[0];
#pragma unroll
    for (int q = 1; q < BEAM_NT / 32; ++q) bound = fmaxf(bound, bound_s[q]);
    // Rank of an entry = its position in its own list + the number of entries of every OTHER list that beat it; the lists are
    // sorted by the very order that decides (key descending, index ascending; sentinels last), so that number is a binary
    // search: (nlists - 1) * log2(K) steps per entry instead of n comparisons.  (The first version compared every entry at or
    // above the bound with all n: with garbage-level scores in every tile the bound prunes only half of them, 80 k
    // comparisons per utterance and 40 of the kernel's 57 us at W = 20 -- ncu r3d.)
    for (int q = tid; q < n; q += BEAM_NT) {
        const float2 me = cands[q];
        const int mi = __float_as_int(me.y);
        if (mi == 0x7fffffff || me.x < bound) continue;
        const int lm = q / K;
        int rank = q - lm * K;
        for (int l = 0; l < nlists && rank < K; ++l) {
            if (l == lm) continue;
            const float2 *lst = cands + l * K;
            int lo = 0, hi = K;  // entries [0, lo) beat me, entries [hi, K) do not
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const float2 c = lst[mid];
                const int ci = __float_as_int(c.y);
                if (ci != 0x7fffffff && cand_beats(c.x, ci, me.x, mi)) lo = mid + 1;
                else hi = mid;
            }
            rank += lo;
        }
        if (rank < K) top[rank].s = me.x, top[rank].i = mi;
    }
    __syncthreads();
    beam_bookkeep(top, bs, o, b, (int)gridDim.x, L, W, V, eos, pad, len_norm);
}

// grid = B * P: CTA (b, p) reduces slice p of the W*V candidates of utterance b to its K best, the last CTA of the
// utterance to arrive (ticket) merges the P partial lists and does the bookkeeping and the copies.
// SPARSE: the candidates are the S pre-beam tokens of each hypothesis, `joint` is (BW,S) and cand_ids (BW,S) names them;
// one CTA per utterance (P = 1).  Candidate i of the utterance has the dense index (i / S) * V + cand_ids[i], so ranking,
// tie-breaking and bookkeeping are those of the dense kernel on a joint tensor that is -inf outside the candidates.
template <int KL, bool SPARSE>
__global__ void __launch_bounds__(BEAM_NT) k_beam_step(const float *__restrict__ joint, const int64_t *__restrict__ cand_ids, int S,
                                                       float *beam_scores, int64_t *__restrict__ best_ids_out,
                                                       int64_t *__restrict__ last_ids_out,
                                                       const int64_t *__restrict__ ids_cur, int64_t *__restrict__ ids_next,
                                                       long long ld_ids, int L, int W, int V, int P, int eos, int pad,
                                                       float len_norm, float *pool_scores, int64_t *pool_lens,
                                                       int64_t *pool_seqs, long long ld_pool, unsigned char *done,
                                                       Cand *part, unsigned int *utt_ticket, unsigned int *ticket,
                                                       long long *done_ring, int ring, long long step_tag) {
    __shared__ Cand wl[BEAM_MAXP * BEAM_MAXK];  // per-warp lists (phase 1), then the P partial lists (phase 2)
    __shared__ Cand top[BEAM_MAXK];             // sorted result of a rank_select
    __shared__ float kth[BEAM_MAXP];
    __shared__ BeamShared bs;
    __shared__ unsigned int is_last;
    const int b = blockIdx.x / P, p = blockIdx.x - b * P;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int K = 2 * W;
    const float NEG = -INFINITY;
    const int n = W * V;

    // ---- 1. per-warp top-K of a contiguous slice: sorted list of 32*KL entries held in registers (entry q of list j
    //         lives in lane q), candidates above the current K-th best are inserted with warp shuffles ----------
    Cand *mine = wl + wid * BEAM_MAXK;
    float ls[KL];
    int li[KL];
#pragma unroll
    for (int j = 0; j < KL; ++j) ls[j] = NEG, li[j] = 0x7fffffff;
    float thr = NEG;  // score of entry K-1
    const int thr_lane = (K - 1) & 31, thr_list = (K - 1) >> 5;
    if (SPARSE) {
        // candidates arrive in score order inside a hypothesis, not in index order: compare (score, index) with the K-th
        int thr_i = 0x7fffffff;
        const int ns = W * S;
        const float *cj = joint + (size_t)b * ns;
        const int64_t *ci64 = cand_ids + (size_t)b * ns;
        for (int i0 = wid * 32; i0 < ns; i0 += BEAM_NT) {
            const int i = i0 + lane;
            float c = NEG;
            int ci = 0x7fffffff;
            if (i < ns) {
                const int w = i / S;
                c = cj[i] + beam_scores[b * W + w];
                ci = w * V + (int)ci64[i];
            }
            unsigned m = __ballot_sync(0xffffffffu, cand_beats(c, ci, thr, thr_i));
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const float cs = __shfl_sync(0xffffffffu, c, src);
                const int cs_i = __shfl_sync(0xffffffffu, ci, src);
                if (cand_beats(cs, cs_i, thr, thr_i)) {
                    warp_list_insert<KL>(ls, li, cs, cs_i, lane);
                    thr = __shfl_sync(0xffffffffu, (KL == 2 && thr_list) ? ls[KL - 1] : ls[0], thr_lane);
                    thr_i = __shfl_sync(0xffffffffu, (KL == 2 && thr_list) ? li[KL - 1] : li[0], thr_lane);
                }
            }
        }
    } else {
        const int per_cta = (((n + P - 1) / P + 127) / 128) * 128;
        const int per_warp = per_cta / BEAM_NW;  // multiple of 32
        const int s0 = min(n, p * per_cta + wid * per_warp), e0 = min(n, s0 + per_warp);
        const float *flat = joint + (size_t)b * n;
        int w = s0 / V;
        for (int s = s0; s < e0;) {
            const int e = min(e0, (w + 1) * V);  // segment [s, e) lies in hypothesis w
            const float bs = beam_scores[b * W + w];
            constexpr int U = 8;  // independent 128-byte loads in flight per warp
            for (int vb0 = s; vb0 < e; vb0 += 32 * U) {
                float cu[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = vb0 + u * 32 + lane;
                    cu[u] = i < e ? __ldg(flat + i) + bs : NEG;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float c = cu[u];
                    unsigned m = __ballot_sync(0xffffffffu, c > thr);
                    while (m) {
                        const int src = __ffs(m) - 1;
                        m &= m - 1;
                        const float cs = __shfl_sync(0xffffffffu, c, src);
                        if (cs > thr) {  // strict: on ties the earlier (lower index) candidate stays
                            warp_list_insert<KL>(ls, li, cs, vb0 + u * 32 + src, lane);
                            thr = __shfl_sync(0xffffffffu, (KL == 2 && thr_list) ? ls[KL - 1] : ls[0], thr_lane);
                        }
                    }
                }
            }
            s = e;
            ++w;
        }
    }
#pragma unroll
    for (int j = 0; j < KL; ++j) mine[j * 32 + lane].s = ls[j], mine[j * 32 + lane].i = li[j];
    if (KL == 1) mine[32 + lane].s = NEG, mine[32 + lane].i = 0x7fffffff;
    if (lane == 0) kth[wid] = thr;
    __syncthreads();

    // ---- 2. CTA merge of the warp lists -> this slice's K best, published for the utterance's last CTA --
    {
        float bound = kth[0];
        for (int q = 1; q < BEAM_NW; ++q) bound = fmaxf(bound, kth[q]);
        for (int k = tid; k < BEAM_MAXK; k += BEAM_NT) top[k].s = NEG, top[k].i = 0x7fffffff;
        __syncthreads();
        rank_select(wl, BEAM_NW * BEAM_MAXK, K, bound, top);
        __syncthreads();
        Cand *dst = part + ((size_t)b * P + p) * BEAM_MAXK;
        for (int k = tid; k < K; k += BEAM_NT) dst[k] = top[k];
        __threadfence();
        __syncthreads();
        if (tid == 0) is_last = (atomicAdd(utt_ticket + b, 1u) == (unsigned)(P - 1)) ? 1u : 0u;
        __syncthreads();
        if (!is_last) return;
    }

    // ---- 3. last CTA of the utterance: merge the P partial lists ---------------------------------------
    {
        const volatile Cand *src = part + (size_t)b * P * BEAM_MAXK;
        for (int q = tid; q < P * BEAM_MAXK; q += BEAM_NT) {
            Cand c;
            if ((q % BEAM_MAXK) < K) c.s = src[q].s, c.i = src[q].i;
            else c.s = NEG, c.i = 0x7fffffff;
            wl[q] = c;
        }
        __syncthreads();
        if (tid < P) kth[tid] = wl[tid * BEAM_MAXK + K - 1].s;  // partial lists are sorted: entry K-1 is the slice's K-th
        __syncthreads();
        float bound = kth[0];
        for (int q = 1; q < P; ++q) bound = fmaxf(bound, kth[q]);
        rank_select(wl, P * BEAM_MAXK, K, bound, top);
        __syncthreads();
    }

    // ---- 4.-6. bookkeeping, copies, done ring ---------------------------------------------------------
    if (tid == 0) utt_ticket[b] = 0;
    BeamOut o;
    o.beam_scores = beam_scores, o.best_ids_out = best_ids_out, o.last_ids_out = last_ids_out, o.ids_cur = ids_cur, o.ids_next = ids_next;
    o.ld_ids = ld_ids, o.pool_scores = pool_scores, o.pool_lens = pool_lens, o.pool_seqs = pool_seqs, o.ld_pool = ld_pool, o.done = done;
    o.ticket = ticket, o.done_ring = done_ring, o.ring = ring, o.step_tag = step_tag;
    beam_bookkeep(top, bs, o, b, (int)gridDim.x / P, L, W, V, eos, pad, len_norm);
}

#include "ctcps_prebeam.cuh"

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// Widest hypothesis group of the materialising kernel (CTCPS_SCORE_MAX_GROUP, 2..5).  Default 2: the kernel is bound by the
// store stream of r, so what counts is how many CTAs keep stores in flight, not how often the x tile is re-read from L2:
// measured (profiles/r2p_score_group_ab.md) C1 0.584 -> 0.734 of the HBM peak (320 -> 800 tiles on 592 slots), C2 0.860 -> 0.868.
constexpr int SCORE_DEFAULT_GROUP = 2;
int g_score_max_group = -1;
int score_max_group() {
    if (g_score_max_group < 0) {
        const char *ev = getenv("CTCPS_SCORE_MAX_GROUP");
        g_score_max_group = ev != nullptr ? atoi(ev) : SCORE_DEFAULT_GROUP;
        if (g_score_max_group < 2 || g_score_max_group > MAX_HW) g_score_max_group = SCORE_DEFAULT_GROUP;
    }
    return g_score_max_group;
}
void pick_hw(int W, int *HW, int *G) {
    int best = 1, best_pad = 1 << 30;
    for (int hw = score_max_group(); hw >= 1; --hw) {
        const int g = (W + hw - 1) / hw;
        const int pad = g * hw - W;
        // prefer no padded lanes, then the widest group (more reuse of x per shared-memory read)
        if (pad < best_pad && hw >= 2) best = hw, best_pad = pad;
        if (pad == 0 && hw >= 2) break;
    }
    if (W == 1) best = 1;
    *HW = best;
    *G = (W + best - 1) / best;
}

inline int tpad_of(int T) { return ((T + TT - 1) / TT) * TT + TT; }

struct Workspace {
    size_t aux_off, aux_bytes, g_off, g_bytes, total;
};
Workspace plan_workspace(int B, int T, int W) {
    int HW, G;
    pick_hw(W, &HW, &G);
    Workspace ws;
    ws.aux_off = 0;
    ws.aux_bytes = (size_t)B * G * tpad_of(T) * (HW + 1) * sizeof(float4);
    ws.g_off = (ws.aux_bytes + 255) & ~(size_t)255;
    ws.g_bytes = (size_t)B * W * sizeof(float);
    ws.total = ws.g_off + ((ws.g_bytes + 255) & ~(size_t)255);
    return ws;
}

// lazy mode: hypotheses per thread of k_psi_full (even: packed FMAs pair two hypotheses) and the padded width of its lin
// stream.  Up to 20 hypotheses in one group (W = 20 reads and exponentiates its x tile once); wider beams are split into
// the group size with the fewest padded lanes.
int g_psi_max_group = -1;  // widest hypothesis group k_psi_full may use (CTCPS_PSI_MAX_GROUP / ctcps_set_psi_max_group; default 20)
int psi_max_group() {
    if (g_psi_max_group < 0) {
        const char *ev = getenv("CTCPS_PSI_MAX_GROUP");
        g_psi_max_group = ev != nullptr ? atoi(ev) : 20;
        if (g_psi_max_group < 2 || g_psi_max_group > 20) g_psi_max_group = 20;
    }
    return g_psi_max_group;
}
void pick_hw_psi(int W, int *HW, int *HWP, int *G) {
    static const int all[] = {20, 16, 12, 10, 8, 6, 4, 2};
    int cand[8], nc = 0;
    for (int hw : all)
        if (hw <= psi_max_group() || hw == 2) cand[nc++] = hw;
    int best = 2;
    if (W <= cand[0]) {
        for (int i = 0; i < nc; ++i)
            if (cand[i] >= W) best = cand[i];  // the smallest group that holds every hypothesis
    } else {
        int best_pad = 1 << 30;
        for (int i = 0; i < nc; ++i) {
            const int hw = cand[i];
            const int pad = ((W + hw - 1) / hw) * hw - W;
            if (pad < best_pad) best = hw, best_pad = pad;  // ties: the widest group (listed first)
        }
    }
    *HW = best;
    *HWP = (best + 3) & ~3;
    *G = (W + best - 1) / best;
}
struct WorkspaceLazy {
    size_t lin_off, lin_bytes, g_off, c_off, r_off, total;
};
WorkspaceLazy plan_workspace_lazy(int B, int T, int W, int V) {
    (void)V;
    int HW, HWP, G;
    pick_hw_psi(W, &HW, &HWP, &G);
    WorkspaceLazy ws;
    ws.lin_off = 0;
    ws.lin_bytes = (size_t)B * G * tpad_of(T) * HWP * sizeof(float);
    ws.g_off = (ws.lin_bytes + 255) & ~(size_t)255;
    const size_t gb = ((size_t)B * W * sizeof(float) + 255) & ~(size_t)255;
    ws.c_off = ws.g_off + gb;
    ws.r_off = ws.c_off + gb;  // (B*G) int2: chunk range of the nonzero part of the lin stream (k_lin_range)
    ws.total = ws.r_off + (((size_t)B * G * sizeof(int2) + 255) & ~(size_t)255);
    return ws;
}

int cuda_rc(cudaError_t e) { return e == cudaSuccess ? 0 : (int)e; }

#define ARG_CHECK(cond, code, msg)                           \
    do {                                                     \
        if (!(cond)) {                                       \
            snprintf(g_errbuf, sizeof(g_errbuf), "%s", msg); \
            return code;                                     \
        }                                                    \
    } while (0)

// SM count of the CURRENT device, cached per device ordinal (a process may drive several GPUs).
constexpr int MAX_DEVICES = 64;
int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev;
}
int sm_count() {
    static int cache[MAX_DEVICES] = {0};
    const int dev = current_device();
    if (dev >= 0 && dev < MAX_DEVICES && cache[dev] > 0) return cache[dev];
    int sms = 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    if (dev >= 0 && dev < MAX_DEVICES) cache[dev] = sms;
    return sms;
}

int grid_for(size_t n, int block) {
    size_t g = (n + block - 1) / block;
    const size_t cap = (size_t)sm_count() * 32;
    return (int)(g < cap ? (g ? g : 1) : cap);
}

// Resident CTAs of a kernel instantiation on the whole current device; one cache slot per (instantiation, device).
template <typename K>
int resident_slots(K kern, int nthreads, size_t smem, int *slots_cache /* [MAX_DEVICES] */, int *out) {
    const int dev = current_device();
    if (dev >= 0 && dev < MAX_DEVICES && slots_cache[dev] > 0) {
        *out = slots_cache[dev];
        return 0;
    }
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nthreads, smem);
    if (e != cudaSuccess) return (int)e;
    const int slots = sm_count() * (per_sm < 1 ? 1 : per_sm);
    if (dev >= 0 && dev < MAX_DEVICES) slots_cache[dev] = slots;
    *out = slots;
    return 0;
}

// Launch policy (measured on B200, tools/kernel_bench.py): the state-writing kernel is fastest with one CTA per tile
// handed out by the hardware scheduler (7.09 ms vs 7.87 ms persistent at C2: statically striding CTAs spread the store
// window); the read-only psi kernel is fastest persistent (422 us vs 435-505 us).
template <int HW, int NT, int MINB>
int launch_score_full(const CUtensorMap &tm, const ScoreArgs &a, cudaStream_t st) {
    using Smem = ScoreSmem<HW, NT>;
    auto kern = k_score_full<HW, NT, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    if (e != cudaSuccess) return (int)e;
    const long long ntiles = (long long)a.B * a.nvt * a.G;
    kern<<<(unsigned)ntiles, NT, sizeof(Smem), st>>>(tm, a);
    return cuda_rc(cudaGetLastError());
}

template <int HW, int HWP, int NT, int MINB, int NSTAGE, bool TOPK>
int launch_psi_full(const CUtensorMap &tm, const PsiArgs &a, cudaStream_t st) {
    using Smem = PsiSmemAll<HW, HWP, NT, NSTAGE, TOPK>;
    auto kern = k_psi_full<HW, HWP, NT, MINB, NSTAGE, TOPK>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    if (e != cudaSuccess) return (int)e;
    const long long ntiles = (long long)a.B * a.nvt * a.G;
    static int slots_cache[MAX_DEVICES] = {0};
    int slots = 0;
    const int rc = resident_slots(kern, NT, sizeof(Smem), slots_cache, &slots);
    if (rc) return rc;
    long long grid = slots;
    if (grid > ntiles) grid = ntiles;
    kern<<<(unsigned)grid, NT, sizeof(Smem), st>>>(tm, a);
    return cuda_rc(cudaGetLastError());
}

// HW <= 10: 4 CTAs per SM (128 registers) with a 3-stage ring; wider groups: 3 CTAs per SM (168 registers), 4 stages.
template <bool TOPK>
int dispatch_psi_full(int HW, const CUtensorMap &tm, const PsiArgs &a, cudaStream_t st) {
    constexpr int NT = PSI_NT;
    switch (HW) {
        case 2: return launch_psi_full<2, 4, NT, 4, 3, TOPK>(tm, a, st);
        case 4: return launch_psi_full<4, 4, NT, 4, 3, TOPK>(tm, a, st);
        case 6: return launch_psi_full<6, 8, NT, 4, 3, TOPK>(tm, a, st);
        case 8: return launch_psi_full<8, 8, NT, 4, 3, TOPK>(tm, a, st);
        case 10: return launch_psi_full<10, 12, NT, 4, 3, TOPK>(tm, a, st);
        case 12: return launch_psi_full<12, 12, NT, 3, 4, TOPK>(tm, a, st);
        case 16: return launch_psi_full<16, 16, NT, 3, 4, TOPK>(tm, a, st);
        default: return launch_psi_full<20, 20, NT, 3, 4, TOPK>(tm, a, st);
    }
}

// 1 = time-parallel warp per hypothesis (k_select_lazy_pscan, the default since it passed its fp64-adjudicated hardware
// gate and measured faster on every BASELINE shape: DESIGN.md section 4), 0 = one thread per hypothesis walks the
// T frames (k_select_lazy_scan); ctcps_set_select_pscan / CTCPS_SELECT_PSCAN
int g_select_pscan = -1;
int select_pscan_mode() {
    if (g_select_pscan < 0) {
        const char *ev = getenv("CTCPS_SELECT_PSCAN");
        g_select_pscan = (ev != nullptr && ev[0] == '0') ? 0 : 1;
    }
    return g_select_pscan;
}

// chunks (8 frames x 512 tokens = 16 KB) that k_psi_full prefetches into L2 beyond its shared-memory ring;
// CTCPS_PSI_PREFETCH / ctcps_set_psi_prefetch.  Default 0: measured on B200 (profiles/r2d_psi_ab.md) the look-ahead does not
// help -- 0 / 2 / 4 / 8 chunks: 0.371 / 0.376 / 0.390 / 0.489 ms per C2 scoring call -- the queue depth is not what limits the stream.
int g_psi_prefetch = -1;
int psi_prefetch_chunks() {
    if (g_psi_prefetch < 0) {
        const char *ev = getenv("CTCPS_PSI_PREFETCH");
        g_psi_prefetch = ev != nullptr ? atoi(ev) : 0;
        if (g_psi_prefetch < 0 || g_psi_prefetch > 64) g_psi_prefetch = 0;
    }
    return g_psi_prefetch;
}

// Native decode loop: do not score the utterances whose beam search has finished (CTCPS_SKIP_DONE, default 1).  Their rows
// never reach a hypothesis (the beam step ignores a finished utterance), so the returned n-best lists are bit-identical.
// Lazy scoring kernel: stream only the chunks in which the lin stream of a hypothesis group is nonzero (CTCPS_FRAME_WINDOW,
// default 1; see k_lin_range).  Bit-identical either way.
int g_frame_window = -1;
bool frame_window() {
    if (g_frame_window < 0) {
        const char *ev = getenv("CTCPS_FRAME_WINDOW");
        g_frame_window = (ev != nullptr && atoi(ev) == 0) ? 0 : 1;
    }
    return g_frame_window != 0;
}
unsigned long long *g_stream_counter = nullptr;  // device word the lazy scoring kernel adds its streamed chunks to (bench.py)

int g_skip_done = -1;
bool skip_done() {
    if (g_skip_done < 0) {
        const char *ev = getenv("CTCPS_SKIP_DONE");
        g_skip_done = (ev != nullptr && atoi(ev) == 0) ? 0 : 1;
    }
    return g_skip_done != 0;
}

// The descriptor only depends on (pointer, ldx, B*T, V): a decode makes one scoring call per output token on the same
// posteriors, so the last few descriptors are kept per host thread instead of calling the driver every step.
struct MapCacheEntry {
    const float *p;
    int ldx, rows, V;
    CUtensorMap tm;
};
int encode_x_map_uncached(CUtensorMap *tm, const float *x_logp, int ldx, int B, int T, int V);
int encode_x_map(CUtensorMap *tm, const float *x_logp, int ldx, int B, int T, int V) {
    constexpr int N = 4;
    thread_local MapCacheEntry cache[N] = {};
    thread_local int next = 0;
    const int rows = B * T;
    for (int i = 0; i < N; ++i)
        if (cache[i].p == x_logp && cache[i].ldx == ldx && cache[i].rows == rows && cache[i].V == V) {
            *tm = cache[i].tm;
            return 0;
        }
    const int rc = encode_x_map_uncached(tm, x_logp, ldx, B, T, V);
    if (rc) return rc;
    cache[next].p = x_logp, cache[next].ldx = ldx, cache[next].rows = rows, cache[next].V = V, cache[next].tm = *tm;
    next = (next + 1) % N;
    return 0;
}
int encode_x_map_uncached(CUtensorMap *tm, const float *x_logp, int ldx, int B, int T, int V) {
    EncodeTiledFn enc = get_encode();
    ARG_CHECK(enc != nullptr, CTCPS_E_NODRIVER, "cuTensorMapEncodeTiled not found");
    cuuint64_t dims[2] = {(cuuint64_t)V, (cuuint64_t)B * T};
    cuuint64_t strides[1] = {(cuuint64_t)ldx * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)BOXC, (cuuint32_t)TT};
    cuuint32_t estr[2] = {1, 1};
    CUresult cr = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(x_logp), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        snprintf(g_errbuf, sizeof(g_errbuf), "cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
        return CTCPS_E_NODRIVER;
    }
    return 0;
}

// index_select_state on a state that was never written (see k_select_lazy_*): shared by the full-vocabulary entry point
// (frame-major x, s_new from the dense log_psi) and the pre-beam one (token-major x, s_new from the candidate scores).
int select_lazy_impl(const XView x, const float *blank_lp, const float *r_prev, const int64_t *last_ids, int ol, const float *scores,
                     const int64_t *cand_ids, int S, const int64_t *best_ids, int B, int W, int T, int V, float *r_new, float *s_new,
                     void *next_workspace, size_t next_workspace_bytes, cudaStream_t st) {
    ARG_CHECK(blank_lp && r_prev && last_ids && scores && best_ids && r_new && s_new, CTCPS_E_BADARG, "select_lazy: null pointer");
    ARG_CHECK(B > 0 && W > 0 && T > 0 && V > 0 && ol >= 0, CTCPS_E_BADARG, "select_lazy: bad size");
    const int BW = B * W;
    const dim3 grid((BW + 127) / 128, (T + LAZY_TC - 1) / LAZY_TC);
    k_select_lazy_stage<<<grid, 128, 0, st>>>(x, r_prev, last_ids, ol, best_ids, cand_ids, S, B, W, T, V, r_new);
    // time-parallel scan: needs at least one frame to walk and a tile of PS_H columns that fits shared memory
    const int ps_start = ol > 1 ? ol : 1;
    const int ps_n = T - ps_start;
    const int ps_F = (ps_n + 31) / 32, ps_Fo = ps_F | 1, ps_HS = 2 * 32 * ps_Fo + 1;
    const size_t ps_smem = (size_t)PS_H * ps_HS * sizeof(float);
    const bool pscan = select_pscan_mode() && ps_n >= 1 && ps_smem <= 200 * 1024;
    if (pscan) {
        cudaError_t e = cudaSuccess;
        if (ps_smem > 48 * 1024) {
            e = cudaFuncSetAttribute(k_select_lazy_pscan<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ps_smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_select_lazy_pscan<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ps_smem);
            if (e != cudaSuccess) return (int)e;
        }
    }
    if (pscan && next_workspace != nullptr && ol + 1 <= T) {
        int HW, HWP, G;
        pick_hw_psi(W, &HW, &HWP, &G);
        const WorkspaceLazy ws = plan_workspace_lazy(B, T, W, V);
        ARG_CHECK(next_workspace_bytes >= ws.total, CTCPS_E_WORKSPACE, "select_lazy: workspace too small");
        ARG_CHECK((((uintptr_t)next_workspace) & 255) == 0, CTCPS_E_ALIGN, "select_lazy: workspace must be 256-byte aligned");
        float *lin = reinterpret_cast<float *>((char *)next_workspace + ws.lin_off);
        float *Gmax = reinterpret_cast<float *>((char *)next_workspace + ws.g_off);
        float *psic = reinterpret_cast<float *>((char *)next_workspace + ws.c_off);
        k_select_lazy_pscan<true><<<(BW + PS_H - 1) / PS_H, PS_H * 32, ps_smem, st>>>(x, blank_lp, ol, scores, best_ids, cand_ids, S, B, W, T, V,
                                                                                  r_new, s_new, lin, Gmax, psic, HW, HWP, G, tpad_of(T),
                                                                                  ps_F, ps_Fo, ps_HS);
        k_lin_range<<<B * G, 128, 0, st>>>(lin, tpad_of(T), HWP, reinterpret_cast<int2 *>((char *)next_workspace + ws.r_off));
    } else if (pscan) {
        k_select_lazy_pscan<false><<<(BW + PS_H - 1) / PS_H, PS_H * 32, ps_smem, st>>>(x, blank_lp, ol, scores, best_ids, cand_ids, S, B, W, T, V,
                                                                                   r_new, s_new, nullptr, nullptr, nullptr, 1, 4, 1, 0, ps_F,
                                                                                   ps_Fo, ps_HS);
    } else if (next_workspace != nullptr && ol + 1 <= T) {  // also prepare the next scoring call (it may then pass workspace_prepared = 1)
        int HW, HWP, G;
        pick_hw_psi(W, &HW, &HWP, &G);
        const WorkspaceLazy ws = plan_workspace_lazy(B, T, W, V);
        ARG_CHECK(next_workspace_bytes >= ws.total, CTCPS_E_WORKSPACE, "select_lazy: workspace too small");
        ARG_CHECK((((uintptr_t)next_workspace) & 255) == 0, CTCPS_E_ALIGN, "select_lazy: workspace must be 256-byte aligned");
        float *lin = reinterpret_cast<float *>((char *)next_workspace + ws.lin_off);
        float *Gmax = reinterpret_cast<float *>((char *)next_workspace + ws.g_off);
        float *psic = reinterpret_cast<float *>((char *)next_workspace + ws.c_off);
        k_select_lazy_scan<true><<<(BW + 63) / 64, 64, 0, st>>>(x, blank_lp, ol, scores, best_ids, cand_ids, S, B, W, T, V, r_new, s_new, lin,
                                                               Gmax, psic, HW, HWP, G, tpad_of(T));
        k_lin_range<<<B * G, 128, 0, st>>>(lin, tpad_of(T), HWP, reinterpret_cast<int2 *>((char *)next_workspace + ws.r_off));
    } else {
        k_select_lazy_scan<false><<<(BW + 63) / 64, 64, 0, st>>>(x, blank_lp, ol, scores, best_ids, cand_ids, S, B, W, T, V, r_new, s_new,
                                                                nullptr, nullptr, nullptr, 1, 4, 1, 0);
    }
    return cuda_rc(cudaGetLastError());
}

// top-S of every row of a (BW,V) score matrix; blank < 0: no scores[:, blank] = logzero side effect
int launch_prebeam_topk(float *att_scores, int BW, int V, int blank, int S, int64_t *scoring_ids, float *cand_att, cudaStream_t st) {
    const size_t smem = (size_t)V * sizeof(unsigned);
    ARG_CHECK(smem <= 200 * 1024, CTCPS_E_TOOBIG, "prebeam_topk: vocabulary too large for the shared-memory row (V <= 51200)");
    const bool fast = (V & 3) == 0 && (V >> 2) <= TOPK_NT * TOPK_MAXU && (((uintptr_t)att_scores) & 15) == 0;
    const int U = fast ? ((V >> 2) + TOPK_NT - 1) / TOPK_NT : 0;
#define CTCPS_TOPK_LAUNCH(UU)                                                                                              \
    do {                                                                                                                   \
        if (smem > 40 * 1024) {                                                                                            \
            cudaError_t e = cudaFuncSetAttribute(k_prebeam_topk<UU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return (int)e;                                                                           \
        }                                                                                                                  \
        k_prebeam_topk<UU><<<BW, TOPK_NT, smem, st>>>(att_scores, V, blank, S, scoring_ids, cand_att);                     \
    } while (0)
    switch (U) {
        case 1: CTCPS_TOPK_LAUNCH(1); break;
        case 2: CTCPS_TOPK_LAUNCH(2); break;
        case 3: CTCPS_TOPK_LAUNCH(3); break;
        case 4: CTCPS_TOPK_LAUNCH(4); break;
        case 5: CTCPS_TOPK_LAUNCH(5); break;
        case 6: CTCPS_TOPK_LAUNCH(6); break;
        case 7: CTCPS_TOPK_LAUNCH(7); break;
        case 8: CTCPS_TOPK_LAUNCH(8); break;
        default: CTCPS_TOPK_LAUNCH(0); break;
    }
#undef CTCPS_TOPK_LAUNCH
    return cuda_rc(cudaGetLastError());
}

// workspace of the beam step: partial lists of the one-kernel dense path, tickets, and the per-hypothesis top-2W lists of the
// two-kernel dense path
struct BeamWorkspace {
    size_t part_off, ticket_off, cid_off, cval_off, total;
};
BeamWorkspace plan_beam_workspace(int B, int W) {
    BeamWorkspace w;
    w.part_off = 0;
    w.ticket_off = (size_t)B * BEAM_MAXP * BEAM_MAXK * sizeof(Cand);
    w.cid_off = (w.ticket_off + ((size_t)B + 1) * sizeof(unsigned int) + 15) & ~(size_t)15;
    w.cval_off = w.cid_off + (size_t)B * W * 2 * W * sizeof(int64_t);
    w.total = w.cval_off + (size_t)B * W * 2 * W * sizeof(float);
    return w;
}

int beam_step_impl(const float *joint, const int64_t *cand_ids, int S, float *beam_scores, const int64_t *ids_cur, int64_t *ids_next,
                   int64_t ld_ids, int L, int B, int W, int V, int eos, int pad, float len_norm, float *pool_scores, int64_t *pool_lens,
                   int64_t *pool_seqs, int64_t ld_pool, unsigned char *done, void *workspace, size_t workspace_bytes,
                   int64_t *done_ring, int ring, int64_t step_tag, int64_t *best_ids_out, int64_t *last_ids_out, cudaStream_t st) {
    ARG_CHECK(joint && beam_scores && ids_cur && ids_next && pool_scores && pool_lens && pool_seqs && done && workspace,
              CTCPS_E_BADARG, "beam_step: null pointer");
    ARG_CHECK(B > 0 && W > 0 && V > 0 && L >= 1 && L < ld_ids && L - 1 <= ld_pool, CTCPS_E_BADARG, "beam_step: bad size");
    ARG_CHECK(2 * W <= BEAM_MAXK && W <= 32, CTCPS_E_TOOBIG, "beam_step: num_beams > 32 is not supported");
    ARG_CHECK((long long)W * V < (1ll << 31) && (long long)W * V >= 2 * W, CTCPS_E_TOOBIG, "beam_step: need 2W <= W*V < 2^31");
    ARG_CHECK(done_ring == nullptr || ring > 0, CTCPS_E_BADARG, "beam_step: done_ring without ring size");
    const BeamWorkspace bw = plan_beam_workspace(B, W);
    ARG_CHECK(workspace_bytes >= bw.total, CTCPS_E_WORKSPACE, "beam_step: workspace too small");
    ARG_CHECK((((uintptr_t)workspace) & 15) == 0, CTCPS_E_ALIGN, "beam_step: workspace must be 16-byte aligned");
    Cand *part = reinterpret_cast<Cand *>((char *)workspace + bw.part_off);
    unsigned int *utt_ticket = reinterpret_cast<unsigned int *>((char *)workspace + bw.ticket_off);
    unsigned int *ticket = utt_ticket + B;
    // Dense scores whose rows fit the register top-k: first the 2W best of every hypothesis row (the overall top 2W of the
    // utterance are among them), then the candidate kernel over those W * 2W -- 21 + 27 us at C2 instead of 88 us for the
    // one-kernel path below (serialised warp-wide insertions over W*V scores).  Same ranking, same ties, same bits.
    if (cand_ids == nullptr && (V & 3) == 0 && (V >> 2) <= TOPK_NT * TOPK_MAXU && 2 * W <= V && (((uintptr_t)joint) & 15) == 0) {
        int64_t *cid = reinterpret_cast<int64_t *>((char *)workspace + bw.cid_off);
        float *cval = reinterpret_cast<float *>((char *)workspace + bw.cval_off);
        const int rc = launch_prebeam_topk(const_cast<float *>(joint), B * W, V, -1, 2 * W, cid, cval, st);
        if (rc) return rc;
        joint = cval, cand_ids = cid, S = 2 * W;
    }
    int P = 1;
    if (cand_ids == nullptr) {
        P = (sm_count() * 8 + B - 1) / B;  // about 8 CTAs of 128 threads per SM in one wave
        P = P < 1 ? 1 : (P > BEAM_MAXP ? BEAM_MAXP : P);
        while (P > 1 && (long long)W * V / P < 4 * 2 * W) --P;  // keep slices much longer than K
    }
#define CTCPS_BEAM_LAUNCH(KL, SP)                                                                                                  \
    k_beam_step<KL, SP><<<B * P, BEAM_NT, 0, st>>>(joint, cand_ids, S, beam_scores, best_ids_out, last_ids_out, ids_cur, ids_next, ld_ids, L, W, V, P, \
                                                   eos, pad, len_norm, pool_scores, pool_lens, pool_seqs, ld_pool, done, part,      \
                                                   utt_ticket, ticket, (long long *)done_ring, ring, step_tag)
    if (cand_ids != nullptr) {
        if (2 * W <= 32) CTCPS_BEAM_LAUNCH(1, true);
        else CTCPS_BEAM_LAUNCH(2, true);
    } else {
        if (2 * W <= 32) CTCPS_BEAM_LAUNCH(1, false);
        else CTCPS_BEAM_LAUNCH(2, false);
    }
#undef CTCPS_BEAM_LAUNCH
    return cuda_rc(cudaGetLastError());
}

}  // namespace

extern "C" {

int ctcps_version(void) { return 120; }  // 120: score_window, score_lazy_lens, score_lazy_topk_active, frame window, 3xFP16 head

const char *ctcps_error_string(int code) {
    if (code == 0) return "ok";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    switch (code) {
        case CTCPS_E_BADARG: return g_errbuf[0] ? g_errbuf : "bad argument";
        case CTCPS_E_ALIGN: return g_errbuf[0] ? g_errbuf : "alignment violated";
        case CTCPS_E_WORKSPACE: return "workspace too small (see ctcps_workspace_bytes)";
        case CTCPS_E_NODRIVER: return "cuTensorMapEncodeTiled unavailable or failed";
        case CTCPS_E_TOOBIG: return "problem too large for 32-bit lane indexing";
        default: return "unknown ctcps error";
    }
}

// 64 floats = 256 B: measured on B200 (tools/membw2.py, profiles/r1_write_pattern.md) the store stream of r runs at
// 6.5 TB/s with 256-byte-aligned rows and at 5.3 TB/s with rows that are only 32-byte aligned (ld = 5000).
int ctcps_padded_ld(int n) { return (n + 63) & ~63; }

int ctcps_set_select_pscan(int mode) {
    const int prev = select_pscan_mode();
    if (mode == 0 || mode == 1) g_select_pscan = mode;
    return prev;
}

int ctcps_set_psi_max_group(int hyps) {
    const int prev = psi_max_group();
    if (hyps >= 2 && hyps <= 20) g_psi_max_group = hyps;
    return prev;
}

int ctcps_set_psi_prefetch(int chunks) {
    const int prev = psi_prefetch_chunks();
    if (chunks >= 0 && chunks <= 64) g_psi_prefetch = chunks;
    return prev;
}

int ctcps_set_frame_window(int on) {
    const int prev = frame_window() ? 1 : 0;
    if (on == 0 || on == 1) g_frame_window = on;
    return prev;
}

int ctcps_set_stream_counter(void *device_u64) {
    g_stream_counter = reinterpret_cast<unsigned long long *>(device_u64);
    return 0;
}

int ctcps_stream_chunk_bytes(int W) {
    int HW, HWP, G;
    pick_hw_psi(W, &HW, &HWP, &G);
    return TT * PSI_NT * 4 * 4 + TT * HWP * 4;  // one chunk of a tile: 8 frames x 512 tokens of x + 8 frames of the group's lin
}

int ctcps_set_skip_done(int on) {
    const int prev = skip_done() ? 1 : 0;
    if (on == 0 || on == 1) g_skip_done = on;
    return prev;
}

int ctcps_workspace_bytes(int B, int T, int V, int W, int S, size_t *out_bytes) {
    (void)V;
    (void)S;
    ARG_CHECK(out_bytes != nullptr && B > 0 && T > 0 && W > 0, CTCPS_E_BADARG, "workspace_bytes: bad sizes");
    const size_t a = plan_workspace(B, T, W).total, l = plan_workspace_lazy(B, T, W, V).total;
    *out_bytes = a > l ? a : l;
    return 0;
}

int ctcps_init(const float *logits, int ld_in, const int64_t *lens, int B, int T, int V, int blank, int apply_log_softmax,
               float *x_logp, int ldx, float *blank_lp, void *stream) {
    ARG_CHECK(logits && x_logp && B > 0 && T > 0 && V > 0, CTCPS_E_BADARG, "init: null pointer or non-positive size");
    ARG_CHECK(blank >= 0 && blank < V, CTCPS_E_BADARG, "init: blank id outside the vocabulary");
    ARG_CHECK(ld_in >= V && ldx >= V, CTCPS_E_BADARG, "init: row stride smaller than V");
    ARG_CHECK((long long)B * T < (1ll << 31), CTCPS_E_TOOBIG, "init: B*T too large");
    k_init<<<B * T, 256, 0, (cudaStream_t)stream>>>(logits, ld_in, lens, T, V, blank, apply_log_softmax, x_logp, ldx, blank_lp);
    return cuda_rc(cudaGetLastError());
}

int ctcps_log_softmax(const float *in, int ld_in, float *out, int ld_out, int rows, int V, void *stream) {
    ARG_CHECK(in && out && rows > 0 && V > 0 && ld_in >= V && ld_out >= V, CTCPS_E_BADARG, "log_softmax: bad argument");
    k_init<<<rows, 256, 0, (cudaStream_t)stream>>>(in, ld_in, nullptr, rows, V, 0, 1, out, ld_out, nullptr);
    return cuda_rc(cudaGetLastError());
}

int ctcps_initial_state(const float *blank_lp, int B, int T, int W, int t_begin, float *r0, void *stream) {
    ARG_CHECK(blank_lp && r0 && B > 0 && T > 0 && W > 0 && t_begin >= 0 && t_begin <= T, CTCPS_E_BADARG,
              "initial_state: bad argument");
    const int BW = B * W;
    k_initial_state<<<(BW + 127) / 128, 128, 0, (cudaStream_t)stream>>>(blank_lp, B, T, W, t_begin, r0);
    return cuda_rc(cudaGetLastError());
}

int ctcps_score_window(const float *x_logp, int ldx, const float *blank_lp, const float *r_prev, const float *s_prev,
                       int64_t s_row_stride, int64_t s_col_stride, const int64_t *last_ids, int ol, int start, int end, int B, int W,
                       int T, int V, int blank, const int64_t *scoring_ids, int S, int64_t *scoring_idmap, float *att_scores,
                       float one_minus_w, float w, float *r, int ldr, float *log_psi, float *token_scores, float *joint,
                       void *workspace, size_t workspace_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    ARG_CHECK(start >= 1 && end >= 1 && end <= T && (ol > 0 || start == 1), CTCPS_E_BADARG, "score: frame window outside [1, T]");
    ARG_CHECK(x_logp && blank_lp && r_prev && last_ids && r && log_psi, CTCPS_E_BADARG, "score: null pointer");
    ARG_CHECK(token_scores != nullptr || (S == 0 && ol <= T), CTCPS_E_BADARG, "score: token_scores may only be omitted on the full-vocabulary path");
    ARG_CHECK(B > 0 && W > 0 && T > 0 && V > 0 && ol >= 0 && S >= 0, CTCPS_E_BADARG, "score: non-positive size");
    ARG_CHECK(blank >= 0 && blank < V, CTCPS_E_BADARG, "score: blank id outside the vocabulary");
    ARG_CHECK(att_scores == nullptr || joint != nullptr, CTCPS_E_BADARG, "score: att_scores given without joint output");
    ARG_CHECK((S == 0) == (scoring_ids == nullptr), CTCPS_E_BADARG, "score: scoring_ids and S disagree");
    ARG_CHECK(S == 0 || scoring_idmap != nullptr, CTCPS_E_BADARG, "score: scoring_idmap output missing");
    const int snum = S > 0 ? S : V;
    ARG_CHECK(ldx >= V && (ldx & 3) == 0 && ldr >= snum && (ldr & 3) == 0, CTCPS_E_ALIGN, "score: ldx/ldr must be multiples of 4 and >= V/snum");
    ARG_CHECK((((uintptr_t)x_logp) & 15) == 0 && (((uintptr_t)r) & 15) == 0, CTCPS_E_ALIGN, "score: x_logp and r must be 16-byte aligned");
    const long long BW = (long long)B * W;
    ARG_CHECK(BW * (long long)(V > snum ? V : snum) < (1ll << 31) && (long long)B * T < (1ll << 31), CTCPS_E_TOOBIG,
              "score: BW*V or B*T exceeds 2^31");

    if (start > end) {  // ctc_scorer.py:138-145
        k_fill_f32<<<grid_for((size_t)T * 2 * BW * ldr, 256), 256, 0, st>>>(r, (size_t)T * 2 * BW * ldr, LZ);
        if (S > 0) {
            k_fill_i64<<<grid_for((size_t)BW * V, 256), 256, 0, st>>>(scoring_idmap, (size_t)BW * V, -1);
            k_build_idmap<<<(int)((BW + 127) / 128), 128, 0, st>>>(scoring_ids, (int)BW, S, V, scoring_idmap);
        }
        k_finalize<<<grid_for((size_t)BW * V, 256), 256, 0, st>>>(log_psi, s_prev, s_row_stride, s_col_stride, att_scores,
                                                                 one_minus_w, w, (int)BW, V, blank, token_scores, joint, 1);
        return cuda_rc(cudaGetLastError());
    }

    if (S > 0) {
        k_fill_i64<<<grid_for((size_t)BW * V, 256), 256, 0, st>>>(scoring_idmap, (size_t)BW * V, -1);
        k_build_idmap<<<(int)((BW + 127) / 128), 128, 0, st>>>(scoring_ids, (int)BW, S, V, scoring_idmap);
        k_fill_f32<<<grid_for((size_t)BW * V, 256), 256, 0, st>>>(log_psi, (size_t)BW * V, LZ);  // :156
        const long long lanes = BW * S;
        k_score_partial<<<(unsigned)((lanes + 127) / 128), 128, 0, st>>>(x_logp, ldx, blank_lp, r_prev, last_ids, scoring_ids,
                                                                        scoring_idmap, ol, start, end, B, W, T, V, S, r, ldr, log_psi);
        k_finalize<<<grid_for((size_t)BW * V, 256), 256, 0, st>>>(log_psi, s_prev, s_row_stride, s_col_stride, att_scores,
                                                                 one_minus_w, w, (int)BW, V, blank, token_scores, joint, 0);
        return cuda_rc(cudaGetLastError());
    }

    // full vocabulary
    int HW, G;
    pick_hw(W, &HW, &G);
    const Workspace ws = plan_workspace(B, T, W);
    ARG_CHECK(workspace != nullptr && workspace_bytes >= ws.total, CTCPS_E_WORKSPACE, "score: workspace too small");
    ARG_CHECK((((uintptr_t)workspace) & 255) == 0, CTCPS_E_ALIGN, "score: workspace must be 256-byte aligned");
    float4 *aux = reinterpret_cast<float4 *>((char *)workspace + ws.aux_off);
    float *Gmax = reinterpret_cast<float *>((char *)workspace + ws.g_off);
    const int Tpad = tpad_of(T);
    {
        const int warps = B * G * HW;
        k_prep<<<(warps + 3) / 4, 128, 0, st>>>(r_prev, blank_lp, B, W, T, HW, G, start, end, Tpad, aux, Gmax);
    }

    CUtensorMap tm;
    {
        const int rc_map = encode_x_map(&tm, x_logp, ldx, B, T, V);
        if (rc_map) return rc_map;
    }
    ScoreArgs a;
    a.aux = aux;
    a.Gmax = Gmax;
    a.s_prev = s_prev;
    a.s_rs = s_row_stride;
    a.s_cs = s_col_stride;
    a.last_ids = last_ids;
    a.att = att_scores;
    a.omw = one_minus_w;
    a.w = w;
    a.r = r;
    a.ldr = ldr;
    a.log_psi = log_psi;
    a.token_scores = token_scores;
    a.joint = joint;
    a.B = B;
    a.W = W;
    a.T = T;
    a.V = V;
    a.blank = blank;
    a.ol = ol;
    a.start = start;
    a.end = end;
    a.G = G;
    a.Tpad = Tpad;
    constexpr int NT = 128;
    a.nvt = (ldr + NT * 4 - 1) / (NT * 4);
    int rc;
    switch (HW) {
        case 1: rc = launch_score_full<1, NT, 4>(tm, a, st); break;
        case 2: rc = launch_score_full<2, NT, 4>(tm, a, st); break;
        case 3: rc = launch_score_full<3, NT, 4>(tm, a, st); break;
        case 4: rc = launch_score_full<4, NT, 4>(tm, a, st); break;
        default: rc = launch_score_full<5, NT, 4>(tm, a, st); break;
    }
    return rc;
}

// the call without attention weights (:133-136): frames max(ol, 1) .. T; ol > T takes the early return of :138-145
int ctcps_score(const float *x_logp, int ldx, const float *blank_lp, const float *r_prev, const float *s_prev,
                int64_t s_row_stride, int64_t s_col_stride, const int64_t *last_ids, int ol, int B, int W, int T, int V,
                int blank, const int64_t *scoring_ids, int S, int64_t *scoring_idmap, float *att_scores, float one_minus_w,
                float w, float *r, int ldr, float *log_psi, float *token_scores, float *joint, void *workspace,
                size_t workspace_bytes, void *stream) {
    return ctcps_score_window(x_logp, ldx, blank_lp, r_prev, s_prev, s_row_stride, s_col_stride, last_ids, ol, ol > 1 ? ol : 1, T, B, W,
                              T, V, blank, scoring_ids, S, scoring_idmap, att_scores, one_minus_w, w, r, ldr, log_psi, token_scores,
                              joint, workspace, workspace_bytes, stream);
}

// shared body of ctcps_score_lazy (dense outputs) and ctcps_score_lazy_topk (per-tile candidate lists)
static int score_lazy_impl(const float *x_logp, int ldx, const float *r_prev, const float *s_prev, int64_t s_row_stride,
                           int64_t s_col_stride, const int64_t *last_ids, int ol, int B, int W, int T, int V, int blank,
                           float *att_scores, float one_minus_w, float w, float *log_psi, float *token_scores, float *joint,
                           const PsiTopk *tk, const int64_t *xlens, void *workspace, size_t workspace_bytes, int workspace_prepared,
                           cudaStream_t st) {
    int HW, HWP, G;
    pick_hw_psi(W, &HW, &HWP, &G);
    const WorkspaceLazy ws = plan_workspace_lazy(B, T, W, V);
    ARG_CHECK(workspace != nullptr && workspace_bytes >= ws.total, CTCPS_E_WORKSPACE, "score_lazy: workspace too small");
    ARG_CHECK((((uintptr_t)workspace) & 255) == 0, CTCPS_E_ALIGN, "score_lazy: workspace must be 256-byte aligned");
    float *lin = reinterpret_cast<float *>((char *)workspace + ws.lin_off);
    float *Gmax = reinterpret_cast<float *>((char *)workspace + ws.g_off);
    float *psic = reinterpret_cast<float *>((char *)workspace + ws.c_off);
    int2 *frange = reinterpret_cast<int2 *>((char *)workspace + ws.r_off);
    const int Tpad = tpad_of(T);
    const int start = ol > 1 ? ol : 1;
    if (!workspace_prepared) {  // else ctcps_select_lazy already wrote lin / Gmax / psic for this very call
        const int warps = B * G * HWP;
        k_prep_psi<<<(warps + 3) / 4, 128, 0, st>>>(r_prev, XView{x_logp, (long long)T * ldx, (long long)ldx, 1}, last_ids, B, W, T, V, HW, HWP, G, start, Tpad, lin, Gmax, psic);
        k_lin_range<<<B * G, 128, 0, st>>>(lin, Tpad, HWP, frange);
    }
    CUtensorMap tm;
    int rc = encode_x_map(&tm, x_logp, ldx, B, T, V);
    if (rc) return rc;
    PsiArgs a;
    a.lin = lin;
    a.Gmax = Gmax;
    a.psic = psic;
    a.s_prev = s_prev;
    a.s_rs = s_row_stride;
    a.s_cs = s_col_stride;
    a.last_ids = last_ids;
    a.att = att_scores;
    a.omw = one_minus_w;
    a.w = w;
    a.log_psi = log_psi;
    a.token_scores = token_scores;
    a.joint = joint;
    a.B = B;
    a.W = W;
    a.T = T;
    a.V = V;
    a.blank = blank;
    a.ol = ol;
    a.G = G;
    a.Tpad = Tpad;
    // 256-token tiles (64 threads, 8 resident CTAs per SM) were measured in round 1 and not kept: no gain for the small
    // shapes (a tile is a serial chain of T/8 TMA chunks, halving its width halves nothing) and 4 % slower at C2; neither
    // was splitting the frame range of the last wave's tiles across CTAs (k_psi_split, removed in round 2: not faster on
    // any BASELINE shape, profiles/r1x_kernels_ncu.md section 3).
    a.nvt = (V + PSI_NT * 4 - 1) / (PSI_NT * 4);
    a.prefetch = psi_prefetch_chunks();
    a.xlens = xlens;
    a.frange = frame_window() ? frange : nullptr;
    {   // posteriors that cannot stay in L2 anyway (> 96 MB) are streamed with evict-first priority (CTCPS_PSI_EVICT_FIRST=0: off)
        static int ef = -1;
        if (ef < 0) {
            const char *ev = getenv("CTCPS_PSI_EVICT_FIRST");
            ef = (ev != nullptr && atoi(ev) == 0) ? 0 : 1;
        }
        a.evict_first = (ef && (size_t)B * T * V * sizeof(float) > ((size_t)96 << 20)) ? 1 : 0;
    }
    a.counter = g_stream_counter;
    a.tk.beam_scores = nullptr, a.tk.lists = nullptr, a.tk.K = 0, a.tk.done = nullptr;
    if (tk != nullptr) {
        a.tk = *tk;
        return dispatch_psi_full<true>(HW, tm, a, st);
    }
    return dispatch_psi_full<false>(HW, tm, a, st);
}

int ctcps_score_lazy(const float *x_logp, int ldx, const float *blank_lp, const float *r_prev, const float *s_prev,
                     int64_t s_row_stride, int64_t s_col_stride, const int64_t *last_ids, int ol, int B, int W, int T, int V,
                     int blank, float *att_scores, float one_minus_w, float w, float *log_psi, float *token_scores,
                     float *joint, void *workspace, size_t workspace_bytes, int workspace_prepared, void *stream) {
    return ctcps_score_lazy_lens(x_logp, ldx, blank_lp, nullptr, r_prev, s_prev, s_row_stride, s_col_stride, last_ids, ol, B, W, T, V, blank,
                                 att_scores, one_minus_w, w, log_psi, token_scores, joint, workspace, workspace_bytes, workspace_prepared, stream);
}

int ctcps_score_lazy_lens(const float *x_logp, int ldx, const float *blank_lp, const int64_t *xlens, const float *r_prev, const float *s_prev,
                          int64_t s_row_stride, int64_t s_col_stride, const int64_t *last_ids, int ol, int B, int W, int T, int V,
                          int blank, float *att_scores, float one_minus_w, float w, float *log_psi, float *token_scores,
                          float *joint, void *workspace, size_t workspace_bytes, int workspace_prepared, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    (void)blank_lp;
    ARG_CHECK(x_logp && r_prev && last_ids && log_psi, CTCPS_E_BADARG, "score_lazy: null pointer");
    ARG_CHECK(B > 0 && W > 0 && T > 0 && V > 0 && ol >= 0, CTCPS_E_BADARG, "score_lazy: non-positive size");
    ARG_CHECK(blank >= 0 && blank < V, CTCPS_E_BADARG, "score_lazy: blank id outside the vocabulary");
    ARG_CHECK(att_scores == nullptr || joint != nullptr, CTCPS_E_BADARG, "score_lazy: att_scores given without joint output");
    ARG_CHECK(ldx >= V && (ldx & 3) == 0 && (((uintptr_t)x_logp) & 15) == 0, CTCPS_E_ALIGN, "score_lazy: ldx must be a multiple of 4 and x_logp 16-byte aligned");
    const long long BW = (long long)B * W;
    ARG_CHECK(BW * (long long)V < (1ll << 31) && (long long)B * T < (1ll << 31), CTCPS_E_TOOBIG, "score_lazy: BW*V or B*T exceeds 2^31");
    const int start = ol > 1 ? ol : 1;
    if (start > T) {  // ctc_scorer.py:138-145
        k_finalize<<<grid_for((size_t)BW * V, 256), 256, 0, st>>>(log_psi, s_prev, s_row_stride, s_col_stride, att_scores,
                                                                 one_minus_w, w, (int)BW, V, blank, token_scores, joint, 1);
        return cuda_rc(cudaGetLastError());
    }
    return score_lazy_impl(x_logp, ldx, r_prev, s_prev, s_row_stride, s_col_stride, last_ids, ol, B, W, T, V, blank, att_scores,
                           one_minus_w, w, log_psi, token_scores, joint, nullptr, xlens, workspace, workspace_bytes, workspace_prepared, st);
}

int ctcps_topk_lists_shape(int B, int W, int V, int *lists_per_utterance, int *K) {
    ARG_CHECK(lists_per_utterance && K && B > 0 && W > 0 && V > 0, CTCPS_E_BADARG, "topk_lists_shape: bad argument");
    int HW, HWP, G;
    pick_hw_psi(W, &HW, &HWP, &G);
    *lists_per_utterance = ((V + PSI_NT * 4 - 1) / (PSI_NT * 4)) * G;
    *K = 2 * W;
    return 0;
}

int ctcps_score_lazy_topk(const float *x_logp, int ldx, const float *r_prev, const float *s_prev, const int64_t *last_ids, int ol,
                          int B, int W, int T, int V, int blank, const float *att_scores, float one_minus_w, float w,
                          const float *beam_scores, float *log_psi, float *tile_lists, void *workspace, size_t workspace_bytes,
                          int workspace_prepared, void *stream) {
    return ctcps_score_lazy_topk_active(x_logp, ldx, r_prev, s_prev, last_ids, ol, B, W, T, V, blank, att_scores, one_minus_w, w, beam_scores,
                                        nullptr, nullptr, log_psi, tile_lists, workspace, workspace_bytes, workspace_prepared, stream);
}

int ctcps_score_lazy_topk_active(const float *x_logp, int ldx, const float *r_prev, const float *s_prev, const int64_t *last_ids, int ol,
                                 int B, int W, int T, int V, int blank, const float *att_scores, float one_minus_w, float w,
                                 const float *beam_scores, const unsigned char *done, const int64_t *xlens, float *log_psi, float *tile_lists,
                                 void *workspace, size_t workspace_bytes, int workspace_prepared, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    ARG_CHECK(x_logp && r_prev && last_ids && att_scores && beam_scores && tile_lists && log_psi, CTCPS_E_BADARG, "score_lazy_topk: null pointer");
    ARG_CHECK(B > 0 && W > 0 && T > 0 && V > 0 && ol >= 0, CTCPS_E_BADARG, "score_lazy_topk: non-positive size");
    ARG_CHECK(blank >= 0 && blank < V, CTCPS_E_BADARG, "score_lazy_topk: blank id outside the vocabulary");
    ARG_CHECK((V & 3) == 0 && ldx >= V && (ldx & 3) == 0 &&
                  ((((uintptr_t)x_logp) | ((uintptr_t)att_scores) | ((uintptr_t)log_psi)) & 15) == 0 && (((uintptr_t)tile_lists) & 7) == 0,
              CTCPS_E_ALIGN, "score_lazy_topk: V and ldx must be multiples of 4, x_logp / att_scores / log_psi 16-byte aligned");
    ARG_CHECK(2 * W <= BEAM_MAXK, CTCPS_E_TOOBIG, "score_lazy_topk: num_beams > 32 is not supported");
    const long long BW = (long long)B * W;
    ARG_CHECK(BW * (long long)V < (1ll << 31) && (long long)B * T < (1ll << 31), CTCPS_E_TOOBIG, "score_lazy_topk: BW*V or B*T exceeds 2^31");
    ARG_CHECK(ol <= T, CTCPS_E_BADARG, "score_lazy_topk: prefix longer than the utterance (use ctcps_score_lazy: every score is logzero)");
    PsiTopk tk;
    tk.beam_scores = beam_scores;
    tk.lists = reinterpret_cast<float2 *>(tile_lists);
    tk.K = 2 * W;
    tk.done = done;
    return score_lazy_impl(x_logp, ldx, r_prev, s_prev, 1, 0, last_ids, ol, B, W, T, V, blank, const_cast<float *>(att_scores), one_minus_w, w,
                           log_psi, nullptr, nullptr, &tk, xlens, workspace, workspace_bytes, workspace_prepared, st);
}

int ctcps_select_lazy(const float *x_logp, int ldx, const float *blank_lp, const float *r_prev, const int64_t *last_ids, int ol,
                      const float *log_psi, const int64_t *best_ids, int B, int W, int T, int V, float *r_new, float *s_new,
                      void *next_workspace, size_t next_workspace_bytes, void *stream) {
    ARG_CHECK(x_logp && log_psi, CTCPS_E_BADARG, "select_lazy: null pointer");
    ARG_CHECK(T > 0 && V > 0 && ldx >= V, CTCPS_E_BADARG, "select_lazy: bad size");
    const XView x = {x_logp, (long long)T * ldx, (long long)ldx, 1};
    return select_lazy_impl(x, blank_lp, r_prev, last_ids, ol, log_psi, nullptr, 0, best_ids, B, W, T, V, r_new, s_new, next_workspace,
                            next_workspace_bytes, (cudaStream_t)stream);
}

int ctcps_beam_step_workspace_bytes(int B, int W, size_t *out_bytes) {
    ARG_CHECK(out_bytes != nullptr && B > 0 && W > 0, CTCPS_E_BADARG, "beam_step_workspace_bytes: bad argument");
    // partial candidate lists + one ticket per utterance + the global ticket (tickets must be zeroed once by the caller)
    // + the per-hypothesis top-2W lists of the two-kernel dense path
    *out_bytes = plan_beam_workspace(B, W).total;
    return 0;
}

int ctcps_beam_step(const float *joint, float *beam_scores, const int64_t *ids_cur, int64_t *ids_next, int64_t ld_ids, int L, int B,
                    int W, int V, int eos, int pad, float len_norm, float *pool_scores, int64_t *pool_lens, int64_t *pool_seqs,
                    int64_t ld_pool, unsigned char *done, void *workspace, size_t workspace_bytes, int64_t *done_ring, int ring,
                    int64_t step_tag, int64_t *best_ids_out, void *stream) {
    return beam_step_impl(joint, nullptr, 0, beam_scores, ids_cur, ids_next, ld_ids, L, B, W, V, eos, pad, len_norm, pool_scores, pool_lens,
                          pool_seqs, ld_pool, done, workspace, workspace_bytes, done_ring, ring, step_tag, best_ids_out, nullptr,
                          (cudaStream_t)stream);
}

int ctcps_beam_step_candidates(const float *cand_joint, const int64_t *cand_ids, int S, float *beam_scores, const int64_t *ids_cur,
                               int64_t *ids_next, int64_t ld_ids, int L, int B, int W, int V, int eos, int pad, float len_norm,
                               float *pool_scores, int64_t *pool_lens, int64_t *pool_seqs, int64_t ld_pool, unsigned char *done,
                               void *workspace, size_t workspace_bytes, int64_t *done_ring, int ring, int64_t step_tag,
                               int64_t *best_ids_out, void *stream) {
    ARG_CHECK(cand_ids != nullptr && S >= 2, CTCPS_E_BADARG, "beam_step_candidates: need candidate ids and S >= 2");
    return beam_step_impl(cand_joint, cand_ids, S, beam_scores, ids_cur, ids_next, ld_ids, L, B, W, V, eos, pad, len_norm, pool_scores,
                          pool_lens, pool_seqs, ld_pool, done, workspace, workspace_bytes, done_ring, ring, step_tag, best_ids_out, nullptr,
                          (cudaStream_t)stream);
}

int ctcps_beam_step_lists(const float *tile_lists, int lists_per_utterance, float *beam_scores, const int64_t *ids_cur, int64_t *ids_next,
                          int64_t ld_ids, int L, int B, int W, int V, int eos, int pad, float len_norm, float *pool_scores,
                          int64_t *pool_lens, int64_t *pool_seqs, int64_t ld_pool, unsigned char *done, void *workspace,
                          size_t workspace_bytes, int64_t *done_ring, int ring, int64_t step_tag, int64_t *best_ids_out,
                          int64_t *last_ids_out, void *stream) {
    ARG_CHECK(tile_lists && beam_scores && ids_cur && ids_next && pool_scores && pool_lens && pool_seqs && done && workspace,
              CTCPS_E_BADARG, "beam_step_lists: null pointer");
    ARG_CHECK(B > 0 && W > 0 && V > 0 && L >= 1 && L < ld_ids && L - 1 <= ld_pool && lists_per_utterance > 0, CTCPS_E_BADARG,
              "beam_step_lists: bad size");
    ARG_CHECK(2 * W <= BEAM_MAXK && W <= 32 && (long long)lists_per_utterance * 2 * W <= MERGE_MAX, CTCPS_E_TOOBIG,
              "beam_step_lists: num_beams > 32 or more than 2048 candidates per utterance");
    ARG_CHECK((((uintptr_t)tile_lists) & 7) == 0, CTCPS_E_ALIGN, "beam_step_lists: tile_lists must be 8-byte aligned");
    ARG_CHECK(done_ring == nullptr || ring > 0, CTCPS_E_BADARG, "beam_step_lists: done_ring without ring size");
    const BeamWorkspace bw = plan_beam_workspace(B, W);
    ARG_CHECK(workspace_bytes >= bw.total, CTCPS_E_WORKSPACE, "beam_step_lists: workspace too small");
    unsigned int *utt_ticket = reinterpret_cast<unsigned int *>((char *)workspace + bw.ticket_off);
    BeamOut o;
    o.beam_scores = beam_scores, o.best_ids_out = best_ids_out, o.last_ids_out = last_ids_out, o.ids_cur = ids_cur, o.ids_next = ids_next;
    o.ld_ids = ld_ids, o.pool_scores = pool_scores, o.pool_lens = pool_lens, o.pool_seqs = pool_seqs, o.ld_pool = ld_pool, o.done = done;
    o.ticket = utt_ticket + B, o.done_ring = (long long *)done_ring, o.ring = ring, o.step_tag = step_tag;
    k_beam_merge<<<B, BEAM_NT, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2 *>(tile_lists), lists_per_utterance, L, W, V, eos, pad,
                                                         len_norm, o);
    return cuda_rc(cudaGetLastError());
}

int ctcps_select(const float *r, int ldr, const float *log_psi, const int64_t *best_ids, const int64_t *scoring_idmap, int B,
                 int W, int T, int V, int S, float *r_new, float *s_new, void *stream) {
    ARG_CHECK(r && log_psi && best_ids && r_new && s_new, CTCPS_E_BADARG, "select: null pointer");
    ARG_CHECK(B > 0 && W > 0 && T > 0 && V > 0 && S >= 0, CTCPS_E_BADARG, "select: non-positive size");
    ARG_CHECK((S == 0) == (scoring_idmap == nullptr), CTCPS_E_BADARG, "select: scoring_idmap and S disagree");
    ARG_CHECK(ldr >= (S > 0 ? S : V), CTCPS_E_ALIGN, "select: ldr smaller than the lane count");
    const size_t n = (size_t)T * 2 * B * W;
    k_select<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(r, ldr, log_psi, best_ids, scoring_idmap, B, W, T, V, S, r_new,
                                                              s_new);
    return cuda_rc(cudaGetLastError());
}

int ctcps_eos_space_trick(const float *att_scores, const float *ctc_scores, float *next, int BW, int V, int eos, int space,
                          float k, void *stream) {
    ARG_CHECK(att_scores && ctc_scores && next && BW > 0 && V > 0, CTCPS_E_BADARG, "eos_space_trick: bad argument");
    if (eos < 0 || eos >= V || space < 0 || space >= V) return 0;  // argmax can never equal an id outside the vocabulary
    k_trick<<<BW, 256, 0, (cudaStream_t)stream>>>(att_scores, ctc_scores, next, V, eos, space, k);
    return cuda_rc(cudaGetLastError());
}

/* ---- pre-beam (candidate) decode step, SURVEY.md 8(f) N2 ------------------------------------------------------ */

int ctcps_padded_lt(int T) { return (T + 7) & ~7; }

int ctcps_transpose_vt(const float *x_logp, int ldx, int B, int T, int V, float *x_vt, int ldt, void *stream) {
    ARG_CHECK(x_logp && x_vt && B > 0 && T > 0 && V > 0, CTCPS_E_BADARG, "transpose_vt: bad argument");
    ARG_CHECK(ldx >= V && ldt >= T && (ldt & 3) == 0 && (((uintptr_t)x_vt) & 15) == 0, CTCPS_E_ALIGN,
              "transpose_vt: ldt must be a multiple of 4 and >= T, x_vt 16-byte aligned");
    ARG_CHECK(B < 65536 && (ldt + 31) / 32 < 65536, CTCPS_E_TOOBIG, "transpose_vt: B or T too large for one launch");
    const dim3 grid((V + 31) / 32, (ldt + 31) / 32, B);
    k_transpose_vt<<<grid, 256, 0, (cudaStream_t)stream>>>(x_logp, ldx, T, V, x_vt, ldt);
    return cuda_rc(cudaGetLastError());
}

int ctcps_prebeam_topk(float *att_scores, int BW, int V, int blank, int S, int64_t *scoring_ids, float *cand_att, void *stream) {
    ARG_CHECK(att_scores && scoring_ids && cand_att && BW > 0 && V > 0, CTCPS_E_BADARG, "prebeam_topk: bad argument");
    ARG_CHECK(blank >= 0 && blank < V, CTCPS_E_BADARG, "prebeam_topk: blank id outside the vocabulary");
    ARG_CHECK(S >= 1 && S <= 64 && S <= V, CTCPS_E_TOOBIG, "prebeam_topk: need 1 <= S <= min(64, V)");
    return launch_prebeam_topk(att_scores, BW, V, blank, S, scoring_ids, cand_att, (cudaStream_t)stream);
}

int ctcps_score_candidates(const float *x_vt, int ldt, const float *r_prev, const float *s_prev, const int64_t *last_ids, int ol,
                           int B, int W, int T, int V, int blank, const int64_t *scoring_ids, int S, const float *cand_att,
                           float one_minus_w, float w, float *cand_log_psi, float *cand_token_scores, float *cand_joint,
                           void *workspace, size_t workspace_bytes, int workspace_prepared, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    ARG_CHECK(x_vt && r_prev && last_ids && scoring_ids && cand_log_psi, CTCPS_E_BADARG, "score_candidates: null pointer");
    ARG_CHECK(B > 0 && W > 0 && T > 0 && V > 0 && ol >= 0 && S > 0, CTCPS_E_BADARG, "score_candidates: non-positive size");
    ARG_CHECK(blank >= 0 && blank < V, CTCPS_E_BADARG, "score_candidates: blank id outside the vocabulary");
    ARG_CHECK(cand_joint == nullptr || cand_att != nullptr, CTCPS_E_BADARG, "score_candidates: cand_joint needs cand_att");
    const int Tpad = tpad_of(T);
    ARG_CHECK(ldt >= T && ldt <= Tpad && (ldt & 3) == 0 && (((uintptr_t)x_vt) & 15) == 0, CTCPS_E_ALIGN,
              "score_candidates: ldt must be ctcps_padded_lt(T) and x_vt 16-byte aligned");
    const long long BW = (long long)B * W;
    ARG_CHECK(BW * (long long)S < (1ll << 31) && (long long)B * V < (1ll << 31), CTCPS_E_TOOBIG, "score_candidates: sizes exceed 2^31");
    const int start = ol > 1 ? ol : 1;
    if (start > T) {  // ctc_scorer.py:138-145
        const size_t n = (size_t)BW * S;
        k_cand_all_logzero<<<grid_for(n, 256), 256, 0, st>>>(cand_att, one_minus_w, w, n, cand_log_psi, cand_token_scores, cand_joint);
        return cuda_rc(cudaGetLastError());
    }
    int HW, HWP, G;
    pick_hw_psi(W, &HW, &HWP, &G);
    const WorkspaceLazy ws = plan_workspace_lazy(B, T, W, V);
    ARG_CHECK(workspace != nullptr && workspace_bytes >= ws.total, CTCPS_E_WORKSPACE, "score_candidates: workspace too small");
    ARG_CHECK((((uintptr_t)workspace) & 255) == 0, CTCPS_E_ALIGN, "score_candidates: workspace must be 256-byte aligned");
    float *lin = reinterpret_cast<float *>((char *)workspace + ws.lin_off);
    float *Gmax = reinterpret_cast<float *>((char *)workspace + ws.g_off);
    float *psic = reinterpret_cast<float *>((char *)workspace + ws.c_off);
    if (!workspace_prepared) {
        const int warps = B * G * HWP;
        k_prep_psi<<<(warps + 3) / 4, 128, 0, st>>>(r_prev, XView{x_vt, (long long)V * ldt, 1, (long long)ldt}, last_ids, B, W, T, V, HW, HWP,
                                                    G, start, Tpad, lin, Gmax, psic);
    }
    CandArgs a;
    a.xt = x_vt, a.ldt = ldt, a.lin = lin, a.Gmax = Gmax, a.psic = psic, a.s_prev = s_prev, a.last_ids = last_ids, a.ids = scoring_ids;
    a.cand_att = cand_att, a.omw = one_minus_w, a.w = w, a.cand_log_psi = cand_log_psi, a.cand_ts = cand_token_scores;
    a.cand_joint = cand_joint, a.B = B, a.W = W, a.T = T, a.V = V, a.S = S, a.blank = blank, a.ol = ol, a.G = G, a.HW = HW, a.HWP = HWP;
    a.Tpad = Tpad;
    const int hper = (HW + CAND_SPLIT - 1) / CAND_SPLIT;  // hypotheses (= warps) per CTA
    const size_t smem = (size_t)hper * ldt * sizeof(float);
    ARG_CHECK(smem <= 200 * 1024, CTCPS_E_TOOBIG, "score_candidates: T too large for the shared-memory stream");
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_psi_cand, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    k_psi_cand<<<(unsigned)(B * G * CAND_SPLIT), 32 * hper, smem, st>>>(a);
    return cuda_rc(cudaGetLastError());
}

int ctcps_candidates_to_dense(const float *att_scores, const float *s_prev, const int64_t *scoring_ids, const float *cand_log_psi,
                              const float *cand_token_scores, const float *cand_joint, int BW, int V, int S, float one_minus_w,
                              float w, int ol, int T, float *log_psi, float *token_scores, float *joint, void *stream) {
    ARG_CHECK(scoring_ids && cand_log_psi && BW > 0 && V > 0 && S > 0, CTCPS_E_BADARG, "candidates_to_dense: bad argument");
    ARG_CHECK(joint == nullptr || (att_scores && cand_joint), CTCPS_E_BADARG, "candidates_to_dense: joint needs att_scores and cand_joint");
    ARG_CHECK(token_scores == nullptr || cand_token_scores, CTCPS_E_BADARG, "candidates_to_dense: token_scores needs cand_token_scores");
    const int start = ol > 1 ? ol : 1;
    k_cand_to_dense<<<BW, 256, 0, (cudaStream_t)stream>>>(att_scores, s_prev, scoring_ids, cand_log_psi, cand_token_scores, cand_joint, V,
                                                         S, one_minus_w, w, start > T ? 1 : 0, log_psi, token_scores, joint);
    return cuda_rc(cudaGetLastError());
}

int ctcps_select_lazy_candidates(const float *x_vt, int ldt, const float *blank_lp, const float *r_prev, const int64_t *last_ids,
                                 int ol, const int64_t *scoring_ids, int S, const float *cand_log_psi, const int64_t *best_ids, int B,
                                 int W, int T, int V, float *r_new, float *s_new, void *next_workspace,
                                 size_t next_workspace_bytes, void *stream) {
    ARG_CHECK(x_vt && scoring_ids && cand_log_psi && S > 0, CTCPS_E_BADARG, "select_lazy_candidates: null pointer");
    ARG_CHECK(T > 0 && V > 0 && ldt >= T, CTCPS_E_BADARG, "select_lazy_candidates: bad size");
    const XView x = {x_vt, (long long)V * ldt, 1, (long long)ldt};
    return select_lazy_impl(x, blank_lp, r_prev, last_ids, ol, cand_log_psi, scoring_ids, S, best_ids, B, W, T, V, r_new, s_new,
                            next_workspace, next_workspace_bytes, (cudaStream_t)stream);
}

/* ---- native decode-step driver ----------------------------------------------------------------------------- */

int ctcps_async_create(void **side_stream, void **ev_step, void **ev_select) {
    ARG_CHECK(side_stream && ev_step && ev_select, CTCPS_E_BADARG, "async_create: null pointer");
    cudaStream_t st;
    cudaEvent_t a, b;
    cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    if (e != cudaSuccess) return (int)e;
    e = cudaEventCreateWithFlags(&a, cudaEventDisableTiming);
    if (e != cudaSuccess) return (int)e;
    e = cudaEventCreateWithFlags(&b, cudaEventDisableTiming);
    if (e != cudaSuccess) return (int)e;
    *side_stream = st, *ev_step = a, *ev_select = b;
    return 0;
}

int ctcps_async_destroy(void *side_stream, void *ev_step, void *ev_select) {
    if (ev_step) cudaEventDestroy((cudaEvent_t)ev_step);
    if (ev_select) cudaEventDestroy((cudaEvent_t)ev_select);
    if (side_stream) cudaStreamDestroy((cudaStream_t)side_stream);
    return 0;
}

int ctcps_event_create(void **event) {
    ARG_CHECK(event != nullptr, CTCPS_E_BADARG, "event_create: null pointer");
    cudaEvent_t e;
    cudaError_t rc = cudaEventCreate(&e);
    if (rc != cudaSuccess) return (int)rc;
    *event = e;
    return 0;
}
int ctcps_event_destroy(void *event) { return event ? cuda_rc(cudaEventDestroy((cudaEvent_t)event)) : 0; }
int ctcps_event_elapsed_ms(void *begin, void *end, float *ms) {
    ARG_CHECK(begin && end && ms, CTCPS_E_BADARG, "event_elapsed_ms: null pointer");
    return cuda_rc(cudaEventElapsedTime(ms, (cudaEvent_t)begin, (cudaEvent_t)end));
}

int ctcps_decode_step(const ctcps_decode_session *s, float *att_scores, int step, void *ev_score_begin, void *ev_score_end,
                      void *stream) {
    ARG_CHECK(s != nullptr && att_scores != nullptr && step >= 0, CTCPS_E_BADARG, "decode_step: bad argument");
    ARG_CHECK(s->S == 0 || s->S >= 2, CTCPS_E_BADARG, "decode_step: S must be 0 (full vocabulary) or >= 2");
    cudaStream_t main_st = (cudaStream_t)stream;
    cudaStream_t side = s->side_stream ? (cudaStream_t)s->side_stream : main_st;
    const int cur = step & 1, nxt = cur ^ 1;
    const int L = step + 1, ol = step;
    const int B = s->B, W = s->W, T = s->T, V = s->V, S = s->S;
    const float *r_prev = step == 0 ? s->r0 : s->r_sel[cur];
    const float *s_prev = step == 0 ? nullptr : s->s_sel[cur];
    const int prepared = (step > 0 && step <= T) ? 1 : 0;  // the select of the previous step prepared this step's workspace
    const float len_norm = powf((float)L, s->length_penalty);
    int rc;
    if (S > 0) {  // the candidates do not depend on the CTC state: rank them while the select of the previous step still runs
        rc = ctcps_prebeam_topk(att_scores, B * W, V, s->blank, S, s->cand_ids[cur], s->cand_att[cur], main_st);
        if (rc) return rc;
    }
    if (step > 0 && side != main_st) {
        cudaError_t e = cudaStreamWaitEvent(main_st, (cudaEvent_t)s->ev_select, 0);
        if (e != cudaSuccess) return (int)e;
    }
    if (ev_score_begin) cudaEventRecord((cudaEvent_t)ev_score_begin, main_st);
    // Fused scoring + per-tile top-2W (full vocabulary): no (BW,V) tensor is written; the beam step merges the tile lists and
    // hands the prefix scores of the new rows straight to the state selection.
    int nlists = 0, Klist = 0;
    if (S == 0) ctcps_topk_lists_shape(B, W, V, &nlists, &Klist);
    const bool fused = S == 0 && s->tile_lists != nullptr && (V & 3) == 0 && ol <= T && 2 * W <= BEAM_MAXK && W <= 32 &&
                       (long long)nlists * Klist <= MERGE_MAX && s->log_psi[0] != nullptr && s->log_psi[1] != nullptr;
    const int64_t tag = s->tag_base + (int64_t)step;
    if (fused) {
        rc = ctcps_score_lazy_topk_active(s->x_logp, s->ldx, r_prev, s_prev, s->last_ids[cur], ol, B, W, T, V, s->blank, att_scores,
                                          s->one_minus_w, s->w, s->beam_scores, skip_done() ? s->done : nullptr, s->xlens, s->log_psi[cur],
                                          s->tile_lists, s->score_ws, s->score_ws_bytes, prepared, main_st);
        if (rc) return rc;
        if (ev_score_end) cudaEventRecord((cudaEvent_t)ev_score_end, main_st);
        rc = ctcps_beam_step_lists(s->tile_lists, nlists, s->beam_scores, s->ids[cur], s->ids[nxt], s->ld_ids, L, B, W, V, s->eos, s->pad,
                                   len_norm, s->pool_scores, s->pool_lens, s->pool_seqs, s->ld_pool, s->done, s->beam_ws, s->beam_ws_bytes,
                                   s->done_ring, s->ring, tag, s->best_ids, s->last_ids[nxt], main_st);
        if (rc) return rc;
        if (side != main_st) {
            cudaError_t e = cudaEventRecord((cudaEvent_t)s->ev_step, main_st);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(side, (cudaEvent_t)s->ev_step, 0);
            if (e != cudaSuccess) return (int)e;
        }
        const int64_t *sel = s->use_beam_idx ? s->best_ids : s->last_ids[nxt];
        rc = ctcps_select_lazy(s->x_logp, s->ldx, s->blank_lp, r_prev, s->last_ids[cur], ol, s->log_psi[cur], sel, B, W, T, V, s->r_sel[nxt],
                               s->s_sel[nxt], s->score_ws, s->score_ws_bytes, side);
        if (rc) return rc;
        if (side != main_st) {
            cudaError_t e = cudaEventRecord((cudaEvent_t)s->ev_select, side);
            if (e != cudaSuccess) return (int)e;
        }
        return 0;
    }
    if (S > 0) {
        rc = ctcps_score_candidates(s->x_vt, s->ldt, r_prev, s_prev, s->last_ids[cur], ol, B, W, T, V, s->blank, s->cand_ids[cur], S,
                                    s->cand_att[cur], s->one_minus_w, s->w, s->cand_log_psi[cur], nullptr, s->cand_joint, s->score_ws,
                                    s->score_ws_bytes, prepared, main_st);
    } else {
        ARG_CHECK(s->joint != nullptr && s->log_psi[0] != nullptr && s->log_psi[1] != nullptr, CTCPS_E_BADARG,
                  "decode_step: this step needs the dense (BW,V) buffers (joint, log_psi[2]) of the session");
        rc = ctcps_score_lazy_lens(s->x_logp, s->ldx, s->blank_lp, s->xlens, r_prev, s_prev, 1, 0, s->last_ids[cur], ol, B, W, T, V, s->blank,
                                   att_scores, s->one_minus_w, s->w, s->log_psi[cur], nullptr, s->joint, s->score_ws, s->score_ws_bytes, prepared,
                                   main_st);
    }
    if (rc) return rc;
    if (ev_score_end) cudaEventRecord((cudaEvent_t)ev_score_end, main_st);
    rc = beam_step_impl(S > 0 ? s->cand_joint : s->joint, S > 0 ? s->cand_ids[cur] : nullptr, S, s->beam_scores, s->ids[cur], s->ids[nxt],
                        s->ld_ids, L, B, W, V, s->eos, s->pad, len_norm, s->pool_scores, s->pool_lens, s->pool_seqs, s->ld_pool, s->done,
                        s->beam_ws, s->beam_ws_bytes, s->done_ring, s->ring, tag, s->best_ids, s->last_ids[nxt], main_st);
    if (rc) return rc;
    if (side != main_st) {
        cudaError_t e = cudaEventRecord((cudaEvent_t)s->ev_step, main_st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(side, (cudaEvent_t)s->ev_step, 0);
        if (e != cudaSuccess) return (int)e;
    }
    // state of the next step: ESPnet ids (source hypothesis * V + token) or, like the reference's processor, tokens only
    const int64_t *sel_ids = s->use_beam_idx ? s->best_ids : s->last_ids[nxt];
    if (S > 0) {
        rc = ctcps_select_lazy_candidates(s->x_vt, s->ldt, s->blank_lp, r_prev, s->last_ids[cur], ol, s->cand_ids[cur], S,
                                          s->cand_log_psi[cur], sel_ids, B, W, T, V, s->r_sel[nxt], s->s_sel[nxt], s->score_ws,
                                          s->score_ws_bytes, side);
    } else {
        rc = ctcps_select_lazy(s->x_logp, s->ldx, s->blank_lp, r_prev, s->last_ids[cur], ol, s->log_psi[cur], sel_ids, B, W, T, V,
                               s->r_sel[nxt], s->s_sel[nxt], s->score_ws, s->score_ws_bytes, side);
    }
    if (rc) return rc;
    if (side != main_st) {
        cudaError_t e = cudaEventRecord((cudaEvent_t)s->ev_select, side);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

int ctcps_decode_finish(const ctcps_decode_session *s, void *stream) {
    ARG_CHECK(s != nullptr, CTCPS_E_BADARG, "decode_finish: null session");
    if (s->side_stream == nullptr || s->side_stream == stream) return 0;
    return cuda_rc(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)s->ev_select, 0));
}

int ctcps_split_tf32(const float *x, int64_t n, int d, int weight_order, float *out, void *stream) {
    ARG_CHECK(x && out && n > 0 && d > 0, CTCPS_E_BADARG, "split_tf32: bad argument");
    ARG_CHECK((d & 3) == 0 && ((((uintptr_t)x) | ((uintptr_t)out)) & 15) == 0, CTCPS_E_ALIGN, "split_tf32: d must be a multiple of 4, pointers 16-byte aligned");
    k_split_tf32<<<grid_for((size_t)n * (d >> 2), 256), 256, 0, (cudaStream_t)stream>>>(x, n, d, weight_order, out);
    return cuda_rc(cudaGetLastError());
}

size_t ctcps_decode_session_size(void) { return sizeof(ctcps_decode_session); }

}  // extern "C"
