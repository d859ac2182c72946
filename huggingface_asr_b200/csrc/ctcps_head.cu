// ctcps_head.cu -- N4: the CTC head in front of the scorer as a hand-written tcgen05 kernel.
//
// Replaces, for BUTSpeechFIT/huggingface_asr, the step before the path:
//   logits = Wav2Vec2ForCTC.lm_head(hidden)                 src/reguler/e_branchformer.py:245-252   (fp32 Linear: hidden W^T + b)
//   x      = F.log_softmax(logits, dim=-1) + length padding src/decoding/ctc_scorer.py:279, :39-46
//
// k_head_gemm: logits tile by tile on the 5th-generation tensor cores at fp32 accuracy (three half-precision products, below),
// bias and the row-wise softmax statistics (running max, sum of exponentials) fused into the TMEM -> register epilogue; the raw
// logits go straight into the scorer's padded (B,T,ldx) buffer, the per-row statistics into a small side array.
// k_head_finish: one streaming pass turns the buffer into log-posteriors in place -- (z - max) - log(sum), the order
// torch.log_softmax uses -- applies the length padding and extracts the blank column (what k_init does after a library GEMM,
// without its two block-wide reductions).  Fusing this pass into the GEMM kernel was tried three ways and measured (C2, r2u):
// the grid works through the tiles in bands of 148 (~19 MB of logits, L2-resident), reports finished tiles to per-band counters
// in global memory (cooperative launch) and normalises the rows a band completed out of L2 -- by the drain warps after a grid
// barrier (2.6 ms: lock step), by the drain warps one band behind (2.26 ms) or by six warps of their own (2.33 ms, and the
// rows fall out of L2 before their turn).  HBM traffic did drop to one write of the posteriors, but the SM's load / store path
// is what the drain is short of, so the two-kernel form (1.90 ms) stayed.
//
// 3xFP16.  An fp16 significand has 11 bits -- exactly TF32's -- in 2 bytes instead of 4, and kind::f16 runs at twice the rate
// of kind::tf32.  What fp16 lacks is range, so every row of h (and W as a whole) is first scaled by a power of two that puts
// its largest magnitude into [2^13, 2^14) (exact; undone in the epilogue), then split
//   h s = H1 + 2^-11 H2,   H1 = fp16(h s),  H2 = fp16((h s - H1) 2^11)        (22 significand bits; the remainder is exact in fp32)
//   h W^T ~= [ H1 W1^T + 2^-11 (H2 W1^T + H1 W2^T) ] / (s_h s_W)               (the dropped H2 W2^T is 2^-22 relative)
// The first version of this kernel did the same with TF32 operands (3xTF32, 8 bytes per element and half the MMA rate): ncu
// showed it bound by the L2 -> SM operand stream (1.5 MB per 128 x 256 tile, 8.3 TB/s against a ~12 TB/s fabric limit).
// The tensor core accumulates in fp32 but rounds toward zero at every k-step, a bias proportional to the running sum
// (round 1 measured it through cuBLAS: 7.6e-5 when the large term shares an accumulator with 192 k-steps, 1.8e-5 with 64).
// So the two small cross terms accumulate in their OWN TMEM accumulator and the large term in another (32 k-steps at
// d = 512); the epilogue adds the two in fp32 registers.
//
// Structure (one CTA per SM, persistent, 640 threads, 96 registers each, all 227 KB of shared memory):
//   warp 0      TMA producer: per k-block of 32 halves the tiles H1, H2 (128 x 32) and W1, W2 (256 x 32), each ONE
//               contiguous bulk copy of a pre-swizzled image (k_split_blocked) -- 48 KB per stage, 4 stages -- behind
//               full / empty mbarriers
//   warp 1      TMEM allocation (512 columns: two 128 x 256 fp32 accumulators) and, one elected lane, the MMA issue:
//               per k-block 2 x (UMMA 128x256x16, kind::f16) x 3 products; tcgen05.commit frees the stage / publishes the tile
//   warps 4-19  drain: (TMEM lane quarter, column quarter) per warp; phase 1 moves the thread's 64 logits into registers
//               (tcgen05.ld 32 lanes x 16 columns, (big + 2^-11 small) / scale + bias) and hands the accumulators back; phase 2,
//               under the MMAs of the next tile: online max / sum-exp, rows through a swizzled shared-memory box, one TMA
//               store (cp.async.bulk.tensor, SASS UTMASTG) per 32 x 16 box
//   (warps 2-3 idle: the drain starts on a warpgroup boundary so that warp % 4 is the TMEM lane quarter)
// Work item = (128-row tile, quarter of the vocabulary tiles): 4 x 746 items at C2 keep the last wave short; the partial
// statistics of a row's quarters are combined by k_head_finish.
// Measured at C2 (B*T = 95 488 rows, d = 512, V = 5000; profiles/r2_head_ncu.md): split 0.06 ms + GEMM 1.19 ms (1.47 PFLOP of
// fp16 tensor work: 1.23 PFLOP/s = 0.78 of the measured bf16 burst peak, tensor pipe active 67 % of the time, the rest is
// phase 1) + finish 0.63 ms (3.78 GB at 6.0 TB/s) = 1.90 ms; 3xTF32 version 3.44 ms; split + cuBLAS + K-a 2.42 ms.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "ctcps.h"

namespace {

constexpr float LZ = CTCPS_LOGZERO;
constexpr int BLOCK_M = 128, BLOCK_N = 256, BLOCK_K = 32;  // BLOCK_K halves = 64 bytes = one SWIZZLE_64B row
constexpr int UMMA_K = 16;                                 // kind::f16: 32 bytes of K per instruction
constexpr float LO_SCALE = 2048.f, LO_UNSCALE = 1.f / 2048.f;  // the low parts are stored times 2^11 (fp16 range)
constexpr int NSTAGE = 4;                                  // 4 x 48 KB: the first version (2 x 96 KB, 128-byte rows) starved the MMAs
constexpr float LOG2E_F = 1.4426950408889634f;
constexpr int N_DRAIN = 16;                                // drain warps: (TMEM lane quarter, column quarter)
constexpr int HEAD_NT = (4 + N_DRAIN) * 32;                // warp 0 TMA, warp 1 MMA, warps 2-3 idle (the drain starts on a warpgroup boundary), 4..19 drain
constexpr int NCHUNK = 4;                                  // vocabulary quarters per row tile
constexpr int NPART = NCHUNK * 4;                          // partial softmax statistics per row: (vocabulary quarter, column quarter of the tiles)
constexpr uint32_t A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;   // 8 KB
constexpr uint32_t B_TILE_BYTES = BLOCK_N * BLOCK_K * 2;   // 16 KB
constexpr uint32_t STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
constexpr uint32_t TMEM_COLS = 512;
constexpr int GCOLS = 16;                                  // columns per TMEM load of a drain warp
constexpr int STORE_COLS = GCOLS;                          // columns per TMA store of a drain warp: a 32-row x 64-byte box

struct HeadSmem {
    alignas(1024) unsigned char a_hi[NSTAGE][A_TILE_BYTES];
    alignas(1024) unsigned char a_lo[NSTAGE][A_TILE_BYTES];
    alignas(1024) unsigned char b_hi[NSTAGE][B_TILE_BYTES];
    alignas(1024) unsigned char b_lo[NSTAGE][B_TILE_BYTES];
    alignas(1024) float stage_out[N_DRAIN][32][STORE_COLS];  // per drain warp: the 32 x 16 box its next TMA store reads (SWIZZLE_64B image)
    alignas(16) float bias[2][BLOCK_N];                             // bias of the current / next vocabulary tile
    alignas(8) uint64_t full[NSTAGE];
    alignas(8) uint64_t empty[NSTAGE];
    alignas(8) uint64_t tmem_full;
    alignas(8) uint64_t tmem_empty;
    uint32_t tmem_base;
};

static_assert(sizeof(HeadSmem) <= 227 * 1024, "HeadSmem must fit the 227 KB a CTA can have");

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
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
// TMA bulk copy of one contiguous, pre-swizzled operand tile (SASS UBLKCP)
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .b32 rx;\n"
        ".reg .pred px;\n"
        "elect.sync rx|px, 0xffffffff;\n"
        "selp.b32 %0, 1, 0, px;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// tcgen05.commit: the mbarrier gets one arrival when every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, both operands K-major, kind::f16 (fp16 x fp16 -> fp32), issued by one thread for the CTA
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// shared-memory matrix descriptor of a K-major tile whose rows are 64 bytes, SWIZZLE_64B (what TMA wrote): 8-row groups are
// 512 bytes apart (stride byte offset), the leading byte offset is unused for swizzled K-major layouts, descriptor version 1
// (sm_100), layout type 4 = SWIZZLE_64B.  A K-step of 32 bytes inside the swizzle row advances the start address.
__device__ __forceinline__ uint64_t smem_desc_sw64(const void *tile, uint32_t byte_offset) {
    const uint32_t addr = smem_u32(tile) + byte_offset;
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);          // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                         // leading byte offset (ignored), bits [16,30)
    d |= (uint64_t)(512 >> 4) << 32;                // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                         // version, bits [46,48)
    d |= (uint64_t)4 << 61;                         // layout type SWIZZLE_64B, bits [61,64)
    return d;
}
// instruction descriptor of kind::f16 (cute::UMMA::InstrDescriptor): D = F32 (bits 4-5 = 1), A = B = F16 (bits 7-9, 10-12 = 0),
// both K-major, dense, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// power-of-two scale that puts a largest magnitude m into [2^13, 2^14) (fp16 overflows at 2^16), and its inverse; both normal
// fp32 numbers for every m (0, subnormal, huge, inf / nan included)
__device__ __forceinline__ void scale_from_absmax(float m, float &s, float &inv) {
    const int e = (int)((__float_as_uint(m) >> 23) & 0xffu);
    int k = 13 - (e - 127);
    k = k > 100 ? 100 : (k < -120 ? -120 : k);
    s = __uint_as_float((uint32_t)(127 + k) << 23);
    inv = __uint_as_float((uint32_t)(127 - k) << 23);
}
// 32 TMEM lanes (this warp's quarter) x 32 consecutive columns -> 32 registers per thread (thread = lane = tile row);
// the loads of both accumulators are issued before the one wait
__device__ __forceinline__ void tmem_ld_32x32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// 32 TMEM lanes x 16 consecutive columns -> 16 registers per thread
__device__ __forceinline__ void tmem_ld_32x16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// shared memory by its own address space (through generic pointers the compiler emitted LD.E / ST.E with global-memory
// scoreboards for the epilogue's staging buffers)
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// TMA store of one box, shared -> global through a tensor map (columns / rows beyond the tensor are clipped); SASS UTMASTG
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tm, uint32_t smem_addr, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(smem_addr), "r"(c0), "r"(c1)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// the bulk stores this thread committed have finished READING shared memory (the buffer may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct HeadArgs {
    const unsigned char *a_hi, *a_lo;  // blocked, pre-swizzled operand tiles: [row tile][k-block][128 rows][64 bytes]
    const unsigned char *b_hi, *b_lo;  //                                      [vocabulary tile][k-block][256 rows][64 bytes]
    const float *a_inv;                // (padded n) 1 / scale of every row of h
    const uint32_t *w_absmax;          // bits of max |W| (the weight's scale is derived from it)
    const float *bias;                 // (V) or null
    float *z;                          // (n, ldz) raw logits out (the scorer's padded posterior buffer)
    float2 *stats;                     // (n, NPART): running max and sum of exp(z - max) over the columns of a (vocabulary quarter, column quarter)
    int n, d_pad, V, ldz;
    int n_mtiles, n_ntiles, tiles_per_chunk;
};

__global__ void __launch_bounds__(HEAD_NT, 1) k_head_gemm(const __grid_constant__ CUtensorMap tmz, const HeadArgs a) {
    // all 227 KB: there is no room to round the base up, the declared alignment has to hold (swizzled tiles repeat every 512 bytes)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    if (smem_u32(smem_raw) & 1023u) __trap();
    HeadSmem &sm = *reinterpret_cast<HeadSmem *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_kblocks = a.d_pad / BLOCK_K;
    const int n_items = a.n_mtiles * NCHUNK;

    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < NSTAGE; ++s) mbar_init(&sm.full[s], 1), mbar_init(&sm.empty[s], 1);
            mbar_init(&sm.tmem_full, 1);
            mbar_init(&sm.tmem_empty, N_DRAIN);  // one arrival per drain warp
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        // the whole warp allocates all 512 TMEM columns (one CTA per SM) and lets go of the permit
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            uint32_t it = 0;  // k-blocks issued so far (stage = it % NSTAGE)
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int mt = item / NCHUNK, ch = item % NCHUNK;
                const int nt0 = ch * a.tiles_per_chunk, nt1 = min(a.n_ntiles, nt0 + a.tiles_per_chunk);
                for (int nt = nt0; nt < nt1; ++nt)
                    for (int kb = 0; kb < n_kblocks; ++kb, ++it) {
                        const int s = it % NSTAGE;
                        mbar_wait(&sm.empty[s], ((it / NSTAGE) & 1) ^ 1);  // first pass over the ring: passes at once
                        mbar_expect_tx(&sm.full[s], STAGE_BYTES);
                        const size_t ao = ((size_t)mt * n_kblocks + kb) * A_TILE_BYTES, bo = ((size_t)nt * n_kblocks + kb) * B_TILE_BYTES;
                        bulk_load_1d(sm.a_hi[s], a.a_hi + ao, A_TILE_BYTES, &sm.full[s]);
                        bulk_load_1d(sm.a_lo[s], a.a_lo + ao, A_TILE_BYTES, &sm.full[s]);
                        bulk_load_1d(sm.b_hi[s], a.b_hi + bo, B_TILE_BYTES, &sm.full[s]);
                        bulk_load_1d(sm.b_lo[s], a.b_lo + bo, B_TILE_BYTES, &sm.full[s]);
                    }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N);
        const uint32_t d_big = tmem_base, d_small = tmem_base + BLOCK_N;
        uint32_t it = 0, tile = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int ch = item % NCHUNK;
            const int nt0 = ch * a.tiles_per_chunk, nt1 = min(a.n_ntiles, nt0 + a.tiles_per_chunk);
            for (int nt = nt0; nt < nt1; ++nt, ++tile) {
                mbar_wait(&sm.tmem_empty, (tile & 1) ^ 1);  // the drain warps have emptied the accumulators of the previous tile
                tc_fence_after();
                for (int kb = 0; kb < n_kblocks; ++kb, ++it) {
                    const int s = it % NSTAGE;
                    mbar_wait(&sm.full[s], (it / NSTAGE) & 1);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                            const uint32_t off = k * UMMA_K * 2;
                            const uint64_t ah = smem_desc_sw64(sm.a_hi[s], off), al = smem_desc_sw64(sm.a_lo[s], off);
                            const uint64_t bh = smem_desc_sw64(sm.b_hi[s], off), bl = smem_desc_sw64(sm.b_lo[s], off);
                            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
                            umma_f16(d_small, al, bh, idesc, acc);  // H2 W1^T
                            umma_f16(d_small, ah, bl, idesc, 1u);   // H1 W2^T
                            umma_f16(d_big, ah, bh, idesc, acc);    // H1 W1^T
                        }
                    }
                    __syncwarp();
                    if (elect_one()) {
                        umma_commit(&sm.empty[s]);                          // the stage is free once these MMAs have read it
                        if (kb == n_kblocks - 1) umma_commit(&sm.tmem_full);  // and the tile is complete
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 4) {
        // ===== drain (warps 4..19): TMEM lane quarter = warp % 4 (a hardware rule), column quarter = (warp - 4) / 4; a thread owns
        // one row of the tile and 64 of its 256 columns, 16 at a time.  History of this part, all measured with ncu at C2
        // (profiles/r2*_head*.md): the MMAs of a 128 x 256 tile take ~4 us, so the tile time IS the drain.
        //   v1  8 warps, wait after every TMEM load                                  MMA warp idle 43 % of the time
        //   v2  + loads of the next group in flight                                  14 us per tile: every float4 of the bias a
        //       global load in front of its first use (32 dependent L2 round trips per tile), float4 stores to 32 different
        //       rows per instruction (half-written sectors, twice the bytes towards L2)
        //   v3  + bias staged in shared memory while the MMAs run, stores through a per-warp shared-memory transposition
        //       (8 rows x 64 contiguous bytes per instruction)                       8.6 us
        //   v4  + 16-column groups (no spills), shared memory addressed as such (the generic loads carried global-memory
        //       scoreboards), columns beyond the vocabulary masked by a -inf bias instead of a compare per element   6.8 us
        //   v5  16 warps x 64 columns: four warps per scheduler hide the chain TMEM load -> max -> exp -> transposition -> store
        const int quarter = warp & 3, cq = (warp - 4) >> 2;
        const int row_in_tile = quarter * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cq * (BLOCK_N / 4));
        constexpr int NGROUP = BLOCK_N / 4 / GCOLS;  // 4 groups of 16 columns per thread and tile
        const uint32_t st_base = smem_u32(sm.stage_out[warp - 4]);
        // a lane = a row of the box: its four 16-byte chunks land where TMA's SWIZZLE_64B expects them (chunk ^ ((row >> 1) & 3)),
        // which also keeps the row-wise writes free of bank conflicts
        const uint32_t st_wr = st_base + (uint32_t)lane * (STORE_COLS * 4), st_wx = (uint32_t)((lane >> 1) & 3);
        float w_s, w_inv;
        scale_from_absmax(__uint_as_float(__ldg(a.w_absmax)), w_s, w_inv);
        uint32_t tile = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int mt = item / NCHUNK, ch = item % NCHUNK;
            const int nt0 = ch * a.tiles_per_chunk, nt1 = min(a.n_ntiles, nt0 + a.tiles_per_chunk);
            const long long row = (long long)mt * BLOCK_M + row_in_tile;
            const bool row_ok = row < a.n;
            const int box_row0 = mt * BLOCK_M + quarter * 32;
            float m_run = -INFINITY, s_run = 0.f;
            const float unscale = __ldg(a.a_inv + row) * w_inv;  // a_inv covers the padded rows
            for (int nt = nt0; nt < nt1; ++nt, ++tile) {
                {   // this tile's bias while its MMAs are still running: 0 without a bias, -inf beyond the vocabulary, which
                    // takes those columns out of the row maximum and the sum of exponentials without a compare per element
                    const int et = (int)threadIdx.x - 128;
                    if (et < BLOCK_N) {
                        const int v = nt * BLOCK_N + et;
                        sm.bias[tile & 1][et] = v < a.V ? (a.bias != nullptr ? __ldg(a.bias + v) : 0.f) : -INFINITY;
                    }
                    asm volatile("bar.sync 1, 512;" ::: "memory");  // the 16 drain warps; two buffers, so one barrier per tile is enough
                }
                const uint32_t bias_s = smem_u32(&sm.bias[tile & 1][cq * (BLOCK_N / 4)]);
                const int vbase = nt * BLOCK_N + cq * (BLOCK_N / 4);
                mbar_wait(&sm.tmem_full, tile & 1);
                tc_fence_after();
                // Phase 1: the thread's 64 logits of the tile into registers (accumulators combined, scale, bias), the accumulators back
                // to the MMA warp.  Phase 2, while the MMAs of the next tile run: row statistics and the stores.  (Timing probes at C2,
                // r2v: of 6.7 us of drain per tile the stores are 3.4 us -- 128 KB through the SM's ~32 B/clk path to L2 -- the TMEM
                // loads 1.7 us, the arithmetic 0.8 us; as long as the stores held the accumulators the MMAs waited for them.)
                float zv[NGROUP * GCOLS];
                {
                    uint32_t big[GCOLS], small[GCOLS];
                    tmem_ld_32x16_issue(lane_base, big);
                    tmem_ld_32x16_issue(lane_base + BLOCK_N, small);
#pragma unroll
                    for (int g = 0; g < NGROUP; ++g) {
                        tmem_ld_wait();  // group g has landed
#pragma unroll
                        for (int q = 0; q < GCOLS / 4; ++q) {
                            const float4 b4 = lds_v4(bias_s + (uint32_t)(g * GCOLS + q * 4) * 4);
                            float *zq = &zv[g * GCOLS + q * 4];
                            zq[0] = fmaf(fmaf(__uint_as_float(small[q * 4 + 0]), LO_UNSCALE, __uint_as_float(big[q * 4 + 0])), unscale, b4.x);
                            zq[1] = fmaf(fmaf(__uint_as_float(small[q * 4 + 1]), LO_UNSCALE, __uint_as_float(big[q * 4 + 1])), unscale, b4.y);
                            zq[2] = fmaf(fmaf(__uint_as_float(small[q * 4 + 2]), LO_UNSCALE, __uint_as_float(big[q * 4 + 2])), unscale, b4.z);
                            zq[3] = fmaf(fmaf(__uint_as_float(small[q * 4 + 3]), LO_UNSCALE, __uint_as_float(big[q * 4 + 3])), unscale, b4.w);
                        }
                        if (g + 1 < NGROUP) {  // the load registers are free again: next group on its way
                            tmem_ld_32x16_issue(lane_base + (uint32_t)((g + 1) * GCOLS), big);
                            tmem_ld_32x16_issue(lane_base + (uint32_t)(BLOCK_N + (g + 1) * GCOLS), small);
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.tmem_empty);
                }
#pragma unroll
                for (int g = 0; g < NGROUP; ++g) {
                    const int v0 = vbase + g * GCOLS;
                    if (v0 >= a.V) continue;  // warp-uniform: a column group beyond the vocabulary
                    const float *z = &zv[g * GCOLS];
                    float g0 = fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3])), g1 = fmaxf(fmaxf(z[4], z[5]), fmaxf(z[6], z[7]));
                    float g2 = fmaxf(fmaxf(z[8], z[9]), fmaxf(z[10], z[11])), g3 = fmaxf(fmaxf(z[12], z[13]), fmaxf(z[14], z[15]));
                    const float m_new = fmaxf(fmaxf(m_run, fmaxf(g0, g1)), fmaxf(g2, g3));  // finite: the group has a column inside the vocabulary
                    const float ml2 = m_new * LOG2E_F;
                    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
                    for (int j = 0; j < GCOLS; j += 2) {
                        acc0 += ex2_fast(fmaf(z[j], LOG2E_F, -ml2));
                        acc1 += ex2_fast(fmaf(z[j + 1], LOG2E_F, -ml2));
                    }
                    s_run = fmaf(s_run, ex2_fast((m_run - m_new) * LOG2E_F), acc0 + acc1);  // 2^-inf = 0 on the first group
                    m_run = m_new;
                    // rows -> shared memory -> one TMA store of the 32 x 16 box (the stores of a tile were 3.4 us of LSU / L1 time
                    // when every warp wrote its rows with st.global: r2v)
                    if (lane == 0) tma_store_wait_read();  // the previous box has been read out of the buffer
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < GCOLS / 4; ++q)
                        sts_v4(st_wr + (((uint32_t)q ^ st_wx) << 4), z[q * 4], z[q * 4 + 1], z[q * 4 + 2], z[q * 4 + 3]);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes visible to the TMA engine
                    __syncwarp();
                    if (lane == 0) tma_store_2d(&tmz, st_base, v0, box_row0);
                }
            }
            if (row_ok && a.stats != nullptr) a.stats[((size_t)row * NCHUNK + ch) * 4 + cq] = make_float2(m_run, s_run);
        }
        if (lane == 0) tma_store_wait_read();  // shared memory must outlive the last bulk store's read
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

// (rows, d) fp32 row-major -> two fp16 operand images in the layout the GEMM streams: [tile of RPT rows][k-block of 32 halves]
// [RPT rows][64 bytes], the four 16-byte chunks of a row XOR-ed with (row >> 1) & 3 -- exactly what TMA's SWIZZLE_64B (CuTe
// Swizzle<2,4,3>) would leave in shared memory, so that a tile is ONE contiguous bulk copy.  (The first versions loaded
// 64 / 128-byte row segments 2 KB apart through a tensor map: ncu showed the MMA warp waiting for operands 80 % of the time
// at a third of the L2 bandwidth.)  One warp per row: the row's largest magnitude gives its power-of-two scale (absmax == null;
// 1 / scale goes to inv_scale) or the whole tensor shares one (absmax = bits of max |x|, the weight); hi = fp16(x s),
// lo = fp16((x s - hi) 2^11).  Rows beyond `rows` (the padding of the last tile) and columns beyond d (the padding of the
// last k-block) are zero.
__global__ void __launch_bounds__(256) k_split_blocked(const float *__restrict__ x, long long rows, int d, int d_pad, int rpt, long long rows_padded,
                                                       const uint32_t *__restrict__ absmax, unsigned char *__restrict__ hi,
                                                       unsigned char *__restrict__ lo, float *__restrict__ inv_scale) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int c4n = d >> 2;        // float4 per row
    const int nchunk = d_pad >> 3;  // 16-byte chunks of 8 halves per row
    const int nkb = d_pad / BLOCK_K;
    for (long long row = warp0; row < rows_padded; row += nwarps) {
        const float4 *xr = reinterpret_cast<const float4 *>(x + row * d);
        const bool live = row < rows;
        float s, inv;
        if (absmax != nullptr) {
            scale_from_absmax(__uint_as_float(__ldg(absmax)), s, inv);
        } else {
            float m = 0.f;
            if (live)
                for (int c = lane; c < c4n; c += 32) {
                    const float4 v = __ldg(xr + c);
                    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
                }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            scale_from_absmax(m, s, inv);
            if (lane == 0 && inv_scale != nullptr) inv_scale[row] = inv;
        }
        const long long tile = row / rpt;
        const int r = (int)(row - tile * rpt);
        for (int c = lane; c < nchunk; c += 32) {
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
            if (live && 2 * c < c4n) v0 = __ldg(xr + 2 * c);
            if (live && 2 * c + 1 < c4n) v1 = __ldg(xr + 2 * c + 1);
            const float in[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            __align__(16) __half h[8], l[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float xs = in[j] * s;
                h[j] = __float2half_rn(xs);
                l[j] = __float2half_rn((xs - __half2float(h[j])) * LO_SCALE);
            }
            const int kb = c >> 2, ch = c & 3;
            const size_t off = (((size_t)tile * nkb + kb) * rpt + r) * 64 + (size_t)((ch ^ ((r >> 1) & 3)) * 16);
            *reinterpret_cast<uint4 *>(hi + off) = *reinterpret_cast<const uint4 *>(h);
            *reinterpret_cast<uint4 *>(lo + off) = *reinterpret_cast<const uint4 *>(l);
        }
    }
}

// bits of max |x| over a tensor (non-negative floats order like their bit patterns); *out zeroed by the caller
__global__ void __launch_bounds__(256) k_absmax(const float *__restrict__ x, long long n4, uint32_t *out) {
    float m = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(x) + i);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

// x -> (TF32-exact high part, fp32 remainder): hi = round-to-nearest TF32, lo = x - hi (exact in fp32)
__global__ void __launch_bounds__(256) k_split_hi_lo(const float *__restrict__ x, long long n4, float *__restrict__ hi, float *__restrict__ lo) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4 *>(x)[i];
        const float in[4] = {v.x, v.y, v.z, v.w};
        float h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            unsigned t;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(in[j]));
            h[j] = __uint_as_float(t);
            l[j] = in[j] - h[j];
        }
        reinterpret_cast<float4 *>(hi)[i] = make_float4(h[0], h[1], h[2], h[3]);
        reinterpret_cast<float4 *>(lo)[i] = make_float4(l[0], l[1], l[2], l[3]);
    }
}

// One CTA per (b, t) row: combine the row's partial statistics, then stream the row once: x = (z - max) - log(sum),
// length padding (:39-42) and the blank column (:44-46).  No reduction, no barrier: what k_init does in three phases.
__global__ void __launch_bounds__(256) k_head_finish(float *x, int ldx, const float2 *__restrict__ stats, const int64_t *__restrict__ lens,
                                                     int T, int V, int blank, float *__restrict__ blank_lp) {
    const int row = blockIdx.x;
    const int b = row / T, t = row - b * T;
    float *dst = x + (size_t)row * ldx;
    if (lens != nullptr) {
        const long long lraw = lens[b];
        long long l = lraw < 0 ? lraw + T : lraw;
        if (l < 0) l = 0;
        if (lraw < T && t >= l) {  // ctc_scorer.py:39-42
            for (int v = threadIdx.x; v < V; v += blockDim.x) dst[v] = (v == blank) ? 0.f : LZ;
            if (blank_lp != nullptr && threadIdx.x == 0) blank_lp[row] = 0.f;
            return;
        }
    }
    // every warp combines the NPART <= 32 partials on its own: one load per lane, two butterfly reductions
    static_assert(NPART <= 32, "one partial per lane");
    const int lane = threadIdx.x & 31;
    const float2 p = lane < NPART ? __ldg(&stats[(size_t)row * NPART + lane]) : make_float2(-INFINITY, 0.f);
    float m = p.x;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = p.y > 0.f ? p.y * expf(p.x - m) : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float ls = logf(s);
    if ((V & 3) == 0 && (ldx & 3) == 0) {
        const int n4 = V >> 2;
        for (int i = threadIdx.x; i < n4; i += blockDim.x) {
            const float4 z = reinterpret_cast<const float4 *>(dst)[i];
            const float4 o = make_float4((z.x - m) - ls, (z.y - m) - ls, (z.z - m) - ls, (z.w - m) - ls);
            reinterpret_cast<float4 *>(dst)[i] = o;
            if (blank_lp != nullptr && (blank >> 2) == i) blank_lp[row] = (&o.x)[blank & 3];
        }
    } else {
        for (int v = threadIdx.x; v < V; v += blockDim.x) {
            const float o = (dst[v] - m) - ls;
            dst[v] = o;
            if (v == blank && blank_lp != nullptr) blank_lp[row] = o;
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// tensor map of the (n, V) logits inside the (n, ldz) buffer for the drain's stores: 32-row x 16-column boxes, SWIZZLE_64B
int encode_z_map(CUtensorMap *tm, float *z, long long n, int V, int ldz) {
    static EncodeTiledFn enc = nullptr;
    if (enc == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return CTCPS_E_NODRIVER;
        enc = reinterpret_cast<EncodeTiledFn>(p);
    }
    cuuint64_t dims[2] = {(cuuint64_t)V, (cuuint64_t)n};
    cuuint64_t strides[1] = {(cuuint64_t)ldz * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)STORE_COLS, 32u};
    cuuint32_t estr[2] = {1, 1};
    const CUresult cr = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, z, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return cr == CUDA_SUCCESS ? 0 : CTCPS_E_NODRIVER;
}

long long padded_rows(long long rows, int rpt) { return (rows + rpt - 1) / rpt * rpt; }
int padded_k(int d) { return (d + BLOCK_K - 1) / BLOCK_K * BLOCK_K; }
size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
// one fp16 operand image of a (rows, d) matrix: rows padded to whole tiles, K to whole k-blocks
size_t image_bytes(long long rows, int rpt, int d) { return (size_t)padded_rows(rows, rpt) * padded_k(d) * sizeof(__half); }

int launch_split_blocked(const float *x, long long rows, int d, int rpt, const uint32_t *absmax, unsigned char *hi, unsigned char *lo,
                         float *inv_scale, cudaStream_t st) {
    const long long rp = padded_rows(rows, rpt);
    long long g = (rp + 7) / 8;  // a warp per row, 8 warps per CTA
    if (g > 148 * 32) g = 148 * 32;
    k_split_blocked<<<(unsigned)g, 256, 0, st>>>(x, rows, d, padded_k(d), rpt, rp, absmax, hi, lo, inv_scale);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace

extern "C" {

int ctcps_head_workspace_bytes(int64_t n, int d, size_t *out_bytes) {
    if (out_bytes == nullptr || n <= 0 || d <= 0) return CTCPS_E_BADARG;
    // blocked H1, H2 (rows padded to whole 128-row tiles) + 1 / scale of every padded row + the partial softmax statistics
    // (n, NPART) float2
    *out_bytes = 2 * align256(image_bytes(n, BLOCK_M, d)) + align256((size_t)padded_rows(n, BLOCK_M) * sizeof(float)) +
                 (size_t)n * NPART * sizeof(float2) + 512;
    return 0;
}

int ctcps_head_weight_bytes(int V, int d, size_t *out_bytes) {
    if (out_bytes == nullptr || V <= 0 || d <= 0) return CTCPS_E_BADARG;
    *out_bytes = align256(image_bytes(V, BLOCK_N, d)) + 256;  // each of w_hi, w_lo; w_hi ends with the bits of max |W|
    return 0;
}

int ctcps_head_prepare_weight(const float *weight, int V, int d, float *w_hi, float *w_lo, void *stream) {
    if (!weight || !w_hi || !w_lo || V <= 0 || d <= 0) return CTCPS_E_BADARG;
    if ((d % 16) != 0 || ((((uintptr_t)weight) | ((uintptr_t)w_hi) | ((uintptr_t)w_lo)) & 15)) return CTCPS_E_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t *absmax = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(w_hi) + align256(image_bytes(V, BLOCK_N, d)));
    cudaError_t e = cudaMemsetAsync(absmax, 0, 256, st);
    if (e != cudaSuccess) return (int)e;
    const long long n4 = (long long)V * d / 4;
    long long g = (n4 + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    k_absmax<<<(unsigned)g, 256, 0, st>>>(weight, n4, absmax);
    if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
    return launch_split_blocked(weight, V, d, BLOCK_N, absmax, reinterpret_cast<unsigned char *>(w_hi), reinterpret_cast<unsigned char *>(w_lo),
                                nullptr, st);
}

int ctcps_split_hi_lo(const float *x, int64_t count, float *hi, float *lo, void *stream) {
    if (!x || !hi || !lo || count <= 0 || (count & 3)) return CTCPS_E_BADARG;
    if ((((uintptr_t)x) | ((uintptr_t)hi) | ((uintptr_t)lo)) & 15) return CTCPS_E_ALIGN;
    const long long n4 = count >> 2;
    long long g = (n4 + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    k_split_hi_lo<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(x, n4, hi, lo);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

int ctcps_ctc_head(const float *hidden, const float *w_hi, const float *w_lo, const float *bias, const int64_t *lens, int B, int T, int d,
                   int V, int blank, int apply_log_softmax, float *x_logp, int ldx, float *blank_lp, void *workspace,
                   size_t workspace_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!hidden || !w_hi || !w_lo || !x_logp || !workspace || B <= 0 || T <= 0 || d <= 0 || V <= 0) return CTCPS_E_BADARG;
    if (blank < 0 || blank >= V || ldx < V) return CTCPS_E_BADARG;
    if ((d % 16) != 0 || (ldx & 3) != 0 || (bias != nullptr && (((uintptr_t)bias) & 15))) return CTCPS_E_ALIGN;
    if ((((uintptr_t)hidden) | ((uintptr_t)w_hi) | ((uintptr_t)w_lo) | ((uintptr_t)x_logp) | ((uintptr_t)workspace)) & 15) return CTCPS_E_ALIGN;
    const long long n = (long long)B * T;
    if (n >= (1ll << 31)) return CTCPS_E_TOOBIG;
    size_t need = 0;
    ctcps_head_workspace_bytes(n, d, &need);
    if (workspace_bytes < need) return CTCPS_E_WORKSPACE;
    unsigned char *h_hi = reinterpret_cast<unsigned char *>(workspace);
    unsigned char *h_lo = h_hi + align256(image_bytes(n, BLOCK_M, d));
    float *a_inv = reinterpret_cast<float *>(h_lo + align256(image_bytes(n, BLOCK_M, d)));
    float2 *stats = reinterpret_cast<float2 *>(reinterpret_cast<unsigned char *>(a_inv) + align256((size_t)padded_rows(n, BLOCK_M) * sizeof(float)));
    int rc = launch_split_blocked(hidden, n, d, BLOCK_M, nullptr, h_hi, h_lo, a_inv, st);
    if (rc) return rc;
    HeadArgs a;
    a.a_hi = h_hi, a.a_lo = h_lo;
    a.b_hi = reinterpret_cast<const unsigned char *>(w_hi), a.b_lo = reinterpret_cast<const unsigned char *>(w_lo);
    a.a_inv = a_inv;
    a.w_absmax = reinterpret_cast<const uint32_t *>(a.b_hi + align256(image_bytes(V, BLOCK_N, d)));
    a.bias = bias, a.z = x_logp, a.stats = apply_log_softmax ? stats : nullptr, a.n = (int)n, a.d_pad = padded_k(d), a.V = V, a.ldz = ldx;
    a.n_mtiles = (int)((n + BLOCK_M - 1) / BLOCK_M);
    a.n_ntiles = (V + BLOCK_N - 1) / BLOCK_N;
    a.tiles_per_chunk = (a.n_ntiles + NCHUNK - 1) / NCHUNK;
    CUtensorMap tmz;
    if ((rc = encode_z_map(&tmz, x_logp, n, V, ldx)) != 0) return rc;
    const size_t smem = sizeof(HeadSmem);
    cudaError_t e = cudaFuncSetAttribute(k_head_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int items = a.n_mtiles * NCHUNK;
    k_head_gemm<<<items < sms ? items : sms, HEAD_NT, smem, st>>>(tmz, a);
    if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
    if (!apply_log_softmax) return 0;  // raw logits (tests, callers that want the head alone)
    k_head_finish<<<(unsigned)n, 256, 0, st>>>(x_logp, ldx, stats, lens, T, V, blank, blank_lp);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

}  // extern "C"
