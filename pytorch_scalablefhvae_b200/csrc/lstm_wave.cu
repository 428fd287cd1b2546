// Layer-wavefront LSTM recurrence (H = 256, stacks of 1 or 2 layers) without thread-block clusters.
//
// A *group* of 8 CTAs owns 32 batch rows of one layer for all T steps; CTA j of the group owns hidden
// units [32j, 32j+32) = 128 gate columns, with its W_hh slice resident in TMEM (A operand of
// tcgen05.mma .ts) exactly like lstm_cluster.cu.  What changes:
//
//  * No clusters: only 15 clusters of 8 CTAs are co-resident on the 148 SMs of a B200 (GPC packing), but
//    a two-layer wavefront needs 2 x 8 groups.  Groups are plain CTAs (128 of them, one per SM) that
//    exchange h_t through an L2-resident buffer with an LL protocol: every 16-byte word carries 8 bytes
//    of payload and two copies of a 32-bit sequence flag {d0, flag, d1, flag}; an 8-byte half is
//    self-validating, so a consumer polls the data words themselves -- no separate flag, no fence, no
//    barrier: one L2 round trip from the producer's store to the consumer's operand buffer.
//  * Wavefront over layers: layer 1 of step t only needs h0_t, so the layer-1 groups run two steps behind
//    the layer-0 groups in the SAME launch: 2T dependent steps become T+2.  A layer-0 CTA has the gathered
//    h0_t in its MMA operand buffer one step later anyway, so it keeps the W_ih1 slice of ITS gate columns
//    resident in shared memory (bf16 hi/lo UMMA layout) and computes the layer-1 input projection
//    P1[t] = W_ih1 * h0_t behind the recurrent MMAs (a dedicated issuer warp), publishing it as LL words
//    that the layer-1 CTA of the same rank consumes like a GEMM-produced P.  The layer-1 input
//    projection GEMM and its (T,B,4H) HBM buffer disappear.
//  * Flags are unique per launch without host involvement (CUDA-graph replays reuse kernel arguments):
//    each CTA keeps a private launch counter in the header of the exchange buffer.
//
// The BPTT kernel mirrors it, top layer first: the layer-1 CTA that owns gate columns also multiplies
// its dgates_t by its W_ih1^T slice (second resident A operand) and publishes the 8 split-K partials of
// dh0_t, which the layer-0 group sums when it loads dh -- the dgrad GEMM of layer 1 disappears.
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace fhvae {

using namespace tc;

constexpr int WH = 256;            // hidden size
constexpr int WG = 8;              // CTAs per group
constexpr int WU = WH / WG;        // 32 units per CTA
constexpr int WNC = 4 * WU;        // 128 gate columns per CTA
constexpr int WNT = 512;           // compute threads per CTA
constexpr int WNTA = WNT + 128;    // + the tcgen05 issuer warp (a whole warpgroup, so that setmaxnreg can rebalance registers)
constexpr int WNB = 32;            // batch rows per group
constexpr int WSLICE = 512;        // uint4 words per published slice (2 parts x 32 rows x 4 chunks x 2 halves)
constexpr int WPSLICE = 2048;      // uint4 words per published P1 slice (32 rows x 128 gate columns fp32, 2 per word)
constexpr int WHDR = 512;          // header of the exchange buffer in uint4 (8 KB of launch counters)
constexpr int WMAXG = 32;          // counters for up to 32 groups per layer
constexpr int WMAXT = 63;          // flag = epoch * 64 + t + 1

// ---- experiment switches (tools/build_variant.sh + tools/wave_bench.py); the defaults are the measured best
#ifndef WAVE_POLL_SEQ_FWD
#define WAVE_POLL_SEQ_FWD 1  // forward h gather, wait_ll_all MODE: 1 word by word, 0 invalid words of a round together (28 more
#endif                       // registers -> spills: 173 vs 116 us/launch), 2 re-read everything per round
#ifndef WAVE_POLL_SEQ_BWD
#define WAVE_POLL_SEQ_BWD 0  // BPTT reduce-scatter: MODE 0 (measured: 111 vs 134 us/launch with MODE 1)
#endif
#ifndef WAVE_P_FIRST
#define WAVE_P_FIRST 0       // 1: request this step's input projection before the exchange instead of after it
#endif
#ifndef WAVE_SENTINEL
#define WAVE_SENTINEL 0      // 1: spin on ONE word of the last peer slice before the bulk loads of the exchange
#endif
#ifndef WAVE_FLAGPOLL
#define WAVE_FLAGPOLL 0      // 1: ONE warp polls one sentinel word per producer; the other warps sleep at a named barrier
#endif                       //    and read the LL words once they are (almost surely) there -- no poll spam in L2
#ifndef WAVE_TMA
#define WAVE_TMA 0           // 1: forward h exchange = plain payload + release/acquire flag + cp.async.bulk straight into the
#endif                       //    MMA operand buffer (half the bytes of LL words, no register staging, no compute-warp work)
#ifndef WAVE_FWD2
#define WAVE_FWD2 0          // 1: forward as two independent 16-row chains per CTA (lstm_wave_fwd2_kernel).  Bit-identical, but
#endif                       // measured SLOWER (131 vs 105 us per launch): see the kernel's header comment
#ifndef WAVE_FWD2_QUIET
#define WAVE_FWD2_QUIET 0
#endif
#ifndef WAVE_COOPERATIVE
#define WAVE_COOPERATIVE -1  // 1 / 0: always / never launch cooperatively; -1: env FHVAE_WAVE_COOPERATIVE=1 decides (default off)
#endif
#ifndef WAVE_PDL_EARLY
#define WAVE_PDL_EARLY 0      // 1: griddepcontrol.launch_dependents at the top of the wavefront kernels (the launches behind them
#endif                        //    become resident on the free SMs while the recurrence runs): +38 us per step -- they take the
                              //    SMs the side-stream weight gradients live on
#ifndef WAVE_SAVE_FIRST
#define WAVE_SAVE_FIRST 0    // 1: the HBM stores of the previous step are issued before the exchange loads
#endif

__device__ __forceinline__ uint4 ld_ll(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_ll(uint4* p, uint32_t d0, uint32_t d1, uint32_t flag) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(d0), "r"(flag), "r"(d1), "r"(flag) : "memory");
}
__device__ __forceinline__ uint32_t wait_ll(uint4& v, const uint4* p, uint32_t flag) {
    uint32_t spins = 0;
    while (v.y != flag || v.w != flag) {
        if (++spins > FHVAE_SPIN_LIMIT) __trap();
        v = ld_ll(p);
    }
    return spins;
}

// ---- flag + bulk-copy exchange (WAVE_TMA) ------------------------------------------------------------------
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_flag(const uint32_t* p, uint32_t flag) {
    uint32_t spins = 0;
    while (ld_acquire_gpu(p) != flag)
        if (++spins > FHVAE_SPIN_LIMIT) __trap();
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy (TMA engine, async proxy); completion is counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
#ifndef WAVE_TMA_FENCE
#define WAVE_TMA_FENCE 1     // consumer-side proxy fence before the bulk copies: 0 none, 1 .global, 2 all state spaces
#endif
__device__ __forceinline__ void fence_proxy_async_all() {
#if WAVE_TMA_FENCE == 2
    asm volatile("fence.proxy.async;" ::: "memory");
#elif WAVE_TMA_FENCE == 1
    asm volatile("fence.proxy.async.global;" ::: "memory");
#endif
}

// Wait for N LL words at once: every retry round re-issues ALL still-invalid loads back to back, so a round costs
// one L2 round trip (~260 cycles) however many words are outstanding.  (Waiting word by word -- wait_ll in a loop --
// serialises the round trips: 7 peer slices that all miss their first poll cost 7 dependent trips, which was
// 2900-3700 of the 10600 cycles of a forward wavefront step in round 1.)
template <int N, int MODE, typename AddrFn>
__device__ __forceinline__ uint32_t wait_ll_all(uint4 (&v)[N], AddrFn addr, uint32_t flag) {
    uint32_t spins = 0;
    if constexpr (MODE == 1) {
        // word by word: a word that missed its first poll costs its own dependent L2 round trip
#pragma unroll
        for (int i = 0; i < N; ++i) spins += wait_ll(v[i], addr(i), flag);
        return spins;
    } else if constexpr (MODE == 2) {
        // re-read EVERYTHING until every word is valid: one round trip per retry round, no predicates, no extra
        // registers (valid words stay valid within a step, so re-reading them is harmless)
        for (;;) {
            bool ok = true;
#pragma unroll
            for (int i = 0; i < N; ++i) ok = ok && v[i].y == flag && v[i].w == flag;
            if (ok) return spins;
            if (++spins > FHVAE_SPIN_LIMIT) __trap();
#pragma unroll
            for (int i = 0; i < N; ++i) v[i] = ld_ll(addr(i));
        }
    } else {
        // re-issue only the still-invalid words of a round together (needs the retry loads in separate registers)
        for (;;) {
            uint32_t bad = 0;
#pragma unroll
            for (int i = 0; i < N; ++i) bad |= (v[i].y != flag || v[i].w != flag) ? (1u << i) : 0u;
            if (!bad) return spins;
            if (++spins > FHVAE_SPIN_LIMIT) __trap();
#pragma unroll
            for (int i = 0; i < N; ++i)
                if ((bad >> i) & 1u) v[i] = ld_ll(addr(i));
        }
    }
}

#ifdef FHVAE_TIMELINE
__device__ long long g_wave_tl[2][32][16];
#define WTL(step, slot) do { if (rank == 0 && grp == 0 && threadIdx.x == 0 && (step) < 32) g_wave_tl[layer][step][slot] = clock64(); } while (0)
#define WTL1(step, slot) do { if (rank == 0 && grp == 0 && (step) < 32) g_wave_tl[layer][step][slot] = clock64(); } while (0)   // caller = one thread
extern "C" int fhvae_debug_wave_timeline(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_wave_tl, sizeof(g_wave_tl)); }
#else
#define WTL(step, slot) do { } while (0)
#define WTL1(step, slot) do { } while (0)
#endif

// ================================================================================================
// Pre-packed weight operands.  Staging a CTA's W_hh slice into TMEM (and, for the cross-layer products, its W_ih1
// slice into shared memory) from the fp32 weights costs 17-30 k cycles per launch (strided 16-byte loads of rows
// 1 KB apart + fp32 -> bf16 hi/lo conversion, by 128 CTAs at once): 9-15 us of every one of the six recurrent
// launches of a step.  wave_pack_kernel does the conversion ONCE per step (the weights only change in Adam) into
// images laid out exactly as the kernels consume them:
//   T images (128 x 256 bf16 per part, for tcgen05.st): uint4 [warp 16][v 8*parts][lane 32] -- a warp's load
//            instruction reads 512 contiguous bytes; v = part*8 + hf*4 + i holds the 4 words {4i..4i+3} of the
//            16-word tcgen05.st of half hf.
//   S images (the UMMA no-swizzle K-major shared-memory operand, 64 KB per part): copied verbatim by cp.async.bulk.
// Order in the buffer: fwdT[layer][rank], fwdS[rank] (2-layer stacks), bwdT[layer][rank], bwdS[rank].
// ================================================================================================
template <bool X3, int CH = WH> struct WavePack {
    static constexpr int NG = CH / WU, HF = CH / 128;                  // CTAs per group, 128-unit halves
    static constexpr size_t IMG = (size_t)(X3 ? 2 : 1) * HF * 32768;   // every image: parts x CH x 128 bf16
    __host__ __device__ static constexpr int n_images(int L) { return 2 * NG * L + (L == 2 ? 2 * NG : 0); }
    __host__ __device__ static constexpr size_t fwdT(int L, int layer, int rank) { return (size_t)(layer * NG + rank) * IMG; }
    __host__ __device__ static constexpr size_t fwdS(int L, int rank) { return (size_t)(L * NG + rank) * IMG; }
    __host__ __device__ static constexpr size_t bwdT(int L, int layer, int rank) {
        return (size_t)(L * NG + (L == 2 ? NG : 0) + layer * NG + rank) * IMG;
    }
    __host__ __device__ static constexpr size_t bwdS(int L, int rank) { return (size_t)(2 * L * NG + NG + rank) * IMG; }
};

struct WavePackArgs { const float* Whh[2]; const float* Wih1; uint8_t* out; int L; };

template <bool X3, int CH>
__global__ void __launch_bounds__(WNT) wave_pack_kernel(const __grid_constant__ WavePackArgs a) {
    using PK = WavePack<X3, CH>;
    constexpr int UC = WU, NC = WNC, NT = WNT, WG = CH / WU, HF = CH / 128, KQ = CH / 4;   // KQ: k-values per column group
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, cg = (warp >> 2) & 3;
    const int L = a.L;
    int img = blockIdx.x;
    const int nT = L * WG, nS = (L == 2) ? WG : 0;
    auto store_T = [&](uint8_t* dst, const uint32_t (&w)[HF * 16], int part) {     // HF*16 words: [hf][16]
        uint4* o = reinterpret_cast<uint4*>(dst) + (size_t)(warp * ((X3 ? 2 : 1) * HF * 4) + part * HF * 4) * 32 + lane;
#pragma unroll
        for (int v = 0; v < HF * 4; ++v) o[v * 32] = make_uint4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
    };
    if (img < nT) {
        // ---- forward TMEM image: lane n = gate q * 32 + unit, 32-bit column c = (k, k+1) pair, k = 2c
        const int layer = img / WG, rank = img % WG;
        const float* src = a.Whh[layer] + (size_t)(q * CH + rank * UC + lane) * CH + cg * KQ;
        uint32_t hi[HF * 16], lo[HF * 16];
#pragma unroll
        for (int i = 0; i < KQ / 4; ++i) {
            const float4 w4 = __ldcg(reinterpret_cast<const float4*>(src) + i);
            const float v[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint32_t h = pack_bf16(v[2 * j], v[2 * j + 1]);
                hi[2 * i + j] = h;
                lo[2 * i + j] = pack_bf16(v[2 * j] - __uint_as_float(h << 16), v[2 * j + 1] - __uint_as_float(h & 0xffff0000u));
            }
        }
        uint8_t* dst = a.out + PK::fwdT(L, layer, rank);
        store_T(dst, hi, 0);
        if (X3) store_T(dst, lo, 1);
        return;
    }
    img -= nT;
    if (img < nS) {
        // ---- forward shared-memory image of the W_ih1 slice: [part][K chunk CH/8][row 128] x 16 B
        const int rank = img;
        const float* s1 = a.Wih1 + (size_t)(q * CH + rank * UC + lane) * CH + cg * KQ;
        const int r = q * 32 + lane;
        uint8_t* dst = a.out + PK::fwdS(L, rank);
#pragma unroll 2
        for (int i = 0; i < KQ / 8; ++i) {
            const float4 x0 = __ldcg(reinterpret_cast<const float4*>(s1) + 2 * i);
            const float4 x1 = __ldcg(reinterpret_cast<const float4*>(s1) + 2 * i + 1);
            const float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            const uint32_t off = (uint32_t)(cg * (KQ / 8) + i) * (NC * 16) + (uint32_t)r * 16;
            uint4 hi, lo;
            split_bf16(v, hi, lo);
            *reinterpret_cast<uint4*>(dst + off) = hi;
            if (X3) *reinterpret_cast<uint4*>(dst + PK::IMG / 2 + off) = lo;
        }
        return;
    }
    img -= nS;
    if (img < nT) {
        // ---- BPTT TMEM image of W_hh^T: lane = unit n (HF halves of 128), column = pair of the CTA's 128 gate columns
        const int layer = img / WG, rank = img % WG, g = cg;
        const float* W = a.Whh[layer];
        uint32_t hi[HF * 16], lo[HF * 16];
#pragma unroll
        for (int hf = 0; hf < HF; ++hf) {
            const int n = hf * 128 + q * 32 + lane;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float v0 = __ldcg(W + (size_t)(g * CH + rank * UC + 2 * j) * CH + n);
                const float v1 = __ldcg(W + (size_t)(g * CH + rank * UC + 2 * j + 1) * CH + n);
                const uint32_t h = pack_bf16(v0, v1);
                hi[hf * 16 + j] = h;
                lo[hf * 16 + j] = pack_bf16(v0 - __uint_as_float(h << 16), v1 - __uint_as_float(h & 0xffff0000u));
            }
        }
        uint8_t* dst = a.out + PK::bwdT(L, layer, rank);
        store_T(dst, hi, 0);
        if (X3) store_T(dst, lo, 1);
        return;
    }
    img -= nT;
    {
        // ---- BPTT shared-memory image of W_ih1^T: [part][unit half HF][K chunk 16][unit 128] x 16 B
        const int rank = img;
        uint8_t* dst = a.out + PK::bwdS(L, rank);
        constexpr int WIT = CH * (NC / 8) / NT;
#pragma unroll 2
        for (int i = 0; i < WIT; ++i) {
            const int item = tid + i * NT;
            const int n = item % CH, kc = item / CH;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = kc * 8 + j, g = k >> 5, u = k & 31;
                v[j] = __ldcg(a.Wih1 + (size_t)(g * CH + rank * UC + u) * CH + n);
            }
            const uint32_t off = (uint32_t)(n >> 7) * (128 * NC * 2) + (uint32_t)(kc * 128 + (n & 127)) * 16;
            uint4 hi, lo;
            split_bf16(v, hi, lo);
            *reinterpret_cast<uint4*>(dst + off) = hi;
            if (X3) *reinterpret_cast<uint4*>(dst + PK::IMG / 2 + off) = lo;
        }
    }
}

// packed T image -> this thread's TMEM columns: HF*4 coalesced 16-byte loads + HF tcgen05.st.x16 per bf16 part
template <bool X3, int HF>
__device__ __forceinline__ void wave_load_T_image(const uint8_t* img, int warp, int lane, uint32_t ta, uint32_t half_stride,
                                                  uint32_t part_stride) {
#pragma unroll
    for (int part = 0; part < (X3 ? 2 : 1); ++part) {
        const uint4* src = reinterpret_cast<const uint4*>(img) + (size_t)(warp * ((X3 ? 2 : 1) * HF * 4) + part * HF * 4) * 32 + lane;
        uint4 u[HF * 4];
#pragma unroll
        for (int v = 0; v < HF * 4; ++v) u[v] = __ldcg(src + v * 32);
#pragma unroll
        for (int hf = 0; hf < HF; ++hf) {
            const uint32_t w[16] = {u[4 * hf].x,     u[4 * hf].y,     u[4 * hf].z,     u[4 * hf].w,
                                    u[4 * hf + 1].x, u[4 * hf + 1].y, u[4 * hf + 1].z, u[4 * hf + 1].w,
                                    u[4 * hf + 2].x, u[4 * hf + 2].y, u[4 * hf + 2].z, u[4 * hf + 2].w,
                                    u[4 * hf + 3].x, u[4 * hf + 3].y, u[4 * hf + 3].z, u[4 * hf + 3].w};
            tmem_st16(ta + part * part_stride + hf * half_stride, w);
        }
    }
}
// packed S image -> shared memory, 16-KB bulk copies counted on `bar` (issued by one thread)
__device__ __forceinline__ void wave_load_S_image(const uint8_t* img, uint32_t dst_smem, uint32_t bytes, uint64_t* bar) {
    mbar_arrive_expect_tx(bar, bytes);
    for (uint32_t o = 0; o < bytes; o += 16384) bulk_g2s(dst_smem + o, img + o, 16384, bar);
}

struct WaveFwdArgs {
    const float* P0; const float* Q0; const float* Whh0; float* h0; float* c0; float* a0;
    const float* Wih1; const float* b1; const float* Whh1; float* h1; float* c1; float* a1;
    uint4* xchg;
    int T, B, b_off, G, Gs, L;     // B = row stride (whole batch), G groups in this launch, Gs = slot stride
    __nv_bfloat16* hp0; __nv_bfloat16* hp1; long long hps;   // optional bf16 hi/lo planes of h (lo at +hps), for gemm_wgrad.cu
    const uint8_t* packed;          // optional pre-packed operand images of the weights (wave_pack_kernel); NULL: convert here
};

template <bool X3, int CH = WH>
struct WaveFwdSmem {
    static constexpr int H_PART = WNB * CH * 2;                // 32 x 256 bf16 = 16 KB
    static constexpr int H_BUF = (X3 ? 2 : 1) * H_PART;
    static constexpr int H_OFF = 0;                            // h_{t-1} operand, double-buffered
    static constexpr int G_OFF = 2 * H_BUF;
    static constexpr int G_BYTES = 4 * WNB * (WU + 1) * 4;     // gates[4][NB][33] fp32
    static constexpr int BAR_OFF = G_OFF + G_BYTES;
    static constexpr int W_OFF = (BAR_OFF + 256 + 1023) / 1024 * 1024;
    static constexpr int W_PART = WNC * CH * 2;                // 128 x 256 bf16 = 64 KB
    static constexpr int W_BYTES = (X3 ? 2 : 1) * W_PART;
    static constexpr int TOTAL1 = W_OFF;                       // single layer: no resident W_ih
    static constexpr int TOTAL2 = W_OFF + W_BYTES;
};

// pull NS slices (all 8, or the 7 peers) of one published step into an MMA operand buffer
template <bool X3, int CH, int NS, bool SKIP_OWN>
__device__ __forceinline__ void wave_load(const uint4* slot, int rank, uint4 (&v)[NS], uint32_t flag) {
    constexpr int NW = (X3 ? 2 : 1) * 256;
    if ((int)threadIdx.x < NW) {
#if WAVE_SENTINEL
        {   // one word of the LAST slice first: when it has landed the others almost surely have too, so the bulk
            // pass below is not a wasted 57-KB read of words that are still in flight
            const int src = SKIP_OWN ? (NS - 1) + ((NS - 1) >= rank ? 1 : 0) : NS - 1;
            uint4 s = ld_ll(slot + src * WSLICE + threadIdx.x);
            wait_ll(s, slot + src * WSLICE + threadIdx.x, flag);
        }
#endif
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const int src = SKIP_OWN ? i + (i >= rank ? 1 : 0) : i;
            v[i] = ld_ll(slot + src * WSLICE + threadIdx.x);
        }
    }
}
template <bool X3, int CH, int NS, bool SKIP_OWN>
__device__ __forceinline__ void wave_store(const uint4* slot, int rank, uint32_t flag, uint4 (&v)[NS], uint8_t* dst) {
    using S = WaveFwdSmem<X3, CH>;
    constexpr int NW = (X3 ? 2 : 1) * 256;
    const int tid = threadIdx.x;
    if (tid < NW) {
        const int part = tid >> 8, rem = tid & 255, chunk = rem >> 1, half = rem & 1;
        const int kcl = chunk / WNB, row = chunk % WNB;
        wait_ll_all<NS, WAVE_POLL_SEQ_FWD>(v, [&](int i) { return slot + (SKIP_OWN ? i + (i >= rank ? 1 : 0) : i) * WSLICE + tid; }, flag);
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const int src = SKIP_OWN ? i + (i >= rank ? 1 : 0) : i;
            *reinterpret_cast<uint2*>(dst + part * S::H_PART + (uint32_t)(src * 4 + kcl) * (WNB * 16) + row * 16 + half * 8) =
                make_uint2(v[i].x, v[i].z);
        }
    }
}
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(WNT) : "memory"); }

// Roles (CTA-uniform): layer 0 of a 2-layer stack additionally owns the layer-1 input projection
//   P1[t] = W_ih1_slice * h0_t^T   (its h operand buffer already holds the gathered h0_t one step later),
// issued by the dedicated MMA warp behind the recurrent MMAs and published as LL words; layer 1 consumes it
// exactly like layer 0 consumes the GEMM-produced P0.  Warps 0..15 = gate/cell math + exchange, warp 16 =
// tcgen05 issuer (so the ~2000-cycle SS-mode projection never blocks a compute warp).
template <bool X3, int CH_>
__global__ void __launch_bounds__(WNTA, 1) lstm_wave_fwd_kernel(const __grid_constant__ WaveFwdArgs a) {
    using S = WaveFwdSmem<X3, CH_>;
    constexpr int NB = WNB, NT = WNT, CH = CH_, UC = WU;
    constexpr int WG = CH / WU, HF = CH / 128, KQ = CH / 4;   // CTAs per group (shadows the H = 256 constant), 128-k halves, k per column group
    constexpr int NW = (X3 ? 2 : 1) * 256;     // LL words of one slice actually used
    constexpr int NCG = NT / 128;              // column groups: warps sharing one TMEM lane quarter
    constexpr int CPW = NB / NCG;              // batch rows (TMEM columns) per thread in the gate phase
    constexpr int RPT = NB * 32 / NT;          // batch rows per thread in the cell phase
    constexpr int H4 = 4 * CH;
    static_assert(CPW == 8 && RPT == 2, "LL packing of P1 assumes 8 rows per thread");
    static_assert(CH / 16 == 2 * WG, "two K=16 steps per exchanged slice");
    extern __shared__ __align__(1024) uint8_t smem[];
    float (*gates)[NB][UC + 1] = reinterpret_cast<float (*)[NB][UC + 1]>(smem + S::G_OFF);
    uint64_t* hb_full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);   // [2 buffers] compute warps -> issuer
    uint64_t* rec_done = hb_full + 2;                                     // issuer (commit) -> compute warps
    uint64_t* p1_done = rec_done + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p1_done + 1);
    uint32_t* epoch_slot = tmem_slot + 1;
    volatile uint32_t* p1_safe = epoch_slot + 1;    // WAVE_TMA: last step whose projection MMAs are known complete
    uint64_t* w_full = hb_full + 6;                 // packed W_ih1 image landed in shared memory (bulk copies)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, cg = (warp >> 2) & 3;   // gate (= TMEM lane quarter), column group
    const int layer = blockIdx.x / (a.G * WG);
    const int grp = (blockIdx.x / WG) % a.G;
    const int rank = blockIdx.x % WG;
    const int T = a.T, B = a.B;
    const int b0 = a.b_off + grp * NB;
    const bool p1_duty = (layer == 0 && a.L == 2);
    const int nsteps = p1_duty ? T + 1 : T;
    const float* P = layer ? nullptr : a.P0;
    const float* Q = layer ? nullptr : a.Q0;
    const float* W_hh = layer ? a.Whh1 : a.Whh0;
    float* h_all = layer ? a.h1 : a.h0;
    __nv_bfloat16* h_pl = layer ? a.hp1 : a.hp0;
    float* c_all = layer ? a.c1 : a.c0;
    float* acts = layer ? a.a1 : a.a0;
    // exchange buffer: [header][h: layer][group][t][src CTA][WSLICE] uint4, then [p1: group][t][CTA][WPSLICE]
    uint32_t* cnt = reinterpret_cast<uint32_t*>(a.xchg) + (((a.L - 1) * 2 + layer) * WMAXG + grp) * WG + rank;
    uint4* own = a.xchg + WHDR + ((size_t)(layer * a.Gs + grp) * T) * (WG * WSLICE);
    uint4* p1x = a.xchg + WHDR + ((size_t)(2 * a.Gs) * T) * (WG * WSLICE) + ((size_t)grp * T) * (WG * WPSLICE) + rank * WPSLICE;

    WTL(0, 15);
#if WAVE_PDL_EARLY
    pdl_launch_dependents();
#endif
    // ---- prologue: TMEM, barriers, W_hh slice -> TMEM (lane n = gate*32 + unit, column k/2 = bf16 pair)
    constexpr int NACC = 2;
    constexpr int WCOLS = CH / 2;
    constexpr int ACOL = (X3 ? 2 : 1) * WCOLS;
    constexpr int TCOLS = (ACOL + 2 * NACC * NB) <= 256 ? 256 : 512;
    if (warp == NT / 32) tmem_alloc<TCOLS>(tmem_slot);
    if (tid == 32) {
        // WAVE_TMA: + one arrive.expect_tx for the 7 peer slices (pull lanes of warp 17)
        mbar_init(&hb_full[0], NT / 32 + (WAVE_TMA ? 1 : 0)); mbar_init(&hb_full[1], NT / 32 + (WAVE_TMA ? 1 : 0));
        mbar_init(rec_done, 1); mbar_init(p1_done, 1); mbar_init(w_full, 1);
        fence_mbar_init();
        *p1_safe = 0;
        if (p1_duty && a.packed)
            wave_load_S_image(a.packed + WavePack<X3, CH>::fwdS(a.L, rank), smem_u32(smem + S::W_OFF), S::W_BYTES, w_full);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    uint32_t fbase = 0;              // flag base = launch epoch << 6: read by the compute warps after pdl_wait()
    WTL(3, 15);
    constexpr uint32_t W_LBO = WNC * 16, H_LBO = NB * 16, SBO_ = 128;
    const uint32_t tmem_d = tmem_base + ACOL;                  // recurrent accumulators [NACC][NB]
    const uint32_t tmem_p = tmem_d + NACC * NB;                // layer-1 projection accumulators [NACC][NB]
    constexpr uint32_t idesc = make_idesc_bf16(WNC, NB);
    const uint32_t hb_u = smem_u32(smem + S::H_OFF), wih_u = smem_u32(smem + S::W_OFF);

    if (warp >= NT / 32) {
        // ================= tcgen05 issuer (warp 16; warps 17..19 only donate their registers) =================
        reg_dec<32>();
        tc_fence_before();
        __syncthreads();                                        // operands staged by the compute warps
        tc_fence_after();
        if (warp == NT / 32 && elect_one()) {
            if (p1_duty && a.packed) mbar_wait(w_full, 0);
            for (int t = 1; t < nsteps; ++t) {
                const uint32_t hbt = hb_u + (t & 1) * S::H_BUF;
                const uint64_t dhh0 = make_smem_desc(hbt, H_LBO, SBO_);
                const uint64_t dhl0 = make_smem_desc(hbt + S::H_PART, H_LBO, SBO_);
                mbar_wait(&hb_full[t & 1], ((t - 1) >> 1) & 1);        // h_{t-1} gathered in buffer t&1
                WTL1(t, 13);
                tc_fence_after();
                if (t < T) {
                    // rolled over the 8 slices (this warp runs on 32 registers)
#pragma unroll 1
                    for (int j = 0; j < WG; ++j) {
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            const int s = 2 * j + ks;
                            const uint64_t ih = (uint64_t)((s * 2 * H_LBO) >> 4);
                            const uint32_t awh = tmem_base + (uint32_t)(s * 8);
                            const uint32_t td = tmem_d + (uint32_t)(ks * NB);       // NACC == 2: one accumulator per ks
                            const uint32_t first = j > 0 ? 1u : 0u;
                            if (X3) {
                                umma_bf16_ts(td, awh + WCOLS, dhh0 + ih, idesc, first);
                                umma_bf16_ts(td, awh, dhl0 + ih, idesc, 1u);
                                umma_bf16_ts(td, awh, dhh0 + ih, idesc, 1u);
                            } else {
                                umma_bf16_ts(td, awh, dhh0 + ih, idesc, first);
                            }
                        }
                    }
                }
                if (t < T) umma_commit(rec_done);
                WTL1(t, 14);
                if (p1_duty) {
                    uint64_t dh = dhh0, dl = dhl0;
                    uint64_t dwh = make_smem_desc(wih_u, W_LBO, SBO_), dwl = make_smem_desc(wih_u + S::W_PART, W_LBO, SBO_);
#pragma unroll 1
                    for (int s = 0; s < CH / 16; ++s) {
                        const uint32_t td = tmem_p + (uint32_t)((s % NACC) * NB);
                        const uint32_t first = s >= NACC ? 1u : 0u;
                        if (X3) {
                            umma_bf16(td, dwl, dh, idesc, first);
                            umma_bf16(td, dwh, dl, idesc, 1u);
                            umma_bf16(td, dwh, dh, idesc, 1u);
                        } else {
                            umma_bf16(td, dwh, dh, idesc, first);
                        }
                        dh += (uint64_t)((2 * H_LBO) >> 4); dl += (uint64_t)((2 * H_LBO) >> 4);
                        dwh += (uint64_t)((2 * W_LBO) >> 4); dwl += (uint64_t)((2 * W_LBO) >> 4);
                    }
                    umma_commit(p1_done);
                }
            }
        }
#if WAVE_TMA
        else if (warp == NT / 32 + 1 && lane < WG - 1) {
            // ---- pull lanes: lane i owns peer slice src(i).  Poll the slice's flag (acquire), then bulk-copy the
            // payload (already in the UMMA operand layout) straight into the operand buffer of step t; the issuer's
            // hb_full barrier counts the bytes.  No compute warp touches the exchange.
            const int src = lane + (lane >= rank ? 1 : 0);
            pdl_wait();
            fbase = *reinterpret_cast<volatile uint32_t*>(cnt) << 6;
            constexpr uint32_t PART_BYTES = 4 * H_LBO;                       // 4 K-chunks x 32 rows x 16 B = 2 KB
            constexpr uint32_t PULL_MASK = (1u << (WG - 1)) - 1;
            for (int t = 1; t < nsteps; ++t) {
                const uint4* sl = own + (size_t)(t - 1) * (WG * WSLICE) + src * WSLICE;
                // buffer t&1 was last read by the projection MMAs of step t-2 (p1_duty only; the recurrent MMAs of
                // step t-2 are complete by causality: no peer can publish h_{t-1} before it received our h_{t-2})
                if (p1_duty && t >= 3) {
                    uint32_t spins = 0;
                    while (*p1_safe < (uint32_t)(t - 2))
                        if (++spins > FHVAE_SPIN_LIMIT) __trap();
                }
                // the 7 lanes poll in lockstep (one converged loop: no divergent re-issue of the copies)
                const uint32_t* fp = reinterpret_cast<const uint32_t*>(sl + 256);
                bool seen = false;
                uint32_t spins = 0;
                do {
                    if (!seen) seen = ld_acquire_gpu(fp) == fbase + t;
                    if (++spins > FHVAE_SPIN_LIMIT) __trap();
                } while (!__all_sync(PULL_MASK, seen));
                if (lane == WG - 2) WTL1(t, 11);
                fence_proxy_async_all();
                uint64_t* bar = &hb_full[t & 1];
                if (lane == 0) mbar_arrive_expect_tx(bar, (WG - 1) * (X3 ? 2u : 1u) * PART_BYTES);
                __syncwarp(PULL_MASK);
                const uint32_t dst = hb_u + (t & 1) * S::H_BUF + (uint32_t)src * PART_BYTES;
                bulk_g2s(dst, sl, PART_BYTES, bar);
                if (X3) bulk_g2s(dst + S::H_PART, reinterpret_cast<const uint8_t*>(sl) + PART_BYTES, PART_BYTES, bar);
                if (lane == WG - 2) WTL1(t, 12);
            }
        }
#endif
        __syncwarp();
    } else {
        // ================= compute warps =================
        reg_inc<112>();   // 20 warps x 96 regs at launch = 16 x 112 + 4 x 32 (setmaxnreg only moves registers inside the CTA)
        WTL(4, 15);
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * (KQ / 2));
        if (a.packed) {
            wave_load_T_image<X3, HF>(a.packed + WavePack<X3, CH>::fwdT(a.L, layer, rank), warp, lane, ta, 16, WCOLS);
        } else {
            const float* src = W_hh + (size_t)(q * CH + rank * UC + lane) * CH + cg * KQ;
            float4 wv[KQ / 4];
#pragma unroll
            for (int i = 0; i < KQ / 4; ++i) wv[i] = __ldcg(reinterpret_cast<const float4*>(src) + i);
#pragma unroll
            for (int hf = 0; hf < HF; ++hf) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 w4 = wv[hf * 8 + i];
                    const float v[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        hi[2 * i + j] = pack_bf16(v[2 * j], v[2 * j + 1]);
                        lo[2 * i + j] = pack_bf16(v[2 * j] - __uint_as_float(hi[2 * i + j] << 16),
                                                  v[2 * j + 1] - __uint_as_float(hi[2 * i + j] & 0xffff0000u));
                    }
                }
                tmem_st16(ta + hf * 16, hi);
                if (X3) tmem_st16(ta + WCOLS + hf * 16, lo);
            }
        }
        tmem_wait_st();
        WTL(5, 15);
        if (p1_duty && !a.packed) {
            // resident W_ih1 slice as an smem A operand: row n = gate*32 + unit, K-chunk planes of W_LBO bytes
            const float* s1 = a.Wih1 + (size_t)(q * CH + rank * UC + lane) * CH + cg * KQ;
            const int r = q * 32 + lane;
#pragma unroll 2
            for (int i = 0; i < KQ / 8; ++i) {
                const float4 x0 = __ldcg(reinterpret_cast<const float4*>(s1) + 2 * i);
                const float4 x1 = __ldcg(reinterpret_cast<const float4*>(s1) + 2 * i + 1);
                const float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                const uint32_t off = (uint32_t)(cg * (KQ / 8) + i) * W_LBO + (uint32_t)r * 16;
                if (X3) {
                    uint4 hi, lo;
                    split_bf16(v, hi, lo);
                    *reinterpret_cast<uint4*>(smem + S::W_OFF + off) = hi;
                    *reinterpret_cast<uint4*>(smem + S::W_OFF + S::W_PART + off) = lo;
                } else {
                    *reinterpret_cast<uint4*>(smem + S::W_OFF + off) =
                        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                }
            }
        }
        WTL(6, 15);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        WTL(7, 15);
        // everything above touched only this launch's own state and the weight images (older than any launch still
        // in flight); from here on: outputs of the preceding launches, the exchange buffer and its epoch counter
        pdl_wait();
        fbase = *reinterpret_cast<volatile uint32_t*>(cnt) << 6;
        // time-invariant addend for this thread's (gate q, unit lane) column, rows cg*CPW ..
        const int col = q * CH + rank * UC + lane;
        float qv[CPW];
        {
            const float bias = (layer == 1 && a.b1) ? __ldcg(a.b1 + col) : 0.f;
#pragma unroll
            for (int b = 0; b < CPW; ++b) qv[b] = bias + (Q ? __ldcg(Q + (size_t)(b0 + cg * CPW + b) * H4 + col) : 0.f);
        }
        WTL(1, 15);
        float creg[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) creg[i] = 0.f;
        const int kglob = rank * UC + lane;                        // this thread's unit in the cell phase
        const uint32_t hoff_k = (uint32_t)(kglob >> 3) * H_LBO + (uint32_t)(kglob & 7) * 2;
        const int llw = (cg * 4) * 128 + q * 32 + lane;            // this thread's first P1 word (+128 per row pair)
        float aval[CPW], hreg[RPT];
        // saved-for-backward state of step ts -> HBM; issued while the next step's exchange loads are in flight
        auto store_saved = [&](int ts) {
            float* At = acts + ((size_t)ts * B + b0 + cg * CPW) * H4 + col;
#pragma unroll
            for (int b = 0; b < CPW; ++b) At[(size_t)b * H4] = aval[b];
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const size_t o = ((size_t)ts * B + b0 + warp * RPT + i) * CH + kglob;
                h_all[o] = hreg[i];
                c_all[o] = creg[i];
                if (h_pl) {                                   // x = hi + lo planes: the weight-gradient GEMM's TMA operand
                    const __nv_bfloat16 hh = __float2bfloat16_rn(hreg[i]);
                    h_pl[o] = hh;
                    h_pl[o + a.hps] = __float2bfloat16_rn(hreg[i] - __bfloat162float(hh));
                }
            }
        };

        for (int t = 0; t < nsteps; ++t) {
            const bool real = t < T;
            WTL(t, 0);
            uint8_t* hbt = smem + S::H_OFF + (t & 1) * S::H_BUF;
            float pv[CPW];
            uint4 pl[CPW / 2];
            auto request_p = [&]() {
                if (real) {
                    if (layer == 0) {
                        const float* Pt = P ? P + ((size_t)t * B + b0 + cg * CPW) * H4 + col : nullptr;
#pragma unroll
                        for (int b = 0; b < CPW; ++b) pv[b] = Pt ? __ldcg(Pt + (size_t)b * H4) : 0.f;
                    } else {
#pragma unroll
                        for (int j = 0; j < CPW / 2; ++j) pl[j] = ld_ll(p1x + (size_t)t * (WG * WPSLICE) + llw + j * 128);
                    }
                }
            };
#if WAVE_P_FIRST
            request_p();
#endif
#if WAVE_TMA
            if (t > 0) store_saved(t - 1);   // (the pull lanes of warp 17 gather h_{t-1}; nothing to do here)
#else
            if (t > 0) {
                // h_{t-1}: the 7 peer slices (own slice was written locally by the cell phase)
                const uint4* slot = own + (size_t)(t - 1) * (WG * WSLICE);
                uint4 hv[WG - 1];
#if WAVE_SAVE_FIRST
                store_saved(t - 1);
#endif
#if WAVE_FLAGPOLL
                if (warp == 0) {
                    if (lane < WG - 1) {
                        const uint4* sp = slot + (lane + (lane >= rank ? 1 : 0)) * WSLICE + (NW - 1);
                        uint4 sv = ld_ll(sp);
                        wait_ll(sv, sp, fbase + t);
                    }
                    __syncwarp();
                }
                bar_compute();
#endif
                wave_load<X3, CH, WG - 1, true>(slot, rank, hv, fbase + t);
                WTL(t, 6);
                wave_store<X3, CH, WG - 1, true>(slot, rank, fbase + t, hv, hbt);
                WTL(t, 7);
                fence_proxy_async();         // also covers this thread's own-slice stores of the previous cell phase
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&hb_full[t & 1]);
#if !WAVE_SAVE_FIRST
                store_saved(t - 1);          // HBM stores of the previous step ride in the shadow of this step's MMAs
#endif
            }
#endif
#if !WAVE_P_FIRST
            request_p();                     // this step's input projection: layer 0 from the GEMM-produced P0, layer 1 from layer 0's LL words
#endif
            WTL(t, 1);
            if (real) {
                float acc[CPW];
                if (t > 0) {
                    mbar_wait(rec_done, (t - 1) & 1);
                    WTL(t, 2);
                    tc_fence_after();
                    tmem_ld_nb<CPW>(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * CPW), acc);
#pragma unroll
                    for (int k = 1; k < NACC; ++k) {
                        float part[CPW];
                        tmem_ld_nb<CPW>(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(k * NB + cg * CPW), part);
#pragma unroll
                        for (int b = 0; b < CPW; ++b) acc[b] += part[b];
                    }
                } else {
#pragma unroll
                    for (int b = 0; b < CPW; ++b) acc[b] = 0.f;
                }
                WTL(t, 3);
                if (layer == 1) {
                    wait_ll_all<CPW / 2, WAVE_POLL_SEQ_FWD>(pl, [&](int j) { return p1x + (size_t)t * (WG * WPSLICE) + llw + j * 128; }, fbase + t + 1);
#pragma unroll
                    for (int j = 0; j < CPW / 2; ++j) {
                        pv[2 * j] = __uint_as_float(pl[j].x);
                        pv[2 * j + 1] = __uint_as_float(pl[j].z);
                    }
                }
#pragma unroll
                for (int b = 0; b < CPW; ++b) {
                    const float pre = acc[b] + pv[b] + qv[b];
                    aval[b] = (q == 2) ? tanhf_fast(pre) : sigmoidf_fast(pre);
                    gates[q][cg * CPW + b][lane] = aval[b];
                }
                tc_fence_before();
                bar_compute();
                WTL(t, 4);
                // cell update: thread = (unit = lane, rows warp*RPT ..); h_t goes into the OTHER operand buffer
                uint8_t* hbn = smem + S::H_OFF + ((t + 1) & 1) * S::H_BUF;
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    const int b = warp * RPT + i;
                    const float ig = gates[0][b][lane], fg = gates[1][b][lane], gg = gates[2][b][lane], og = gates[3][b][lane];
                    const float c = fmaf(fg, creg[i], ig * gg);
                    creg[i] = c;
                    const float h = og * tanhf_fast(c);
                    hreg[i] = h;
                    const __nv_bfloat16 hh = __float2bfloat16_rn(h);
                    *reinterpret_cast<__nv_bfloat16*>(hbn + hoff_k + b * 16) = hh;
                    if (X3)
                        *reinterpret_cast<__nv_bfloat16*>(hbn + S::H_PART + hoff_k + b * 16) = __float2bfloat16_rn(h - __bfloat162float(hh));
                }
#if WAVE_TMA
                // own slice of h_t is in operand buffer (t+1)&1: visible to the async proxy, then counted on its barrier
                fence_proxy_async();
                bar_compute();
                WTL(t, 5);
                if (t + 1 < nsteps) {
                    if (lane == 0) mbar_arrive(&hb_full[(t + 1) & 1]);
                    // publish: payload = the slice exactly as it sits in the operand buffer (2 KB per bf16 part), plain
                    // coalesced 8-byte stores; then ONE release store of the flag (cumulative over the CTA barrier)
                    if (tid < NW) {
                        const int part = tid >> 8, rem = tid & 255;
                        const uint2 d = *reinterpret_cast<const uint2*>(hbn + part * S::H_PART + (uint32_t)(rank * 4) * H_LBO + rem * 8);
                        uint2* dst = reinterpret_cast<uint2*>(own + (size_t)t * (WG * WSLICE) + rank * WSLICE) + tid;
                        asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(d.x), "r"(d.y) : "memory");
                    }
                    WTL(t, 8);
                    bar_compute();
                    WTL(t, 9);
                    if (tid == 0)
                        st_release_gpu(reinterpret_cast<uint32_t*>(own + (size_t)t * (WG * WSLICE) + rank * WSLICE + 256), fbase + t + 1);
                    WTL(t, 10);
                }
            }
#else
                bar_compute();
                WTL(t, 5);
                // publish this CTA's slice of h_t: peers need it for step t+1 (and the P1 duty for its extra step)
                if ((t + 1 < T || p1_duty) && tid < NW) {
                    const int part = tid >> 8, rem = tid & 255, chunk = rem >> 1, half = rem & 1;
                    const int kcl = chunk / NB, row = chunk % NB;
                    const uint2 d = *reinterpret_cast<const uint2*>(hbn + part * S::H_PART + (uint32_t)(rank * 4 + kcl) * H_LBO + row * 16 + half * 8);
                    st_ll(own + (size_t)t * (WG * WSLICE) + rank * WSLICE + tid, d.x, d.y, fbase + t + 1);
                }
            }
#endif
            if (p1_duty && t > 0) {
                // P1[t-1] = W_ih1_slice * h0_{t-1}^T has been accumulating behind this step's recurrent MMAs
                mbar_wait(p1_done, (t - 1) & 1);
                tc_fence_after();
#if WAVE_TMA
                if (tid == 0) *p1_safe = (uint32_t)t;      // the projection MMAs of step t (operand buffer t&1) are complete
#endif
                float pa[CPW];
                tmem_ld_nb<CPW>(tmem_p + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * CPW), pa);
#pragma unroll
                for (int k = 1; k < NACC; ++k) {
                    float part[CPW];
                    tmem_ld_nb<CPW>(tmem_p + ((uint32_t)(q * 32) << 16) + (uint32_t)(k * NB + cg * CPW), part);
#pragma unroll
                    for (int b = 0; b < CPW; ++b) pa[b] += part[b];
                }
#pragma unroll
                for (int j = 0; j < CPW / 2; ++j)
                    st_ll(p1x + (size_t)(t - 1) * (WG * WPSLICE) + llw + j * 128, __float_as_uint(pa[2 * j]),
                          __float_as_uint(pa[2 * j + 1]), fbase + t);
            }
        }
        if (!p1_duty) store_saved(T - 1);
        WTL(2, 15);
    }
    if (tid == 0) *cnt = (fbase >> 6) + 1;
    tc_fence_before();
    __syncthreads();
    if (warp == NT / 32) tmem_dealloc<TCOLS>(tmem_base);
}

// ================================================================================================
// Two-chain forward wavefront (WAVE_FWD2).  Per step the single-chain kernel above spends ~40 % of its period waiting
// for the exchange (publish -> L2 -> the peers' polls -> MMAs): pure latency, with every unit of the SM idle.  Here
// the group's 32 batch rows are run as two INDEPENDENT 16-row recurrences ("chains": warps 0-7 own rows 0-15, warps
// 8-15 rows 16-31 -- the gate phase and the cell phase already partition the rows this way), each with its own
// barriers, accumulators (tcgen05.mma N = 16) and LL words, so one chain computes while the other one's h_t is in
// flight.  Same bytes, same MMAs, same arithmetic as the single-chain kernel (bit-identical results); the exchange
// buffer layout is unchanged (a chain publishes / pulls the words of its rows).  The layer-1 input projection stays
// ONE N = 32 product per step (the SS-mode A operand read would otherwise double): it is issued once both chains
// have gathered h_{t-1}, in four chunks between which the issuer keeps serving recurrent products, accumulates into a
// double-buffered TMEM tile and is read out one step later (top of the next iteration, in the shadow of the exchange).
// MEASURED (B200, T=20, B=256, bf16x3): 131 us per launch vs 105 us for the single-chain kernel, with or without quiet
// waiting (WAVE_FWD2_QUIET); per-chain period 11.0 k cycles vs 8.3 k.  The exchange wait is therefore NOT idle latency
// that a second chain can fill: each chain still waits for the slowest of 8 CTAs, and that CTA is now also busy with its
// other chain.  Kept as a compile-time experiment (tools/build_variant.sh fwd2 -DWAVE_FWD2=1); not used by default.
// ================================================================================================
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bar_chain(int ch) { asm volatile("bar.sync %0, 256;" ::"r"(1 + ch) : "memory"); }

template <bool X3>
__global__ void __launch_bounds__(WNTA, 1) lstm_wave_fwd2_kernel(const __grid_constant__ WaveFwdArgs a) {
    using S = WaveFwdSmem<X3, WH>;
    constexpr int NB = WNB, NT = WNT, CH = WH, UC = WU, NBC = WNB / 2;
    constexpr int NWC = (X3 ? 2 : 1) * 128;    // LL words of one slice that belong to one chain
    constexpr int CPW = 8, RPT = 2, H4 = 4 * CH;
    extern __shared__ __align__(1024) uint8_t smem[];
    float (*gates)[NB][UC + 1] = reinterpret_cast<float (*)[NB][UC + 1]>(smem + S::G_OFF);
    uint64_t* hb_full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);   // [chain][buffer]: the chain's 8 warps -> issuer
    uint64_t* rec_done = hb_full + 4;                                     // [chain]
    uint64_t* p1_done = hb_full + 6;                                      // [buffer]
    uint64_t* w_full = hb_full + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(hb_full + 9);
    uint32_t* epoch_slot = tmem_slot + 1;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, cg = (warp >> 2) & 3;
    const int ch = (warp >> 3) & 1, ct = tid & 255;          // chain, thread index inside the chain (compute warps)
    const int layer = blockIdx.x / (a.G * WG);
    const int grp = (blockIdx.x / WG) % a.G;
    const int rank = blockIdx.x % WG;
    const int T = a.T, B = a.B;
    const int b0 = a.b_off + grp * NB;
    const bool p1_duty = (layer == 0 && a.L == 2);
    const int nsteps = p1_duty ? T + 1 : T;
    const float* P = layer ? nullptr : a.P0;
    const float* Q = layer ? nullptr : a.Q0;
    const float* W_hh = layer ? a.Whh1 : a.Whh0;
    float* h_all = layer ? a.h1 : a.h0;
    __nv_bfloat16* h_pl = layer ? a.hp1 : a.hp0;
    float* c_all = layer ? a.c1 : a.c0;
    float* acts = layer ? a.a1 : a.a0;
    uint32_t* cnt = reinterpret_cast<uint32_t*>(a.xchg) + (((a.L - 1) * 2 + layer) * WMAXG + grp) * WG + rank;
    uint4* own = a.xchg + WHDR + ((size_t)(layer * a.Gs + grp) * T) * (WG * WSLICE);
    uint4* p1x = a.xchg + WHDR + ((size_t)(2 * a.Gs) * T) * (WG * WSLICE) + ((size_t)grp * T) * (WG * WPSLICE) + rank * WPSLICE;

    constexpr int NACC = 2;
    constexpr int WCOLS = CH / 2;
    constexpr int ACOL = (X3 ? 2 : 1) * WCOLS;
    constexpr int TCOLS = 512;                 // W (256 / 128) + recurrent acc 2 chains x 2 x 16 + projection acc 2 x 2 x 32
    if (warp == NT / 32) tmem_alloc<TCOLS>(tmem_slot);
    if (tid == 32) {
        for (int i = 0; i < 4; ++i) mbar_init(&hb_full[i], NT / 64);      // 8 warps per chain
        mbar_init(&rec_done[0], 1); mbar_init(&rec_done[1], 1);
        mbar_init(&p1_done[0], 1); mbar_init(&p1_done[1], 1); mbar_init(w_full, 1);
        fence_mbar_init();
        if (p1_duty && a.packed)
            wave_load_S_image(a.packed + WavePack<X3>::fwdS(a.L, rank), smem_u32(smem + S::W_OFF), S::W_BYTES, w_full);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                      // (experimental kernel: no prologue overlap)
    const uint32_t fbase = *reinterpret_cast<volatile uint32_t*>(cnt) << 6;
    constexpr uint32_t W_LBO = WNC * 16, H_LBO = NB * 16, SBO_ = 128;
    const uint32_t tmem_d = tmem_base + ACOL;                  // recurrent accumulators [chain][NACC][NBC]
    const uint32_t tmem_p = tmem_d + 2 * NACC * NBC;           // projection accumulators [buffer][NACC][NB]
    constexpr uint32_t idesc16 = make_idesc_bf16(WNC, NBC), idesc32 = make_idesc_bf16(WNC, NB);
    const uint32_t hb_u = smem_u32(smem + S::H_OFF), wih_u = smem_u32(smem + S::W_OFF);

    if (warp >= NT / 32) {
        // ================= tcgen05 issuer: serves whichever chain has its operand ready, projection chunks in between
        reg_dec<32>();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (warp == NT / 32 && elect_one()) {
            if (p1_duty && a.packed) mbar_wait(w_full, 0);
            int ts0 = 1, ts1 = 1, p1_t = 1, p1_s = 0;
            uint32_t spins = 0;
            while (ts0 < nsteps || ts1 < nsteps || (p1_duty && p1_t < nsteps)) {
                bool progressed = false;
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    const int t = c ? ts1 : ts0;
                    if (t >= nsteps || !mbar_test(&hb_full[c * 2 + (t & 1)], ((t - 1) >> 1) & 1)) continue;
                    tc_fence_after();
                    if (t < T) {
                        const uint32_t hbt = hb_u + (t & 1) * S::H_BUF + (uint32_t)c * (NBC * 16);
                        const uint64_t dhh0 = make_smem_desc(hbt, H_LBO, SBO_);
                        const uint64_t dhl0 = make_smem_desc(hbt + S::H_PART, H_LBO, SBO_);
#pragma unroll 1
                        for (int j = 0; j < WG; ++j) {
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks) {
                                const int s = 2 * j + ks;
                                const uint64_t ih = (uint64_t)((s * 2 * H_LBO) >> 4);
                                const uint32_t awh = tmem_base + (uint32_t)(s * 8);
                                const uint32_t td = tmem_d + (uint32_t)((c * NACC + ks) * NBC);
                                const uint32_t first = j > 0 ? 1u : 0u;
                                if (X3) {
                                    umma_bf16_ts(td, awh + WCOLS, dhh0 + ih, idesc16, first);
                                    umma_bf16_ts(td, awh, dhl0 + ih, idesc16, 1u);
                                    umma_bf16_ts(td, awh, dhh0 + ih, idesc16, 1u);
                                } else {
                                    umma_bf16_ts(td, awh, dhh0 + ih, idesc16, first);
                                }
                            }
                        }
                        umma_commit(&rec_done[c]);
                    }
                    if (c) ++ts1; else ++ts0;
                    progressed = true;
                }
                if (p1_duty && p1_t < nsteps && p1_t < ts0 && p1_t < ts1) {
                    // a quarter of P1(p1_t) = W_ih1_slice * h0_{p1_t - 1}^T over all 32 rows (both chains gathered it)
                    const uint32_t hbt = hb_u + (p1_t & 1) * S::H_BUF;
                    const uint32_t koff = (uint32_t)p1_s * 2 * H_LBO, woff = (uint32_t)p1_s * 2 * W_LBO;
                    uint64_t dh = make_smem_desc(hbt + koff, H_LBO, SBO_), dl = make_smem_desc(hbt + S::H_PART + koff, H_LBO, SBO_);
                    uint64_t dwh = make_smem_desc(wih_u + woff, W_LBO, SBO_), dwl = make_smem_desc(wih_u + S::W_PART + woff, W_LBO, SBO_);
#pragma unroll 1
                    for (int i = 0; i < 4; ++i) {
                        const int s = p1_s + i;
                        const uint32_t td = tmem_p + (uint32_t)(((p1_t & 1) * NACC + (s % NACC)) * NB);
                        const uint32_t first = s >= NACC ? 1u : 0u;
                        if (X3) {
                            umma_bf16(td, dwl, dh, idesc32, first);
                            umma_bf16(td, dwh, dl, idesc32, 1u);
                            umma_bf16(td, dwh, dh, idesc32, 1u);
                        } else {
                            umma_bf16(td, dwh, dh, idesc32, first);
                        }
                        dh += (uint64_t)((2 * H_LBO) >> 4); dl += (uint64_t)((2 * H_LBO) >> 4);
                        dwh += (uint64_t)((2 * W_LBO) >> 4); dwl += (uint64_t)((2 * W_LBO) >> 4);
                    }
                    p1_s += 4;
                    if (p1_s == CH / 16) {
                        umma_commit(&p1_done[p1_t & 1]);
                        p1_s = 0;
                        ++p1_t;
                    }
                    progressed = true;
                }
                if (progressed) spins = 0;
                else if (++spins > FHVAE_SPIN_LIMIT * 16u) __trap();
            }
        }
        __syncwarp();
    } else {
        // ================= compute warps =================
        reg_inc<112>();
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * 32);
        if (a.packed) {
            wave_load_T_image<X3, 2>(a.packed + WavePack<X3>::fwdT(a.L, layer, rank), warp, lane, ta, 16, WCOLS);
        } else {
            const float* src = W_hh + (size_t)(q * CH + rank * UC + lane) * CH + cg * 64;
            float4 wv[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) wv[i] = __ldcg(reinterpret_cast<const float4*>(src) + i);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 w4 = wv[hf * 8 + i];
                    const float v[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        hi[2 * i + j] = pack_bf16(v[2 * j], v[2 * j + 1]);
                        lo[2 * i + j] = pack_bf16(v[2 * j] - __uint_as_float(hi[2 * i + j] << 16),
                                                  v[2 * j + 1] - __uint_as_float(hi[2 * i + j] & 0xffff0000u));
                    }
                }
                tmem_st16(ta + hf * 16, hi);
                if (X3) tmem_st16(ta + WCOLS + hf * 16, lo);
            }
        }
        tmem_wait_st();
        if (p1_duty && !a.packed) {
            const float* s1 = a.Wih1 + (size_t)(q * CH + rank * UC + lane) * CH + cg * 64;
            const int r = q * 32 + lane;
#pragma unroll 2
            for (int i = 0; i < 8; ++i) {
                const float4 x0 = __ldcg(reinterpret_cast<const float4*>(s1) + 2 * i);
                const float4 x1 = __ldcg(reinterpret_cast<const float4*>(s1) + 2 * i + 1);
                const float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                const uint32_t off = (uint32_t)(cg * 8 + i) * W_LBO + (uint32_t)r * 16;
                if (X3) {
                    uint4 hi, lo;
                    split_bf16(v, hi, lo);
                    *reinterpret_cast<uint4*>(smem + S::W_OFF + off) = hi;
                    *reinterpret_cast<uint4*>(smem + S::W_OFF + S::W_PART + off) = lo;
                } else {
                    *reinterpret_cast<uint4*>(smem + S::W_OFF + off) =
                        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                }
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        const int col = q * CH + rank * UC + lane;
        float qv[CPW];
        {
            const float bias = (layer == 1 && a.b1) ? __ldcg(a.b1 + col) : 0.f;
#pragma unroll
            for (int b = 0; b < CPW; ++b) qv[b] = bias + (Q ? __ldcg(Q + (size_t)(b0 + cg * CPW + b) * H4 + col) : 0.f);
        }
        float creg[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) creg[i] = 0.f;
        const int kglob = rank * UC + lane;
        const uint32_t hoff_k = (uint32_t)(kglob >> 3) * H_LBO + (uint32_t)(kglob & 7) * 2;
        const int llw = (cg * 4) * 128 + q * 32 + lane;
        float aval[CPW], hreg[RPT];
        // this thread's LL word of a slice: (bf16 part, K chunk, row of ITS chain, 8-byte half)
        const int xw_part = ct >> 7, xw_kcl = ((ct & 127) >> 1) >> 4, xw_row = ch * NBC + (((ct & 127) >> 1) & 15), xw_half = ct & 1;
        const int xw_word = xw_part * 256 + ((xw_kcl * NB + xw_row) << 1) + xw_half;
        const uint32_t xw_off = (uint32_t)xw_part * S::H_PART + (uint32_t)xw_kcl * H_LBO + (uint32_t)xw_row * 16 + xw_half * 8;
        const bool xw_on = ct < NWC;
        const uint32_t acc_addr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * NACC * NBC + (cg & 1) * CPW);
        auto store_saved = [&](int ts) {
            float* At = acts + ((size_t)ts * B + b0 + cg * CPW) * H4 + col;
#pragma unroll
            for (int b = 0; b < CPW; ++b) At[(size_t)b * H4] = aval[b];
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const size_t o = ((size_t)ts * B + b0 + warp * RPT + i) * CH + kglob;
                h_all[o] = hreg[i];
                c_all[o] = creg[i];
                if (h_pl) {
                    const __nv_bfloat16 hh = __float2bfloat16_rn(hreg[i]);
                    h_pl[o] = hh;
                    h_pl[o + a.hps] = __float2bfloat16_rn(hreg[i] - __bfloat162float(hh));
                }
            }
        };
        // P1(tau) = W_ih1_slice * h0_{tau-1}^T: TMEM tile tau&1 -> LL words of step tau-1 for the layer-1 CTA of this rank
        auto p1_readout = [&](int tau) {
            mbar_wait(&p1_done[tau & 1], ((tau - 1) >> 1) & 1);
            tc_fence_after();
            const uint32_t pb = tmem_p + ((uint32_t)(q * 32) << 16) + (uint32_t)((tau & 1) * NACC * NB + cg * CPW);
            float pa[CPW];
            tmem_ld_nb<CPW>(pb, pa);
#pragma unroll
            for (int k = 1; k < NACC; ++k) {
                float part[CPW];
                tmem_ld_nb<CPW>(pb + (uint32_t)(k * NB), part);
#pragma unroll
                for (int b = 0; b < CPW; ++b) pa[b] += part[b];
            }
#pragma unroll
            for (int j = 0; j < CPW / 2; ++j)
                st_ll(p1x + (size_t)(tau - 1) * (WG * WPSLICE) + llw + j * 128, __float_as_uint(pa[2 * j]),
                      __float_as_uint(pa[2 * j + 1]), fbase + tau);
        };

        for (int t = 0; t < nsteps; ++t) {
            const bool real = t < T;
            WTL(t, 0);
            uint8_t* hbt = smem + S::H_OFF + (t & 1) * S::H_BUF;
            // the projection issued behind the PREVIOUS step: complete before this iteration overwrites the operand rows
            // it read (own slice, cell phase below) -- and read out here, in the shadow of the exchange
            if (p1_duty && t >= 2) p1_readout(t - 1);
            if (t > 0) {
                const uint4* slot = own + (size_t)(t - 1) * (WG * WSLICE);
#if WAVE_FWD2_QUIET
                // quiet wait: ONE lane per peer polls a sentinel word of this chain's rows, the chain's other warps sleep
                // at the named barrier (no poll traffic competing with the chain that is computing)
                if ((warp & 7) == 0) {
                    if (lane < WG - 1) {
                        const int lastw = (X3 ? 256 : 0) + ((3 * NB + ch * NBC + NBC - 1) << 1) + 1;
                        const uint4* sp = slot + (lane + (lane >= rank ? 1 : 0)) * WSLICE + lastw;
                        uint4 sv = ld_ll(sp);
                        wait_ll(sv, sp, fbase + t);
                    }
                    __syncwarp();
                }
                bar_chain(ch);
#endif
                if (xw_on) {
                    uint4 hv[7];
#pragma unroll
                    for (int i = 0; i < 7; ++i) hv[i] = ld_ll(slot + (i + (i >= rank ? 1 : 0)) * WSLICE + xw_word);
                    wait_ll_all<7, 1>(hv, [&](int i) { return slot + (i + (i >= rank ? 1 : 0)) * WSLICE + xw_word; }, fbase + t);
#pragma unroll
                    for (int i = 0; i < 7; ++i) {
                        const int src = i + (i >= rank ? 1 : 0);
                        *reinterpret_cast<uint2*>(hbt + xw_off + (uint32_t)(src * 4) * H_LBO) = make_uint2(hv[i].x, hv[i].z);
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&hb_full[ch * 2 + (t & 1)]);
                store_saved(t - 1);
            }
            float pv[CPW];
            uint4 pl[CPW / 2];
            if (real) {
                if (layer == 0) {
                    const float* Pt = P ? P + ((size_t)t * B + b0 + cg * CPW) * H4 + col : nullptr;
#pragma unroll
                    for (int b = 0; b < CPW; ++b) pv[b] = Pt ? __ldcg(Pt + (size_t)b * H4) : 0.f;
                } else {
#pragma unroll
                    for (int j = 0; j < CPW / 2; ++j) pl[j] = ld_ll(p1x + (size_t)t * (WG * WPSLICE) + llw + j * 128);
                }
                float acc[CPW];
                if (t > 0) {
                    mbar_wait(&rec_done[ch], (t - 1) & 1);
                    WTL(t, 2);
                    tc_fence_after();
                    tmem_ld_nb<CPW>(acc_addr, acc);
#pragma unroll
                    for (int k = 1; k < NACC; ++k) {
                        float part[CPW];
                        tmem_ld_nb<CPW>(acc_addr + (uint32_t)(k * NBC), part);
#pragma unroll
                        for (int b = 0; b < CPW; ++b) acc[b] += part[b];
                    }
                } else {
#pragma unroll
                    for (int b = 0; b < CPW; ++b) acc[b] = 0.f;
                }
                if (layer == 1) {
                    wait_ll_all<CPW / 2, 1>(pl, [&](int j) { return p1x + (size_t)t * (WG * WPSLICE) + llw + j * 128; }, fbase + t + 1);
#pragma unroll
                    for (int j = 0; j < CPW / 2; ++j) {
                        pv[2 * j] = __uint_as_float(pl[j].x);
                        pv[2 * j + 1] = __uint_as_float(pl[j].z);
                    }
                }
#pragma unroll
                for (int b = 0; b < CPW; ++b) {
                    const float pre = acc[b] + pv[b] + qv[b];
                    aval[b] = (q == 2) ? tanhf_fast(pre) : sigmoidf_fast(pre);
                    gates[q][cg * CPW + b][lane] = aval[b];
                }
                tc_fence_before();
                bar_chain(ch);
                WTL(t, 4);
                uint8_t* hbn = smem + S::H_OFF + ((t + 1) & 1) * S::H_BUF;
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    const int b = warp * RPT + i;
                    const float ig = gates[0][b][lane], fg = gates[1][b][lane], gg = gates[2][b][lane], og = gates[3][b][lane];
                    const float c = fmaf(fg, creg[i], ig * gg);
                    creg[i] = c;
                    const float h = og * tanhf_fast(c);
                    hreg[i] = h;
                    const __nv_bfloat16 hh = __float2bfloat16_rn(h);
                    *reinterpret_cast<__nv_bfloat16*>(hbn + hoff_k + b * 16) = hh;
                    if (X3)
                        *reinterpret_cast<__nv_bfloat16*>(hbn + S::H_PART + hoff_k + b * 16) = __float2bfloat16_rn(h - __bfloat162float(hh));
                }
                bar_chain(ch);
                WTL(t, 5);
                if (t + 1 < nsteps && xw_on) {
                    const uint2 d = *reinterpret_cast<const uint2*>(hbn + xw_off + (uint32_t)(rank * 4) * H_LBO);
                    st_ll(own + (size_t)t * (WG * WSLICE) + rank * WSLICE + xw_word, d.x, d.y, fbase + t + 1);
                }
            }
        }
        if (p1_duty) p1_readout(nsteps - 1);
        else store_saved(T - 1);
    }
    if (tid == 0) *cnt = (fbase >> 6) + 1;
    tc_fence_before();
    __syncthreads();
    if (warp == NT / 32) tmem_dealloc<TCOLS>(tmem_base);
}

// ================================================================================================
// BPTT wavefront (top layer first)
// ================================================================================================
// Per step t (descending) the CTA that owns 32 units computes dgates_t pointwise and contributes the split-K
// partial  W_hh^T[:, its 128 gate columns] * dgates_t  of dh_{t-1} for ALL 256 units (TMEM lane = unit); the
// 8 partial tiles of a unit slice are exchanged as LL words (reduce-scatter through L2, 2-deep) and summed in
// a fixed order by the owner.  In a 2-layer stack the layer-1 CTA of rank j also publishes its dgates1_t slice
// (bf16 hi/lo operand words); the layer-0 CTA of rank j multiplies it by ITS W_ih1^T slice (resident smem A
// operand) into the same accumulator that later receives W_hh0^T * dgates0_{t+1}:
//      dh0_t = dgates0_{t+1} W_hh0 + dgates1_t W_ih1          -> ONE reduce-scatter, no dgrad GEMM, no dh buffer.
// The cross-layer MMAs are issued two steps ahead (accumulators double-buffered in TMEM), off the critical path.
struct WaveBwdArgs {
    const float* dh_all;      // top layer: dL/dh_t from the consumer of all outputs (may be NULL)
    const float* dh_last1;    // top layer: extra dL/dh_{T-1} (may be NULL)
    const float* dh_last0;    // bottom layer of a 2-layer stack: extra dL/dh0_{T-1} (may be NULL)
    const float* Whh1; const float* c1; const float* a1; float* dg1; float* dgsum1;   // top layer (the only one if L == 1)
    const float* Wih1;                                                                  // (4H, H) of the top layer
    const float* Whh0; const float* c0; const float* a0; float* dg0; float* dgsum0;   // bottom layer
    uint4* xchg;
    int T, B, b_off, G, Gs, L;
    __nv_bfloat16* dgp1; __nv_bfloat16* dgp0; long long dgps;   // optional bf16 hi/lo planes of dgates (lo at +dgps)
    const uint8_t* packed;          // optional pre-packed operand images of the weights (wave_pack_kernel)
};

template <bool X3, int CH = WH>
struct WaveBwdSmem {
    static constexpr int G_PART = WNB * WNC * 2;               // dgates operand, 32 rows x 128 gate cols bf16 = 8 KB
    static constexpr int G_BUF = (X3 ? 2 : 1) * G_PART;
    static constexpr int G0_OFF = 0;                           // own dgates_t (B operand of the recurrent product)
    static constexpr int G1_OFF = G_BUF;                       // [2] dgates1 of the layer above (bottom layer only)
    static constexpr int BAR_OFF = 3 * G_BUF;
    static constexpr int W_OFF = (BAR_OFF + 64 + 1023) / 1024 * 1024;
    static constexpr int W_HALF = 128 * WNC * 2;               // 128 units x 128 gate cols bf16 = 32 KB
    static constexpr int W_PART = (CH / 128) * W_HALF;
    static constexpr int W_BYTES = (X3 ? 2 : 1) * W_PART;
    static constexpr int TOTAL1 = W_OFF;
    static constexpr int TOTAL2 = W_OFF + W_BYTES;
};
constexpr int WRS = 512;           // uint4 words of one (dst, src) partial tile: 32 units x 16 row pairs
constexpr int WDG = 2048;          // uint4 words of one published dgates slice (2 parts x 16 chunks x 32 rows x 2 halves)

template <bool X3, int CH_>
__global__ void __launch_bounds__(WNTA, 1) lstm_wave_bwd_kernel(const __grid_constant__ WaveBwdArgs a) {
    using S = WaveBwdSmem<X3, CH_>;
    constexpr int NB = WNB, NT = WNT, CH = CH_, UC = WU, NC = WNC;
    constexpr int WG = CH / WU, HF = CH / 128;                 // CTAs per group (shadows the H = 256 constant), 128-unit halves
    constexpr int NCG = NT / 128, CPW = NB / NCG, RPT = NB * 32 / NT, H4 = 4 * CH;
    constexpr int NPARTW = (X3 ? 2 : 1) * 1024;               // LL words of a dgates slice actually used
    static_assert(CPW == 8 && RPT == 2, "LL packing assumes 8 columns per thread / 2 rows per thread");
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* g0 = smem + S::G0_OFF;
    uint64_t* g_full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);    // [2] compute warps -> issuer
    uint64_t* g_pro = g_full + 2;                                          // bottom-layer prologue
    uint64_t* rec_done = g_full + 3;                                       // issuer (commit) -> compute warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g_full + 4);
    uint32_t* epoch_slot = tmem_slot + 1;
    uint64_t* w_full = g_full + 6;                                         // packed W_ih1^T image landed in shared memory

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, cg = (warp >> 2) & 3;
    const int role = blockIdx.x / (a.G * WG);                  // 0: top layer (scheduled first), 1: bottom layer
    const int grp = (blockIdx.x / WG) % a.G;
    const int rank = blockIdx.x % WG;
    const int T = a.T, B = a.B;
    const int b0 = a.b_off + grp * NB;
    const bool bottom = role == 1;                             // bottom layer of a 2-layer stack
    const bool top2 = role == 0 && a.L == 2;                   // top layer that feeds a bottom layer
    const float* W_hh = bottom ? a.Whh0 : a.Whh1;
    const float* c_all = bottom ? a.c0 : a.c1;
    const float* acts = bottom ? a.a0 : a.a1;
    float* dgates = bottom ? a.dg0 : a.dg1;
    __nv_bfloat16* dg_pl = bottom ? a.dgp0 : a.dgp1;
    float* dgsum = bottom ? a.dgsum0 : a.dgsum1;
    const float* dh_all = bottom ? nullptr : a.dh_all;
    const float* dh_last = bottom ? a.dh_last0 : a.dh_last1;
    // exchange buffer: [header][partials: role][group][2][dst][src][WRS] then [dgates1: group][t][rank][WDG]
    uint32_t* cnt = reinterpret_cast<uint32_t*>(a.xchg) + (((a.L - 1) * 2 + role) * WMAXG + grp) * WG + rank;
    uint4* rs = a.xchg + WHDR + ((size_t)(role * a.Gs + grp) * 2) * (WG * WG * WRS);
    uint4* dgx = a.xchg + WHDR + ((size_t)(2 * a.Gs) * 2) * (WG * WG * WRS) + ((size_t)grp * T) * (WG * WDG) + rank * WDG;

#if WAVE_PDL_EARLY
    pdl_launch_dependents();
#endif
    constexpr int NACC = 2;
    constexpr int ABUF = HF * NACC * NB;                       // one accumulator set: HF unit halves x NACC x NB columns
    constexpr int WCOLS = NC / 2;                              // 64 columns per (half, part) of the resident W_hh^T
    constexpr int ACOL = (X3 ? 2 : 1) * HF * WCOLS;            // 256 / 128 (H = 256), 128 / 64 (H = 128)
    constexpr int TCOLS = (ACOL + 2 * ABUF) <= 256 ? 256 : 512;
    if (warp == NT / 32) tmem_alloc<TCOLS>(tmem_slot);
    if (tid == 32) {
        mbar_init(&g_full[0], NT / 32); mbar_init(&g_full[1], NT / 32); mbar_init(g_pro, NT / 32);
        mbar_init(rec_done, 1); mbar_init(w_full, 1);
        fence_mbar_init();
        if (bottom && a.packed)
            wave_load_S_image(a.packed + WavePack<X3, CH>::bwdS(a.L, rank), smem_u32(smem + S::W_OFF), S::W_BYTES, w_full);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    uint32_t fbase = 0;              // flag base = launch epoch << 6: read by the compute warps after pdl_wait()
    const uint32_t tmem_acc = tmem_base + ACOL;
    constexpr uint32_t idesc = make_idesc_bf16(128, NB);
    constexpr uint32_t W_LBO = 128 * 16, G_LBO = NB * 16, SBO_ = 128;
    const uint32_t g0_u = smem_u32(g0), g1_u = smem_u32(smem + S::G1_OFF), w_u = smem_u32(smem + S::W_OFF);

    if (warp >= NT / 32) {
        // ================= tcgen05 issuer =================
        reg_dec<32>();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (warp == NT / 32 && elect_one()) {
            // acc set `as` (+)= W^T * G : A from TMEM (recurrent, W_hh^T) or from smem (cross-layer, W_ih1^T)
            auto rec = [&](uint32_t as, uint32_t accumulate) {
                const uint64_t dgh0 = make_smem_desc(g0_u, G_LBO, SBO_), dgl0 = make_smem_desc(g0_u + S::G_PART, G_LBO, SBO_);
#pragma unroll 1
                for (int hf = 0; hf < HF; ++hf) {
#pragma unroll 2
                    for (int s = 0; s < NC / 16; ++s) {
                        const uint64_t ig = (uint64_t)((s * 2 * G_LBO) >> 4);
                        const uint32_t awh = tmem_base + (uint32_t)(hf * WCOLS + s * 8), awl = awh + HF * WCOLS;
                        const uint32_t td = tmem_acc + as * ABUF + (uint32_t)((hf * NACC + (s % NACC)) * NB);
                        const uint32_t first = (accumulate || s >= NACC) ? 1u : 0u;
                        if (X3) {
                            umma_bf16_ts(td, awl, dgh0 + ig, idesc, first);
                            umma_bf16_ts(td, awh, dgl0 + ig, idesc, 1u);
                            umma_bf16_ts(td, awh, dgh0 + ig, idesc, 1u);
                        } else {
                            umma_bf16_ts(td, awh, dgh0 + ig, idesc, first);
                        }
                    }
                }
            };
            auto cross = [&](uint32_t as, uint32_t gbuf) {
                const uint32_t gb = g1_u + gbuf * S::G_BUF;
                const uint64_t dgh0 = make_smem_desc(gb, G_LBO, SBO_), dgl0 = make_smem_desc(gb + S::G_PART, G_LBO, SBO_);
                const uint64_t dwh0 = make_smem_desc(w_u, W_LBO, SBO_), dwl0 = make_smem_desc(w_u + S::W_PART, W_LBO, SBO_);
#pragma unroll 1
                for (int hf = 0; hf < HF; ++hf) {
#pragma unroll 2
                    for (int s = 0; s < NC / 16; ++s) {
                        const uint64_t iw = (uint64_t)((hf * S::W_HALF + s * 2 * W_LBO) >> 4), ig = (uint64_t)((s * 2 * G_LBO) >> 4);
                        const uint32_t td = tmem_acc + as * ABUF + (uint32_t)((hf * NACC + (s % NACC)) * NB);
                        const uint32_t first = s >= NACC ? 1u : 0u;
                        if (X3) {
                            umma_bf16(td, dwl0 + iw, dgh0 + ig, idesc, first);
                            umma_bf16(td, dwh0 + iw, dgl0 + ig, idesc, 1u);
                            umma_bf16(td, dwh0 + iw, dgh0 + ig, idesc, 1u);
                        } else {
                            umma_bf16(td, dwh0 + iw, dgh0 + ig, idesc, first);
                        }
                    }
                }
            };
            if (bottom) {
                if (a.packed) mbar_wait(w_full, 0);
                mbar_wait(g_pro, 0);                           // dgates1_{T-1} (and _{T-2}) are in G1
                tc_fence_after();
                cross((uint32_t)((T - 1) & 1), (uint32_t)((T - 1) & 1));
                umma_commit(rec_done);                         // dh0_{T-1} partial = cross part only
                if (T > 1) cross((uint32_t)((T - 2) & 1), (uint32_t)((T - 2) & 1));
            }
            for (int k = 0; k + 1 < T; ++k) {
                const int t = T - 1 - k;                       // dgates_t ready -> partial of dh_{t-1}
                mbar_wait(&g_full[k & 1], (k >> 1) & 1);
                tc_fence_after();
                rec((uint32_t)((t - 1) & 1), bottom ? 1u : 0u);
                umma_commit(rec_done);
                if (bottom && t >= 2) cross((uint32_t)(t & 1), (uint32_t)(t & 1));   // dh0_{t-2} += dgates1_{t-2} W_ih1
            }
        }
        __syncwarp();
    } else {
        // ================= compute warps =================
        reg_inc<112>();
        if (a.packed) {
            const int layer_of = bottom ? 0 : a.L - 1;
            wave_load_T_image<X3, HF>(a.packed + WavePack<X3, CH>::bwdT(a.L, layer_of, rank), warp, lane,
                                      tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * 16), WCOLS, HF * WCOLS);
            tmem_wait_st();
        } else {   // resident A operand of the recurrent product: lane = unit n (per half), column = (k, k+1) pair of the
            // CTA's 128 gate columns k = g*32 + u;  A[n][k] = W_hh[(g*H + 32*rank + u) * H + n];  cg <-> gate g
            const int g = cg;
#pragma unroll 1
            for (int hf = 0; hf < HF; ++hf) {
                const int n = hf * 128 + q * 32 + lane;
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float v0 = __ldcg(W_hh + (size_t)(g * CH + rank * UC + 2 * j) * CH + n);
                    const float v1 = __ldcg(W_hh + (size_t)(g * CH + rank * UC + 2 * j + 1) * CH + n);
                    hi[j] = pack_bf16(v0, v1);
                    lo[j] = pack_bf16(v0 - __uint_as_float(hi[j] << 16), v1 - __uint_as_float(hi[j] & 0xffff0000u));
                }
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * WCOLS + g * 16);
                tmem_st16(ta, hi);
                if (X3) tmem_st16(ta + HF * WCOLS, lo);
            }
            tmem_wait_st();
        }
        if (bottom && !a.packed) {
            // resident smem A operand of the cross-layer product: A2[n][k] = W_ih1[(g*H + 32*rank + u) * H + n]
            constexpr int WIT = CH * (NC / 8) / NT;
#pragma unroll 2
            for (int i = 0; i < WIT; ++i) {
                const int item = tid + i * NT;
                const int n = item % CH, kc = item / CH;
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int k = kc * 8 + j, g = k >> 5, u = k & 31;
                    v[j] = __ldcg(a.Wih1 + (size_t)(g * CH + rank * UC + u) * CH + n);
                }
                const uint32_t off = (uint32_t)(n >> 7) * S::W_HALF + (uint32_t)(kc * 128 + (n & 127)) * 16;
                if (X3) {
                    uint4 hi, lo;
                    split_bf16(v, hi, lo);
                    *reinterpret_cast<uint4*>(smem + S::W_OFF + off) = hi;
                    *reinterpret_cast<uint4*>(smem + S::W_OFF + S::W_PART + off) = lo;
                } else {
                    *reinterpret_cast<uint4*>(smem + S::W_OFF + off) =
                        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                }
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        pdl_wait();                      // (see the forward kernel: the prologue above overlaps the preceding launch)
        fbase = *reinterpret_cast<volatile uint32_t*>(cnt) << 6;

        // dgates1_t slice of the layer above -> G1[buf]: 2048 (1024) LL words, 4 (2) per thread
        auto pull_dg = [&](int t, int buf) {
            const uint4* src = dgx + (size_t)t * (WG * WDG);
            uint8_t* dst = smem + S::G1_OFF + buf * S::G_BUF;
            constexpr int PER = NPARTW / NT;
            uint4 v[PER];
#if WAVE_FLAGPOLL
            if (warp == 0) {
                if (lane == 0) {
                    uint4 sv = ld_ll(src + NPARTW - 1);
                    wait_ll(sv, src + NPARTW - 1, fbase + (uint32_t)(T - t));
                }
                __syncwarp();
            }
            bar_compute();
#endif
#pragma unroll
            for (int i = 0; i < PER; ++i) v[i] = ld_ll(src + tid + i * NT);
            wait_ll_all<PER, WAVE_POLL_SEQ_BWD>(v, [&](int i) { return src + tid + i * NT; }, fbase + (uint32_t)(T - t));
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int w = tid + i * NT;
                *reinterpret_cast<uint2*>(dst + (w >> 10) * S::G_PART + (w & 1023) * 8) = make_uint2(v[i].x, v[i].z);
            }
        };
        if (bottom) {
            pull_dg(T - 1, (T - 1) & 1);
            if (T > 1) pull_dg(T - 2, (T - 2) & 1);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(g_pro);
        }

        float dcreg[RPT], gsum[4][RPT], gkeep[4][RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            dcreg[i] = 0.f;
            gsum[0][i] = gsum[1][i] = gsum[2][i] = gsum[3][i] = 0.f;
        }
        const int ucol = rank * UC + lane;                     // this thread's hidden unit (pointwise phase)
        auto store_dg = [&](int ts) {
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const size_t o = ((size_t)ts * B + b0 + warp * RPT + i) * H4 + ucol;
                if (dgates) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) dgates[o + g * CH] = gkeep[g][i];
                }
                if (dg_pl) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const __nv_bfloat16 hh = __float2bfloat16_rn(gkeep[g][i]);
                        dg_pl[o + g * CH] = hh;
                        dg_pl[o + g * CH + a.dgps] = __float2bfloat16_rn(gkeep[g][i] - __bfloat162float(hh));
                    }
                }
            }
        };
        uint32_t nwait = 0;

        for (int k = 0; k < T; ++k) {
            const int t = T - 1 - k;
            // ---- everything the pointwise step needs (thread = unit `lane`, rows warp*RPT + i)
            float a_i[RPT], a_f[RPT], a_g[RPT], a_o[RPT], c_t[RPT], c_p[RPT], dh[RPT];
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const int b = b0 + warp * RPT + i;
                const size_t r4 = ((size_t)t * B + b) * H4, r1 = ((size_t)t * B + b) * CH;
                a_i[i] = __ldcg(acts + r4 + ucol);
                a_f[i] = __ldcg(acts + r4 + CH + ucol);
                a_g[i] = __ldcg(acts + r4 + 2 * CH + ucol);
                a_o[i] = __ldcg(acts + r4 + 3 * CH + ucol);
                c_t[i] = __ldcg(c_all + r1 + ucol);
                c_p[i] = t ? __ldcg(c_all + r1 - (size_t)B * CH + ucol) : 0.f;
                float d = dh_all ? __ldcg(dh_all + r1 + ucol) : 0.f;
                if (t == T - 1 && dh_last) d += __ldcg(dh_last + (size_t)b * CH + ucol);
                dh[i] = d;
            }
            if (k > 0 || bottom) {
                // ---- this CTA's split-K partial of dh_t for all 256 units: TMEM -> LL words, one tile per owner
                mbar_wait(rec_done, nwait & 1);
                ++nwait;
                tc_fence_after();
                uint4* wr = rs + (size_t)(k & 1) * (WG * WG * WRS);
#pragma unroll
                for (int hf = 0; hf < HF; ++hf) {
                    float pv[CPW];
                    const uint32_t tb = tmem_acc + (uint32_t)((t & 1) * ABUF + hf * NACC * NB) + ((uint32_t)(q * 32) << 16);
                    tmem_ld_nb<CPW>(tb + (uint32_t)(cg * CPW), pv);
#pragma unroll
                    for (int x = 1; x < NACC; ++x) {
                        float part[CPW];
                        tmem_ld_nb<CPW>(tb + (uint32_t)(x * NB + cg * CPW), part);
#pragma unroll
                        for (int b = 0; b < CPW; ++b) pv[b] += part[b];
                    }
                    const int dst = hf * 4 + q;
                    uint4* wp = wr + ((size_t)dst * WG + rank) * WRS + (cg * 4) * 32 + lane;
#pragma unroll
                    for (int j = 0; j < CPW / 2; ++j)
                        st_ll(wp + j * 32, __float_as_uint(pv[2 * j]), __float_as_uint(pv[2 * j + 1]), fbase + k + 1);
                }
                tc_fence_before();
                // ---- the 8 partial tiles of this CTA's units (fixed summation order: deterministic)
                const uint4* rd = wr + ((size_t)rank * WG) * WRS + warp * 32 + lane;
                uint4 pr[WG];
#if WAVE_FLAGPOLL
                if (warp == 0) {
                    if (lane < WG) {
                        const uint4* sp = wr + ((size_t)rank * WG + lane) * WRS + (WRS - 1);
                        uint4 sv = ld_ll(sp);
                        wait_ll(sv, sp, fbase + k + 1);
                    }
                    __syncwarp();
                }
                bar_compute();
#endif
#pragma unroll
                for (int src = 0; src < WG; ++src) pr[src] = ld_ll(rd + src * WRS);
                wait_ll_all<WG, WAVE_POLL_SEQ_BWD>(pr, [&](int src) { return rd + src * WRS; }, fbase + k + 1);
#pragma unroll
                for (int src = 0; src < WG; ++src) {                      // fixed summation order: deterministic
                    dh[0] += __uint_as_float(pr[src].x);
                    dh[1] += __uint_as_float(pr[src].z);
                }
            }
            // ---- pointwise BPTT; bf16 operand of the next products in smem
            uint8_t* g_hi = g0;
            uint8_t* g_lo = g0 + S::G_PART;
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const int bl = warp * RPT + i;
                const float tc = tanhf_fast(c_t[i]);
                const float dc = dcreg[i] + dh[i] * a_o[i] * (1.f - tc * tc);
                dcreg[i] = dc * a_f[i];
                float gq[4];
                gq[0] = dc * a_g[i] * a_i[i] * (1.f - a_i[i]);
                gq[1] = dc * c_p[i] * a_f[i] * (1.f - a_f[i]);
                gq[2] = dc * a_i[i] * (1.f - a_g[i] * a_g[i]);
                gq[3] = dh[i] * tc * a_o[i] * (1.f - a_o[i]);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    gkeep[g][i] = gq[g];
                    gsum[g][i] += gq[g];
                    const int kk = g * 32 + lane;
                    const uint32_t off = (uint32_t)(kk >> 3) * G_LBO + (uint32_t)bl * 16 + (uint32_t)(kk & 7) * 2;
                    const __nv_bfloat16 hh = __float2bfloat16_rn(gq[g]);
                    *reinterpret_cast<__nv_bfloat16*>(g_hi + off) = hh;
                    if (X3) *reinterpret_cast<__nv_bfloat16*>(g_lo + off) = __float2bfloat16_rn(gq[g] - __bfloat162float(hh));
                }
            }
            if (top2) {
                // publish this CTA's dgates_t slice (operand layout, 8-byte halves) for the bottom-layer CTA of the same rank
                bar_compute();
                uint4* dp = dgx + (size_t)t * (WG * WDG);
#pragma unroll
                for (int i = 0; i < NPARTW / NT; ++i) {
                    const int w = tid + i * NT;
                    const uint2 d = *reinterpret_cast<const uint2*>(g0 + (w >> 10) * S::G_PART + (w & 1023) * 8);
                    st_ll(dp + w, d.x, d.y, fbase + (uint32_t)(T - t));
                }
            }
            if (bottom && t >= 2) pull_dg(t - 2, t & 1);       // operand of the cross-layer MMAs issued behind this step's
            if (t > 0) {
                fence_proxy_async();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&g_full[k & 1]);
            }
            store_dg(t);                                        // HBM stores in the shadow of the MMAs
        }
        if (dgsum) {
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                float* o = dgsum + (size_t)(b0 + warp * RPT + i) * H4 + ucol;
#pragma unroll
                for (int g = 0; g < 4; ++g) o[g * CH] = gsum[g][i];
            }
        }
    }
    if (tid == 0) *cnt = (fbase >> 6) + 1;
    tc_fence_before();
    __syncthreads();
    if (warp == NT / 32) tmem_dealloc<TCOLS>(tmem_base);
}


// Forward progress of a launch needs ALL its CTAs co-resident (they spin on each other's words): the launch size is
// derived from the CURRENT device (SM count x resident CTAs per SM for this kernel's shared memory), not from a
// compile-time constant, and an over-sized grid is refused with an error instead of hanging (MIG slice, MPS SM
// limit, a smaller part).  With FHVAE_WAVE_COOPERATIVE=1 the launch is additionally cooperative: the driver
// gang-schedules the grid, i.e. it starts only when every CTA has an SM.  Default off: measured 0.949 vs 0.928 ms per
// step (a gang-scheduled launch cannot start while the previous stack's weight-gradient CTAs still drain), and without
// it forward progress only needs every OTHER resident kernel to terminate on its own, which holds for every kernel of
// this library (tests/test_gpu_train_step.py::test_wavefront_launch_survives_sm_hogging_neighbours); a protocol
// bug traps after FHVAE_SPIN_LIMIT polls instead of hanging the GPU.
static bool wave_cooperative() {
#if WAVE_COOPERATIVE >= 0
    return WAVE_COOPERATIVE != 0;
#else
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("FHVAE_WAVE_COOPERATIVE");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
#endif
}
static int device_index() {
    int d = 0;
    cudaGetDevice(&d);
    return d < 0 || d >= 64 ? 0 : d;
}
static int device_sm_count() {
    static int cache[64] = {0};
    const int d = device_index();
    if (cache[d] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || n <= 0) n = 0;
        cache[d] = n > 0 ? n : -1;
    }
    return cache[d] > 0 ? cache[d] : 0;
}
static int wave_groups_per_launch(int L, int H) {   // H = 256: 18 or 9 on a full B200 (148 SMs, 1 CTA of 181-214 KB per SM)
    const int g = device_sm_count() / ((H / WU) * L);
    return g > WMAXG ? WMAXG : g;
}
template <typename K>
static int wave_launch(K kern, const char* name, int grid, size_t smem, void* args, cudaStream_t st) {
    // (kernel, device) pairs whose shared-memory attribute / occupancy have been checked -- keyed by the kernel's
    // address: the bf16x3 and bf16 instantiations share one function-pointer TYPE, hence one copy of this template
    static const void* ready_k[16] = {nullptr};
    static int ready_d[16] = {0};
    static int n_ready = 0;
    const int d = device_index();
    bool ready = false;
    for (int i = 0; i < n_ready; ++i) ready = ready || (ready_k[i] == reinterpret_cast<const void*>(kern) && ready_d[i] == d);
    if (!ready) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int occ = 0;
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WNTA, smem);
        if (e != cudaSuccess || occ < 1) {
            set_error("%s: %d B of shared memory per CTA not available on this device (%s)", name, (int)smem,
                      e != cudaSuccess ? cudaGetErrorString(e) : "0 resident CTAs per SM");
            return e != cudaSuccess ? (int)e : FHVAE_ENOSUP;
        }
        if (n_ready < 16) {
            ready_k[n_ready] = reinterpret_cast<const void*>(kern);
            ready_d[n_ready++] = d;
        }
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(WNTA);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    if (wave_cooperative()) {
        at[0].id = cudaLaunchAttributeCooperative;
        at[0].val.cooperative = 1;
    } else {                                    // the prologue (barriers, TMEM, weight images) overlaps the launch before
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = pdl_enabled(PDL_WAVE) ? 1 : 0;
    }
    cfg.attrs = at;
    cfg.numAttrs = 1;
    void* kargs[1] = {args};
    cudaError_t e = cudaLaunchKernelExC(&cfg, reinterpret_cast<const void*>(kern), kargs);
    count_launches(1);
    if (e != cudaSuccess) {
        set_error("%s: launch of %d co-resident CTAs failed: %s", name, grid, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

bool lstm_wave_supported(int T, int B, int H, int L) {
    return (H == 256 || H == 128) && B % WNB == 0 && T >= 1 && T <= WMAXT && (L == 1 || L == 2) &&
           wave_groups_per_launch(L, H) >= 1;
}

int lstm_wave_rows_per_launch(int H, int L) {      // batch rows one launch serves with every group co-resident
    if (!(H == 256 || H == 128) || !(L == 1 || L == 2)) return 0;
    return wave_groups_per_launch(L, H) * WNB;
}

size_t lstm_wave_xchg_bytes(int T, int B, int H, int L) {
    int gs = B / WNB;
    const int gmax = wave_groups_per_launch(L, H), ng = H / WU;
    if (gs > gmax) gs = gmax;
    return ((size_t)WHDR + (size_t)gs * T * ng * (L == 2 ? 2 * WSLICE + WPSLICE : WSLICE)) * sizeof(uint4);
}

template <bool X3, int CH>
static int launch_wave_fwd(WaveFwdArgs a, cudaStream_t st) {
    using S = WaveFwdSmem<X3, CH>;
    constexpr int NG = CH / WU;
#if WAVE_FWD2
    static_assert(CH == WH || true, "");
    auto kern = CH == WH ? lstm_wave_fwd2_kernel<X3> : lstm_wave_fwd_kernel<X3, CH>;
#else
    auto kern = lstm_wave_fwd_kernel<X3, CH>;
#endif
    const int gtot = a.B / WNB, gmax = wave_groups_per_launch(a.L, CH);
    if (gmax < 1) {
        set_error("lstm_wave_fwd: the device has too few SMs for one group of %d CTAs x %d layers", NG, a.L);
        return FHVAE_ENOSUP;
    }
    a.Gs = gtot < gmax ? gtot : gmax;
    for (int g0 = 0; g0 < gtot; g0 += gmax) {          // all CTAs of a launch must be co-resident (one per SM)
        a.G = (gtot - g0) < gmax ? (gtot - g0) : gmax;
        a.b_off = g0 * WNB;
        const int r = wave_launch(kern, "lstm_wave_fwd", a.L * a.G * NG, S::TOTAL2, &a, st);   // (single-layer launches
        if (r) return r;                                                                        //  simply leave W_OFF.. unused)
    }
    return 0;
}

int lstm_wave_fwd(const float* P0, const float* Q0, const float* Whh0, float* h0, float* c0, float* a0,
                  const float* Wih1, const float* b1, const float* Whh1, float* h1, float* c1, float* a1,
                  void* xchg, int T, int B, int H, int L, int mode, cudaStream_t st, void* hp0, void* hp1, long long hps,
                  const void* packed) {
    WaveFwdArgs a{P0, Q0, Whh0, h0, c0, a0, Wih1, b1, Whh1, h1, c1, a1, reinterpret_cast<uint4*>(xchg), T, B, 0, 0, 0, L,
                  reinterpret_cast<__nv_bfloat16*>(hp0), reinterpret_cast<__nv_bfloat16*>(hp1), hps,
                  reinterpret_cast<const uint8_t*>(packed)};
    const bool x3 = mode == FHVAE_MODE_BF16X3;
    if (H == 256) return x3 ? launch_wave_fwd<true, 256>(a, st) : launch_wave_fwd<false, 256>(a, st);
    return x3 ? launch_wave_fwd<true, 128>(a, st) : launch_wave_fwd<false, 128>(a, st);
}

size_t lstm_wave_bwd_xchg_bytes(int T, int B, int H, int L) {
    int gs = B / WNB;
    const int gmax = wave_groups_per_launch(L, H), ng = H / WU;
    if (gs > gmax) gs = gmax;
    return ((size_t)WHDR + (size_t)L * gs * 2 * ng * ng * WRS + (L == 2 ? (size_t)gs * T * ng * WDG : 0)) * sizeof(uint4);
}

template <bool X3, int CH>
static int launch_wave_bwd(WaveBwdArgs a, cudaStream_t st) {
    using S = WaveBwdSmem<X3, CH>;
    constexpr int NG = CH / WU;
    auto kern = lstm_wave_bwd_kernel<X3, CH>;
    const int gtot = a.B / WNB, gmax = wave_groups_per_launch(a.L, CH);
    if (gmax < 1) {
        set_error("lstm_wave_bwd: the device has too few SMs for one group of %d CTAs x %d layers", NG, a.L);
        return FHVAE_ENOSUP;
    }
    a.Gs = gtot < gmax ? gtot : gmax;
    for (int g0 = 0; g0 < gtot; g0 += gmax) {
        a.G = (gtot - g0) < gmax ? (gtot - g0) : gmax;
        a.b_off = g0 * WNB;
        const int r = wave_launch(kern, "lstm_wave_bwd", a.L * a.G * NG, S::TOTAL2, &a, st);
        if (r) return r;
    }
    return 0;
}

int lstm_wave_bwd(const float* dh_all, const float* dh_last1, const float* dh_last0, const float* Whh1, const float* c1,
                  const float* a1, float* dg1, float* dgsum1, const float* Wih1, const float* Whh0, const float* c0,
                  const float* a0, float* dg0, float* dgsum0, void* xchg, int T, int B, int H, int L, int mode, cudaStream_t st,
                  void* dgp1, void* dgp0, long long dgps, const void* packed) {
    WaveBwdArgs a{dh_all, dh_last1, dh_last0, Whh1, c1, a1, dg1, dgsum1, Wih1, Whh0, c0, a0, dg0, dgsum0,
                  reinterpret_cast<uint4*>(xchg), T, B, 0, 0, 0, L,
                  reinterpret_cast<__nv_bfloat16*>(dgp1), reinterpret_cast<__nv_bfloat16*>(dgp0), dgps,
                  reinterpret_cast<const uint8_t*>(packed)};
    const bool x3 = mode == FHVAE_MODE_BF16X3;
    if (H == 256) return x3 ? launch_wave_bwd<true, 256>(a, st) : launch_wave_bwd<false, 256>(a, st);
    return x3 ? launch_wave_bwd<true, 128>(a, st) : launch_wave_bwd<false, 128>(a, st);
}

template <bool X3, int CH>
static size_t pack_bytes(int L) { return WavePack<X3, CH>::n_images(L) * WavePack<X3, CH>::IMG; }

size_t lstm_wave_pack_bytes(int H, int L, int mode) {
    const bool x3 = mode == FHVAE_MODE_BF16X3;
    if (H == 256) return x3 ? pack_bytes<true, 256>(L) : pack_bytes<false, 256>(L);
    return x3 ? pack_bytes<true, 128>(L) : pack_bytes<false, 128>(L);
}

// layer-indexed weights: Whh_l0 (bottom / only layer), Wih1 + Whh_l1 (second layer of a 2-layer stack)
template <bool X3, int CH>
static void launch_pack(const WavePackArgs& a, cudaStream_t st) {
    wave_pack_kernel<X3, CH><<<WavePack<X3, CH>::n_images(a.L), WNT, 0, st>>>(a);
}
int lstm_wave_pack(const float* Whh_l0, const float* Wih1, const float* Whh_l1, void* out, int H, int L, int mode,
                   cudaStream_t st) {
    WavePackArgs a{{Whh_l0, Whh_l1}, Wih1, reinterpret_cast<uint8_t*>(out), L};
    const bool x3 = mode == FHVAE_MODE_BF16X3;
    if (H == 256) { if (x3) launch_pack<true, 256>(a, st); else launch_pack<false, 256>(a, st); }
    else          { if (x3) launch_pack<true, 128>(a, st); else launch_pack<false, 128>(a, st); }
    FHVAE_LAUNCH_CHECK("lstm_wave_pack");
    return 0;
}

}  // namespace fhvae
