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
//  * Wavefront over layers: layer 1 of step t only needs h0_t, so the layer-1 groups run one step behind
//    the layer-0 groups in the SAME launch: 2T dependent steps become T+1.  Layer-1 CTAs keep their
//    W_ih slice resident in shared memory (bf16 hi/lo UMMA layout) and compute their own input
//    projection  W_ih1 * h0_t  from the very words layer 0 published for its own peers (the exchange
//    buffer is T-deep, so nothing is overwritten); those MMAs are issued a step ahead, in the shadow of
//    the wait for h1_{t-1}.  The layer-1 input projection GEMM and its (T,B,4H) buffer disappear.
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
constexpr int WNT = 512;           // threads per CTA
constexpr int WNB = 32;            // batch rows per group
constexpr int WSLICE = 512;        // uint4 words per published slice (2 parts x 32 rows x 4 chunks x 2 halves)
constexpr int WHDR = 512;          // header of the exchange buffer in uint4 (8 KB of launch counters)
constexpr int WMAXG = 32;          // counters for up to 32 groups per layer
constexpr int WMAXT = 63;          // flag = epoch * 64 + t + 1

__device__ __forceinline__ uint4 ld_ll(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_ll(uint4* p, uint32_t d0, uint32_t d1, uint32_t flag) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(d0), "r"(flag), "r"(d1), "r"(flag) : "memory");
}
__device__ __forceinline__ void wait_ll(uint4& v, const uint4* p, uint32_t flag) {
    uint32_t spins = 0;
    while (v.y != flag || v.w != flag) {
        if (++spins > FHVAE_SPIN_LIMIT) __trap();
        v = ld_ll(p);
    }
}

#ifdef FHVAE_TIMELINE
__device__ long long g_wave_tl[2][32][16];
#define WTL(step, slot) do { if (rank == 0 && grp == 0 && threadIdx.x == 0 && (step) < 32) g_wave_tl[layer][step][slot] = clock64(); } while (0)
extern "C" int fhvae_debug_wave_timeline(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_wave_tl, sizeof(g_wave_tl)); }
#else
#define WTL(step, slot) do { } while (0)
#endif

struct WaveFwdArgs {
    const float* P0; const float* Q0; const float* Whh0; float* h0; float* c0; float* a0;
    const float* Wih1; const float* b1; const float* Whh1; float* h1; float* c1; float* a1;
    uint4* xchg;
    int T, B, b_off, G, Gs, L;     // B = row stride (whole batch), G groups in this launch, Gs = slot stride
};

template <bool X3>
struct WaveFwdSmem {
    static constexpr int H_PART = WNB * WH * 2;                // 32 x 256 bf16 = 16 KB
    static constexpr int H_BUF = (X3 ? 2 : 1) * H_PART;
    static constexpr int H_OFF = 0;                            // h_{t-1} of this layer (MMA B operand)
    static constexpr int X_OFF = H_BUF;                        // h0_t of the layer below (layer 1 only)
    static constexpr int G_OFF = 2 * H_BUF;
    static constexpr int G_BYTES = 4 * WNB * (WU + 1) * 4;     // gates[4][NB][33] fp32
    static constexpr int BAR_OFF = G_OFF + G_BYTES;
    static constexpr int W_OFF = (BAR_OFF + 64 + 1023) / 1024 * 1024;
    static constexpr int W_PART = WNC * WH * 2;                // 128 x 256 bf16 = 64 KB
    static constexpr int W_BYTES = (X3 ? 2 : 1) * W_PART;
    static constexpr int TOTAL1 = W_OFF;                       // single layer: no resident W_ih
    static constexpr int TOTAL2 = W_OFF + W_BYTES;
};

// pull NS slices (all 8, or the 7 peers) of one published step into an MMA operand buffer
template <bool X3, int NS, bool SKIP_OWN>
__device__ __forceinline__ void wave_load(const uint4* slot, int rank, uint4 (&v)[NS]) {
    constexpr int NW = (X3 ? 2 : 1) * 256;
    if ((int)threadIdx.x < NW) {
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const int src = SKIP_OWN ? i + (i >= rank ? 1 : 0) : i;
            v[i] = ld_ll(slot + src * WSLICE + threadIdx.x);
        }
    }
}
template <bool X3, int NS, bool SKIP_OWN>
__device__ __forceinline__ void wave_store(const uint4* slot, int rank, uint32_t flag, uint4 (&v)[NS], uint8_t* dst) {
    using S = WaveFwdSmem<X3>;
    constexpr int NW = (X3 ? 2 : 1) * 256;
    const int tid = threadIdx.x;
    if (tid < NW) {
        const int part = tid >> 8, rem = tid & 255, chunk = rem >> 1, half = rem & 1;
        const int kcl = chunk / WNB, row = chunk % WNB;
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const int src = SKIP_OWN ? i + (i >= rank ? 1 : 0) : i;
            wait_ll(v[i], slot + src * WSLICE + tid, flag);
            *reinterpret_cast<uint2*>(dst + part * S::H_PART + (uint32_t)(src * 4 + kcl) * (WNB * 16) + row * 16 + half * 8) =
                make_uint2(v[i].x, v[i].z);
        }
    }
}

template <bool X3>
__global__ void __launch_bounds__(WNT, 1) lstm_wave_fwd_kernel(const __grid_constant__ WaveFwdArgs a) {
    using S = WaveFwdSmem<X3>;
    constexpr int NB = WNB, NT = WNT, CH = WH, UC = WU;
    constexpr int NW = (X3 ? 2 : 1) * 256;     // LL words of one slice actually used
    constexpr int NCG = NT / 128;              // column groups: warps sharing one TMEM lane quarter
    constexpr int CPW = NB / NCG;              // batch rows (TMEM columns) per thread in the gate phase
    constexpr int RPT = NB * 32 / NT;          // batch rows per thread in the cell phase
    constexpr int H4 = 4 * CH;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* hb = smem + S::H_OFF;
    uint8_t* xb = smem + S::X_OFF;
    float (*gates)[NB][UC + 1] = reinterpret_cast<float (*)[NB][UC + 1]>(smem + S::G_OFF);
    uint64_t* mma_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);
    uint32_t* epoch_slot = tmem_slot + 1;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, cg = warp >> 2;    // gate (= TMEM lane quarter), column group
    const int layer = blockIdx.x / (a.G * WG);
    const int grp = (blockIdx.x / WG) % a.G;
    const int rank = blockIdx.x % WG;
    const int T = a.T, B = a.B;
    const int b0 = a.b_off + grp * NB;
    const float* P = layer ? nullptr : a.P0;
    const float* Q = layer ? nullptr : a.Q0;
    const float* W_hh = layer ? a.Whh1 : a.Whh0;
    float* h_all = layer ? a.h1 : a.h0;
    float* c_all = layer ? a.c1 : a.c0;
    float* acts = layer ? a.a1 : a.a0;
    // exchange buffer: [header][layer][group][t][src CTA][WSLICE] uint4
    uint32_t* cnt = reinterpret_cast<uint32_t*>(a.xchg) + (((a.L - 1) * 2 + layer) * WMAXG + grp) * WG + rank;
    uint4* own = a.xchg + WHDR + ((size_t)(layer * a.Gs + grp) * T) * (WG * WSLICE);
    const uint4* below = a.xchg + WHDR + ((size_t)grp * T) * (WG * WSLICE);      // layer 0 of the same group

    // ---- prologue: TMEM, barrier, W_hh slice -> TMEM (lane n = gate*32 + unit, column k/2 = bf16 pair)
    constexpr int NACC = 2;
    constexpr int WCOLS = CH / 2;
    constexpr int ACOL = (X3 ? 2 : 1) * WCOLS;
    constexpr int TCOLS = (ACOL + NACC * NB) <= 256 ? 256 : 512;
    if (warp == 0) tmem_alloc<TCOLS>(tmem_slot);
    if (tid == 32) { mbar_init(mma_bar, 1); fence_mbar_init(); *epoch_slot = *cnt; }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t fbase = (*epoch_slot) << 6;
    {
        const float* src = W_hh + (size_t)(q * CH + rank * UC + lane) * CH + cg * 64;
        float4 wv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) wv[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * 32);
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
        tmem_wait_st();
    }
    constexpr uint32_t W_LBO = WNC * 16, H_LBO = NB * 16, SBO_ = 128;
    if (layer == 1) {
        // resident W_ih1 slice as an smem A operand: row n = gate*32 + unit, K-chunk planes of W_LBO bytes
        const float* src = a.Wih1 + (size_t)(q * CH + rank * UC + lane) * CH + cg * 64;
        const int r = q * 32 + lane;
#pragma unroll 2
        for (int i = 0; i < 8; ++i) {
            const float4 x0 = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
            const float4 x1 = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
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
    // time-invariant addend for this thread's (gate q, unit lane) column, rows cg*CPW ..
    const int col = q * CH + rank * UC + lane;
    float qv[CPW];
    {
        const float bias = (layer == 1 && a.b1) ? __ldg(a.b1 + col) : 0.f;
#pragma unroll
        for (int b = 0; b < CPW; ++b) qv[b] = bias + (Q ? __ldg(Q + (size_t)(b0 + cg * CPW + b) * H4 + col) : 0.f);
    }
    float creg[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) creg[i] = 0.f;

    const uint32_t tmem_d = tmem_base + ACOL;
    constexpr uint32_t idesc = make_idesc_bf16(WNC, NB);
    const uint32_t hb_u = smem_u32(hb), xb_u = smem_u32(xb), wih_u = smem_u32(smem + S::W_OFF);

    // input projection of layer 1 for step `t`: acc = W_ih1_slice * h0_t^T (A and B from shared memory)
    auto issue_input = [&]() {
        const uint64_t dwh0 = make_smem_desc(wih_u, W_LBO, SBO_), dwl0 = make_smem_desc(wih_u + S::W_PART, W_LBO, SBO_);
        const uint64_t dxh0 = make_smem_desc(xb_u, H_LBO, SBO_), dxl0 = make_smem_desc(xb_u + S::H_PART, H_LBO, SBO_);
#pragma unroll
        for (int s = 0; s < CH / 16; ++s) {
            const uint64_t iw = (uint64_t)((s * 2 * W_LBO) >> 4), ix = (uint64_t)((s * 2 * H_LBO) >> 4);
            const uint32_t td = tmem_d + (uint32_t)((s % NACC) * NB);
            const uint32_t first = s >= NACC ? 1u : 0u;
            if (X3) {
                umma_bf16(td, dwl0 + iw, dxh0 + ix, idesc, first);
                umma_bf16(td, dwh0 + iw, dxl0 + ix, idesc, 1u);
                umma_bf16(td, dwh0 + iw, dxh0 + ix, idesc, 1u);
            } else {
                umma_bf16(td, dwh0 + iw, dxh0 + ix, idesc, first);
            }
        }
    };
    if (layer == 1) {
        uint4 xv[8];
        wave_load<X3, 8, false>(below, rank, xv);
        wave_store<X3, 8, false>(below, rank, fbase + 1, xv, xb);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (layer == 1 && warp == 0 && elect_one()) issue_input();
    uint32_t ph = 0;
    const int kglob = rank * UC + lane;                        // this thread's unit in the cell phase
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
        }
    };

    for (int t = 0; t < T; ++t) {
        float pv[CPW];
        const float* Pt = P ? P + ((size_t)t * B + b0 + cg * CPW) * H4 + col : nullptr;
#pragma unroll
        for (int b = 0; b < CPW; ++b) pv[b] = Pt ? __ldg(Pt + (size_t)b * H4) : 0.f;

        WTL(t, 0);
        if (t > 0) {
            // h_{t-1}: the 7 peer slices (own slice was written locally by the cell phase)
            const uint4* slot = own + (size_t)(t - 1) * (WG * WSLICE);
            uint4 hv[7];
            wave_load<X3, 7, true>(slot, rank, hv);
            store_saved(t - 1);
            wave_store<X3, 7, true>(slot, rank, fbase + t, hv, hb);
            fence_proxy_async();
            __syncthreads();
        }
        WTL(t, 1);
        float acc[CPW];
        if (t > 0 || layer == 1) {
            if (warp == 0 && elect_one()) {
                tc_fence_after();
                if (t > 0) {
                    const uint64_t dhh0 = make_smem_desc(hb_u, H_LBO, SBO_);
                    const uint64_t dhl0 = make_smem_desc(hb_u + S::H_PART, H_LBO, SBO_);
#pragma unroll
                    for (int s = 0; s < CH / 16; ++s) {
                        const uint64_t ih = (uint64_t)((s * 2 * H_LBO) >> 4);
                        const uint32_t awh = tmem_base + (uint32_t)(s * 8), awl = awh + WCOLS;
                        const uint32_t td = tmem_d + (uint32_t)((s % NACC) * NB);
                        const uint32_t first = (layer == 1 || s >= NACC) ? 1u : 0u;
                        if (X3) {
                            umma_bf16_ts(td, awl, dhh0 + ih, idesc, first);
                            umma_bf16_ts(td, awh, dhl0 + ih, idesc, 1u);
                            umma_bf16_ts(td, awh, dhh0 + ih, idesc, 1u);
                        } else {
                            umma_bf16_ts(td, awh, dhh0 + ih, idesc, first);
                        }
                    }
                }
                umma_commit(mma_bar);
            }
            WTL(t, 2);
            mbar_wait(mma_bar, ph & 1);
            ++ph;
            WTL(t, 3);
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
        WTL(t, 4);
        // layer 1: start fetching h0_{t+1} (published by layer 0 one or more steps ago); lands during the gate math
        const bool more_in = (layer == 1) && (t + 1 < T);
        const uint4* nslot = below + (size_t)(t + 1) * (WG * WSLICE);
        uint4 xv[8];
        if (more_in) wave_load<X3, 8, false>(nslot, rank, xv);
#pragma unroll
        for (int b = 0; b < CPW; ++b) {
            const float pre = acc[b] + pv[b] + qv[b];
            aval[b] = (q == 2) ? tanhf_fast(pre) : sigmoidf_fast(pre);
            gates[q][cg * CPW + b][lane] = aval[b];
        }
        tc_fence_before();
        __syncthreads();
        WTL(t, 5);
        // cell update: thread = (unit = lane, rows warp*RPT ..)
        const uint32_t hoff_k = (uint32_t)(kglob >> 3) * H_LBO + (uint32_t)(kglob & 7) * 2;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const int b = warp * RPT + i;
            const float ig = gates[0][b][lane], fg = gates[1][b][lane], gg = gates[2][b][lane], og = gates[3][b][lane];
            const float c = fmaf(fg, creg[i], ig * gg);
            creg[i] = c;
            const float h = og * tanhf_fast(c);
            hreg[i] = h;
            const __nv_bfloat16 hh = __float2bfloat16_rn(h);
            *reinterpret_cast<__nv_bfloat16*>(hb + hoff_k + b * 16) = hh;
            if (X3)
                *reinterpret_cast<__nv_bfloat16*>(hb + S::H_PART + hoff_k + b * 16) = __float2bfloat16_rn(h - __bfloat162float(hh));
        }
        if (more_in) wave_store<X3, 8, false>(nslot, rank, fbase + t + 2, xv, xb);
        fence_proxy_async();
        __syncthreads();
        WTL(t, 6);
        // publish this CTA's slice of h_t: peers need it for step t+1, the layer above for its step t
        if ((t + 1 < T || (layer == 0 && a.L == 2)) && tid < NW) {
            const int part = tid >> 8, rem = tid & 255, chunk = rem >> 1, half = rem & 1;
            const int kcl = chunk / NB, row = chunk % NB;
            const uint2 d = *reinterpret_cast<const uint2*>(hb + part * S::H_PART + (uint32_t)(rank * 4 + kcl) * H_LBO + row * 16 + half * 8);
            st_ll(own + (size_t)t * (WG * WSLICE) + rank * WSLICE + tid, d.x, d.y, fbase + t + 1);
        }
        WTL(t, 7);
        if (more_in && warp == 0 && elect_one()) {
            tc_fence_after();
            issue_input();
        }
        WTL(t, 8);
    }
    store_saved(T - 1);
    if (tid == 0) *cnt = (fbase >> 6) + 1;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TCOLS>(tmem_base);
}

bool lstm_wave_supported(int T, int B, int H, int L) {
    return H == WH && B % WNB == 0 && T >= 1 && T <= WMAXT && (L == 1 || L == 2);
}

static int wave_groups_per_launch(int L) { return (kNumSM / (WG * L)) > WMAXG ? WMAXG : kNumSM / (WG * L); }   // 18 or 9

size_t lstm_wave_xchg_bytes(int T, int B, int L) {
    int gs = B / WNB;
    const int gmax = wave_groups_per_launch(L);
    if (gs > gmax) gs = gmax;
    return ((size_t)WHDR + (size_t)L * gs * T * WG * WSLICE) * sizeof(uint4);
}

template <bool X3>
static int launch_wave_fwd(WaveFwdArgs a, cudaStream_t st) {
    using S = WaveFwdSmem<X3>;
    static bool attr = false;
    auto kern = lstm_wave_fwd_kernel<X3>;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL2);
        if (e != cudaSuccess) {
            set_error("lstm_wave_fwd: cudaFuncSetAttribute(%d B): %s", S::TOTAL2, cudaGetErrorString(e));
            return (int)e;
        }
        attr = true;
    }
    const int gtot = a.B / WNB, gmax = wave_groups_per_launch(a.L);
    a.Gs = gtot < gmax ? gtot : gmax;
    for (int g0 = 0; g0 < gtot; g0 += gmax) {          // all CTAs of a launch must be co-resident (one per SM)
        a.G = (gtot - g0) < gmax ? (gtot - g0) : gmax;
        a.b_off = g0 * WNB;
        kern<<<a.L * a.G * WG, WNT, a.L == 2 ? S::TOTAL2 : S::TOTAL1, st>>>(a);
        FHVAE_LAUNCH_CHECK("lstm_wave_fwd");
    }
    return 0;
}

int lstm_wave_fwd(const float* P0, const float* Q0, const float* Whh0, float* h0, float* c0, float* a0,
                  const float* Wih1, const float* b1, const float* Whh1, float* h1, float* c1, float* a1,
                  void* xchg, int T, int B, int L, int mode, cudaStream_t st) {
    WaveFwdArgs a{P0, Q0, Whh0, h0, c0, a0, Wih1, b1, Whh1, h1, c1, a1, reinterpret_cast<uint4*>(xchg), T, B, 0, 0, 0, L};
    return mode == FHVAE_MODE_BF16X3 ? launch_wave_fwd<true>(a, st) : launch_wave_fwd<false>(a, st);
}

}  // namespace fhvae
