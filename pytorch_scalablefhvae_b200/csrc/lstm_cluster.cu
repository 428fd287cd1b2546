// Persistent LSTM recurrence on thread-block clusters (H = 256).
//
// One cluster of 8 CTAs owns NB batch rows for all T steps of one layer.  CTA j of the cluster owns
// hidden units [32j, 32j+32) = 128 gate columns (4 gates x 32 units); its W_hh slice (128 x 256) stays
// resident in shared memory as bf16 hi/lo in the canonical K-major UMMA layout for the whole kernel.
// Per time step each CTA issues   D[128 gate cols, NB] = W_slice (128 x 256) * h_{t-1}^T (NB x 256)
// on tcgen05 (accumulator in TMEM), adds the precomputed input projection P[t] (+ time-invariant Q),
// applies sigmoid/tanh, exchanges the four gates of a unit through shared memory, updates the cell
// (c lives in registers across steps), and writes its 32-unit slice of h_t as bf16 hi/lo straight
// into the h operand buffers of all 8 CTAs (DSMEM vector stores), double-buffered, one cluster
// barrier per step.  h_t / c_t / gate activations also go to HBM in fp32 for BPTT.
//
// The backward kernel mirrors it: dgates_t is computed pointwise by the CTA that owns the units,
// dh_{t-1} = dgates_t W_hh is a split-K contraction over the 8 CTAs' gate-column slices, reduced
// through DSMEM (each CTA receives the 7 partial tiles of its own 32 units).
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace fhvae {

using namespace tc;

constexpr int CH = 256;            // hidden size served by the cluster kernels
constexpr int CL = 8;              // CTAs per cluster
constexpr int UC = CH / CL;        // 32 units per CTA
constexpr int NC = 4 * UC;         // 128 gate columns per CTA
constexpr int CNT = 512;           // threads per CTA: 4 TMEM lane quarters x 4 column groups

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_remote(uint32_t saddr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta));
    return r;
}
__device__ __forceinline__ void st_remote_v4(uint32_t raddr, const uint4& v) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Development aid (tools/timeline.py builds with -DFHVAE_TIMELINE): SM-clock stamps of CTA 0 / thread 0.
#ifdef FHVAE_TIMELINE
__device__ long long g_timeline[2][32][16];
#define TL(k, step, slot) do { if (blockIdx.x == 0 && threadIdx.x == 0 && (step) < 32) g_timeline[k][step][slot] = clock64(); } while (0)
#else
#define TL(k, step, slot) do { } while (0)
#endif

// shared-memory map of the forward kernel (bytes)
template <int NB, bool X3>
struct FwdSmem {
    static constexpr int W_PART = 0, W_BYTES = 0;              // W_hh slice lives in TMEM (A operand)
    static constexpr int H_PART = NB * CH * 2;                 // NB x 256 bf16
    static constexpr int H_BUF = (X3 ? 2 : 1) * H_PART;        // hi [lo]
    static constexpr int H_OFF = W_BYTES;
    static constexpr int G_OFF = H_OFF + 2 * H_BUF;
    static constexpr int G_BYTES = 4 * NB * (UC + 1) * 4;      // gates[4][NB][33] fp32
    static constexpr int BAR_OFF = G_OFF + G_BYTES;
    static constexpr int TOTAL = BAR_OFF + 64;
};

template <int NB, bool X3, int NT>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NT, 1)
lstm_fwd_cluster_kernel(const float* __restrict__ P, const float* __restrict__ Q,
                        const float* __restrict__ W_hh, float* __restrict__ h_all,
                        float* __restrict__ c_all, float* __restrict__ acts, float* __restrict__ xchg,
                        int T, int B) {
    using S = FwdSmem<NB, X3>;
    constexpr int VEC_PER_PART = 4 * NB;       // 16-byte vectors of one CTA's h slice, per bf16 part
    constexpr int NVEC = (X3 ? 2 : 1) * VEC_PER_PART;
    // L2-resident exchange buffer of this cluster: [2 buffers][CL slices][NVEC] uint4 (DSMEM moves only
    // ~17 B/clk/SM; an all-gather through L2 is ~2x faster)
    uint4* xg = reinterpret_cast<uint4*>(xchg) + (size_t)(blockIdx.x / CL) * (2 * CL * NVEC);
    constexpr int NCG = NT / 128;              // column groups: warps sharing one TMEM lane quarter
    constexpr int CPW = NB / NCG;              // batch rows (TMEM columns) per thread in the gate phase
    constexpr int RPT = NB * 32 / NT;          // batch rows per thread in the cell phase
    extern __shared__ __align__(1024) uint8_t smem[];
    float (*gates)[NB][UC + 1] = reinterpret_cast<float (*)[NB][UC + 1]>(smem + S::G_OFF);
    uint64_t* mma_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, cg = warp >> 2;    // gate (= TMEM lane quarter), column group
    const uint32_t rank = cluster_ctarank();
    const int b0 = (blockIdx.x / CL) * NB;
    constexpr int H4 = 4 * CH;

    // ---- prologue: TMEM, barrier, resident W slice as the TMEM A operand:
    //      lane n = gate*32 + unit  <->  W_hh row gate*H + 32*rank + unit;  column k/2 holds (k, k+1) as bf16
    constexpr int NACC = 2;                    // independent TMEM accumulators (k-steps round-robin)
    constexpr int WCOLS = CH / 2;              // 128 columns per bf16 part
    constexpr int ACOL = (X3 ? 2 : 1) * WCOLS; // accumulators start after the W part(s)
    constexpr int TCOLS = (ACOL + NACC * NB) <= 256 ? 256 : 512;
    if (warp == 0) tmem_alloc<TCOLS>(tmem_slot);
    if (tid == 32) { mbar_init(mma_bar, 1); fence_mbar_init(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    {
        // thread (q, lane) owns TMEM lane 32q+lane = W row; column group cg covers k in [64cg, 64cg+64)
        static_assert(NT == 512, "W staging assumes 4 column groups");
        const float* src = W_hh + (size_t)(q * CH + rank * UC + lane) * CH + cg * 64;
        float4 wv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) wv[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * 32);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {       // two 16-column halves keep the register footprint small
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 w4 = wv[hf * 8 + i];
                const float v[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const __nv_bfloat16 a = __float2bfloat16_rn(v[2 * j]), b = __float2bfloat16_rn(v[2 * j + 1]);
                    hi[2 * i + j] = (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
                    lo[2 * i + j] = pack_bf16(v[2 * j] - __bfloat162float(a), v[2 * j + 1] - __bfloat162float(b));
                }
            }
            tmem_st16(ta + hf * 16, hi);
            if (X3) tmem_st16(ta + WCOLS + hf * 16, lo);
        }
        tmem_wait_st();
    }
    // time-invariant addend for this thread's (gate q, unit lane) column, rows cg*CPW ..
    const int col = q * CH + rank * UC + lane;
    float qv[CPW];
#pragma unroll
    for (int b = 0; b < CPW; ++b) qv[b] = Q ? __ldg(Q + (size_t)(b0 + cg * CPW + b) * H4 + col) : 0.f;
    float creg[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) creg[i] = 0.f;

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base + ACOL;
    constexpr uint32_t idesc = make_idesc_bf16(NC, NB);
    constexpr uint32_t H_LBO = NB * 16, SBO_ = 128;
    cluster_arrive();          // pairs with the wait at the top of step 0 (barriers initialised cluster-wide)

    for (int t = 0; t < T; ++t) {
        // prefetch this step's input projection
        float pv[CPW];
        const float* Pt = P ? P + ((size_t)t * B + b0 + cg * CPW) * H4 + col : nullptr;
#pragma unroll
        for (int b = 0; b < CPW; ++b) pv[b] = Pt ? __ldg(Pt + (size_t)b * H4) : 0.f;

        TL(0, t, 0);
        cluster_wait();        // every CTA's h_{t-1} slice is visible in the L2 exchange buffer
        TL(0, t, 1);
        float acc[CPW];
        if (t > 0) {
            {   // pull the 7 peer slices into this CTA's MMA operand buffer (own slice was written locally)
                uint8_t* hb = smem + S::H_OFF + (t & 1) * S::H_BUF;
                const uint4* xr = xg + (size_t)((t & 1) * CL) * NVEC;
                constexpr int PER = CL * NVEC / NT;
                uint4 val[PER];
#pragma unroll
                for (int i = 0; i < PER; ++i) val[i] = __ldcg(xr + tid + i * NT);
#pragma unroll
                for (int i = 0; i < PER; ++i) {
                    const int idx = tid + i * NT, src = idx / NVEC, v = idx % NVEC;
                    const int part = v / VEC_PER_PART, rem = v % VEC_PER_PART;
                    const int kcl = rem / NB, row = rem % NB;
                    if (src != (int)rank)
                        *reinterpret_cast<uint4*>(hb + part * S::H_PART + (uint32_t)(src * 4 + kcl) * (NB * 16) + row * 16) = val[i];
                }
                fence_proxy_async();
                __syncthreads();
            }
            if (warp == 0 && elect_one()) {
                tc_fence_after();
                TL(0, t, 12);
                const uint32_t hb = smem_u32(smem + S::H_OFF + (t & 1) * S::H_BUF);
                // B descriptors differ only in the 16-byte-granular start address: base + constant;
                // A (W_hh slice) comes from TMEM: 8 columns per K=16 step
                const uint64_t dhh0 = make_smem_desc(hb, H_LBO, SBO_);
                const uint64_t dhl0 = make_smem_desc(hb + S::H_PART, H_LBO, SBO_);
#pragma unroll
                for (int s = 0; s < CH / 16; ++s) {
                    const uint64_t ih = (uint64_t)((s * 2 * H_LBO) >> 4);
                    const uint32_t awh = tmem_base + (uint32_t)(s * 8), awl = awh + WCOLS;
                    const uint32_t td = tmem_d + (uint32_t)((s % NACC) * NB);
                    const uint32_t first = s >= NACC ? 1u : 0u;
                    if (X3) {
                        umma_bf16_ts(td, awl, dhh0 + ih, idesc, first);
                        umma_bf16_ts(td, awh, dhl0 + ih, idesc, 1u);
                        umma_bf16_ts(td, awh, dhh0 + ih, idesc, 1u);
                    } else {
                        umma_bf16_ts(td, awh, dhh0 + ih, idesc, first);
                    }
                }
                umma_commit(mma_bar);
                TL(0, t, 2);
            }
            mbar_wait(mma_bar, (t - 1) & 1);
            TL(0, t, 3);
            tc_fence_after();
            tmem_ld_nb<CPW>(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * CPW), acc);
#pragma unroll
            for (int a = 1; a < NACC; ++a) {
                float part[CPW];
                tmem_ld_nb<CPW>(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * NB + cg * CPW), part);
#pragma unroll
                for (int b = 0; b < CPW; ++b) acc[b] += part[b];
            }
            TL(0, t, 4);
        } else {
#pragma unroll
            for (int b = 0; b < CPW; ++b) acc[b] = 0.f;
        }
        // gate activation: q = gate (0:i 1:f 2:g 3:o), lane = unit, registers = batch rows
        // (HBM stores are deferred past the cluster arrive: its release fence would otherwise wait for them)
        float aval[CPW];
#pragma unroll
        for (int b = 0; b < CPW; ++b) {
            const float pre = acc[b] + pv[b] + qv[b];
            aval[b] = (q == 2) ? tanhf_fast(pre) : sigmoidf_fast(pre);
            gates[q][cg * CPW + b][lane] = aval[b];
        }
        TL(0, t, 5);
        tc_fence_before();
        __syncthreads();
        TL(0, t, 6);
        // cell update: thread = (unit = lane, rows warp*RPT ..)
        uint8_t* hnext = smem + S::H_OFF + ((t + 1) & 1) * S::H_BUF;
        const int kglob = rank * UC + lane;                    // this unit's K index in the h operand
        float hreg[RPT];
        const uint32_t hoff_k = (uint32_t)(kglob >> 3) * H_LBO + (uint32_t)(kglob & 7) * 2;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const int b = warp * RPT + i;
            const float ig = gates[0][b][lane], fg = gates[1][b][lane], gg = gates[2][b][lane],
                        og = gates[3][b][lane];
            const float c = fmaf(fg, creg[i], ig * gg);
            creg[i] = c;
            const float h = og * tanhf_fast(c);
            hreg[i] = h;
            const __nv_bfloat16 hh = __float2bfloat16_rn(h);
            *reinterpret_cast<__nv_bfloat16*>(hnext + hoff_k + b * 16) = hh;
            if (X3)
                *reinterpret_cast<__nv_bfloat16*>(hnext + S::H_PART + hoff_k + b * 16) =
                    __float2bfloat16_rn(h - __bfloat162float(hh));
        }
        TL(0, t, 7);
        __syncthreads();
        TL(0, t, 8);
        // publish this CTA's slice (4 K-chunks x NB rows x 16 B per part) in the L2 exchange buffer
        if (t + 1 < T && tid < NVEC) {
            const int part = tid / VEC_PER_PART, rem = tid % VEC_PER_PART;
            const int kcl = rem / NB, row = rem % NB;
            const uint4 val = *reinterpret_cast<const uint4*>(hnext + part * S::H_PART +
                                                             (uint32_t)(rank * 4 + kcl) * H_LBO + row * 16);
            xg[(size_t)(((t + 1) & 1) * CL + rank) * NVEC + tid] = val;
        }
        TL(0, t, 9);
        cluster_arrive();
        TL(0, t, 10);
        // saved-for-backward state -> HBM, off the critical path
        float* At = acts + ((size_t)t * B + b0 + cg * CPW) * H4 + col;
#pragma unroll
        for (int b = 0; b < CPW; ++b) At[(size_t)b * H4] = aval[b];
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const size_t o = ((size_t)t * B + b0 + warp * RPT + i) * CH + kglob;
            h_all[o] = hreg[i];
            c_all[o] = creg[i];
        }
        TL(0, t, 11);
    }
    cluster_wait();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TCOLS>(tmem_base);
}

// ================================================================================================
// BPTT
// ================================================================================================
template <int NB, bool X3>
struct BwdSmem {
    static constexpr int W_HALF = 128 * NC * 2;                // 128 units x 128 gate cols bf16 = 32 KB
    static constexpr int W_PART = 2 * W_HALF;                  // 256 units: 64 KB
    static constexpr int W_BYTES = (X3 ? 2 : 1) * W_PART;
    static constexpr int G_PART = NB * NC * 2;                 // dgates operand, NB x 128 bf16
    static constexpr int G_OFF = W_BYTES;
    static constexpr int G_BYTES = (X3 ? 2 : 1) * G_PART;
    static constexpr int RSTRIDE = NB + 4;                     // floats per (src, unit) row of the receive buffer
    static constexpr int R_OFF = G_OFF + G_BYTES;
    static constexpr int R_BYTES = CL * UC * RSTRIDE * 4;
    static constexpr int BAR_OFF = R_OFF + R_BYTES;
    static constexpr int TOTAL = BAR_OFF + 64;
};

template <int NB, bool X3, int NT>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NT, 1)
lstm_bwd_cluster_kernel(const float* __restrict__ dh_all, const float* __restrict__ dh_last,
                        const float* __restrict__ W_hh, const float* __restrict__ c_all,
                        const float* __restrict__ acts, float* __restrict__ dgates,
                        float* __restrict__ dgsum, float* __restrict__ xchg, int T, int B) {
    using S = BwdSmem<NB, X3>;
    // L2-resident reduce-scatter buffer of this cluster: [2 buffers][dst CTA][src CTA][NB rows][32 units] fp32
    float* xg = xchg + (size_t)(blockIdx.x / CL) * (2 * CL * CL * NB * UC);
    constexpr int NCG = NT / 128;                              // column groups per TMEM lane quarter
    constexpr int CPW = NB / NCG;                              // TMEM columns per thread in the scatter phase
    constexpr int RPT = NB * 32 / NT;                          // batch rows per thread in the pointwise phase
    constexpr int WIT = CH * (NC / 8) / NT;                    // resident-W items per thread
    static_assert(RPT == 1 || RPT == 2 || RPT % 4 == 0, "receive-buffer reads are scalar / float2 / float4");
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* w_hi = smem;
    uint8_t* w_lo = smem + S::W_PART;
    uint8_t* g_hi = smem + S::G_OFF;
    uint8_t* g_lo = g_hi + S::G_PART;
    float* recv = reinterpret_cast<float*>(smem + S::R_OFF);   // [src][unit][RSTRIDE]
    uint64_t* mma_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, cg = warp >> 2;
    const uint32_t rank = cluster_ctarank();
    const int b0 = (blockIdx.x / CL) * NB;
    constexpr int H4 = 4 * CH;

    constexpr int NACC = 2;                                    // independent accumulators per unit half
    constexpr int TCOLS = 2 * NACC * NB < 32 ? 32 : 2 * NACC * NB;
    if (warp == 0) tmem_alloc<TCOLS>(tmem_slot);
    if (tid == 32) { mbar_init(mma_bar, 1); fence_mbar_init(); }
    // resident A operand: A[n][k] = W_hh[(g*H + 32*rank + u) * H + n],  k = g*32 + u, n = output unit
#pragma unroll 2
    for (int i = 0; i < WIT; ++i) {
        const int item = tid + i * NT;
        const int n = item & (CH - 1), kc = item >> 8;         // lanes <-> consecutive n: coalesced
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = kc * 8 + j, g = k >> 5, u = k & 31;
            v[j] = __ldg(W_hh + (size_t)(g * CH + rank * UC + u) * CH + n);
        }
        const uint32_t off = (uint32_t)(n >> 7) * S::W_HALF + (uint32_t)(kc * 128 + (n & 127)) * 16;
        if (X3) {
            uint4 hi, lo;
            split_bf16(v, hi, lo);
            *reinterpret_cast<uint4*>(w_hi + off) = hi;
            *reinterpret_cast<uint4*>(w_lo + off) = lo;
        } else {
            *reinterpret_cast<uint4*>(w_hi + off) =
                make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        }
    }
    float dcreg[RPT], gsum[4][RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        dcreg[i] = 0.f;
        gsum[0][i] = gsum[1][i] = gsum[2][i] = gsum[3][i] = 0.f;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;
    constexpr uint32_t idesc = make_idesc_bf16(128, NB);
    constexpr uint32_t W_LBO = 128 * 16, G_LBO = NB * 16, SBO_ = 128;
    const int ucol = rank * UC + lane;                         // this thread's hidden unit
    cluster_arrive();

    for (int t = T - 1; t >= 0; --t) {
        // ---- prefetch everything the pointwise step needs (thread = unit `lane`, rows warp*RPT + i)
        float a_i[RPT], a_f[RPT], a_g[RPT], a_o[RPT], c_t[RPT], c_p[RPT], dh[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const int b = b0 + warp * RPT + i;
            const size_t r4 = ((size_t)t * B + b) * H4, r1 = ((size_t)t * B + b) * CH;
            a_i[i] = __ldg(acts + r4 + ucol);
            a_f[i] = __ldg(acts + r4 + CH + ucol);
            a_g[i] = __ldg(acts + r4 + 2 * CH + ucol);
            a_o[i] = __ldg(acts + r4 + 3 * CH + ucol);
            c_t[i] = __ldg(c_all + r1 + ucol);
            c_p[i] = t ? __ldg(c_all + r1 - (size_t)B * CH + ucol) : 0.f;
            float d = dh_all ? __ldg(dh_all + r1 + ucol) : 0.f;
            if (t == T - 1 && dh_last) d += __ldg(dh_last + (size_t)b * CH + ucol);
            dh[i] = d;
        }
        TL(1, T - 1 - t, 0);
        cluster_wait();        // the 8 partial dh tiles for this step have landed in recv
        TL(1, T - 1 - t, 1);
        if (t < T - 1) {
            // sum the 8 partial tiles of this CTA's 32 units (fixed order: deterministic), coalesced from L2
            const float* rp = xg + (size_t)((((t + 1) & 1) * CL + rank) * CL) * (NB * UC) + lane;
            float part[CL][RPT];
#pragma unroll
            for (int src = 0; src < CL; ++src)
#pragma unroll
                for (int i = 0; i < RPT; ++i)
                    part[src][i] = __ldcg(rp + (size_t)src * (NB * UC) + (warp * RPT + i) * UC);
#pragma unroll
            for (int src = 0; src < CL; ++src)
#pragma unroll
                for (int i = 0; i < RPT; ++i) dh[i] += part[src][i];
        }
        TL(1, T - 1 - t, 2);
        // ---- pointwise BPTT; running sum, bf16 operand in smem (HBM store deferred)
        float gkeep[4][RPT];
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
                const int k = g * 32 + lane;
                const uint32_t off = (uint32_t)(k >> 3) * G_LBO + (uint32_t)bl * 16 + (uint32_t)(k & 7) * 2;
                const __nv_bfloat16 hh = __float2bfloat16_rn(gq[g]);
                *reinterpret_cast<__nv_bfloat16*>(g_hi + off) = hh;
                if (X3) *reinterpret_cast<__nv_bfloat16*>(g_lo + off) = __float2bfloat16_rn(gq[g] - __bfloat162float(hh));
            }
        }
        TL(1, T - 1 - t, 3);
        if (t > 0) {
            fence_proxy_async();
            __syncthreads();
            TL(1, T - 1 - t, 4);
            // ---- partial dh_{t-1}[unit, b] = sum over this CTA's 128 gate columns
            if (warp == 0 && elect_one()) {
                tc_fence_after();
                const uint64_t dwh0 = make_smem_desc(smem_u32(w_hi), W_LBO, SBO_);
                const uint64_t dwl0 = make_smem_desc(smem_u32(w_lo), W_LBO, SBO_);
                const uint64_t dgh0 = make_smem_desc(smem_u32(g_hi), G_LBO, SBO_);
                const uint64_t dgl0 = make_smem_desc(smem_u32(g_lo), G_LBO, SBO_);
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
                    for (int s = 0; s < NC / 16; ++s) {
                        const uint64_t iw = (uint64_t)((hf * S::W_HALF + s * 2 * W_LBO) >> 4);
                        const uint64_t ig = (uint64_t)((s * 2 * G_LBO) >> 4);
                        const uint32_t td = tmem_d + (uint32_t)((hf * NACC + (s % NACC)) * NB);
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
                umma_commit(mma_bar);
                TL(1, T - 1 - t, 5);
            }
            mbar_wait(mma_bar, (T - 1 - t) & 1);
            TL(1, T - 1 - t, 6);
            tc_fence_after();
            // ---- reduce-scatter: TMEM lane = unit; quarter q holds units 32q.. (-> CTA q) and 128+32q.. (-> CTA 4+q)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                float pv[CPW];
                tmem_ld_nb<CPW>(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * NACC * NB + cg * CPW), pv);
#pragma unroll
                for (int a = 1; a < NACC; ++a) {
                    float part[CPW];
                    tmem_ld_nb<CPW>(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)((hf * NACC + a) * NB + cg * CPW), part);
#pragma unroll
                    for (int b = 0; b < CPW; ++b) pv[b] += part[b];
                }
                const uint32_t dst = hf * 4 + q;
                float* wp = xg + (size_t)(((t & 1) * CL + dst) * CL + rank) * (NB * UC) + (cg * CPW) * UC + lane;
#pragma unroll
                for (int i = 0; i < CPW; ++i) wp[i * UC] = pv[i];           // 128-byte lines per row
            }
            tc_fence_before();
        }
        TL(1, T - 1 - t, 7);
        cluster_arrive();
        TL(1, T - 1 - t, 8);
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            float* dg = dgates + ((size_t)t * B + b0 + warp * RPT + i) * H4 + ucol;
#pragma unroll
            for (int g = 0; g < 4; ++g) dg[g * CH] = gkeep[g][i];
        }
    }
    cluster_wait();
    if (dgsum) {
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            float* o = dgsum + (size_t)(b0 + warp * RPT + i) * H4 + ucol;
#pragma unroll
            for (int g = 0; g < 4; ++g) o[g * CH] = gsum[g][i];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TCOLS>(tmem_d);
}

template <int NB, bool X3>
static int launch_bwd(const float* dh_all, const float* dh_last, const float* W_hh, const float* c_all,
                      const float* acts, float* dgates, float* dgsum, float* xchg, int T, int B, cudaStream_t st) {
    using S = BwdSmem<NB, X3>;
    static bool attr = false;
    auto kern = lstm_bwd_cluster_kernel<NB, X3, CNT>;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) {
            set_error("lstm_bwd_cluster: cudaFuncSetAttribute(%d B): %s", S::TOTAL, cudaGetErrorString(e));
            return (int)e;
        }
        attr = true;
    }
    kern<<<(B / NB) * CL, CNT, S::TOTAL, st>>>(dh_all, dh_last, W_hh, c_all, acts, dgates, dgsum, xchg, T, B);
    FHVAE_LAUNCH_CHECK("lstm_bwd_cluster");
    return 0;
}

int lstm_fwd_simt(const float* P, const float* Q, const float* W_hh, float* h_all, float* c_all,
                  float* acts, int T, int B, int H, cudaStream_t st);

template <int NB, bool X3>
static int launch_fwd(const float* P, const float* Q, const float* W_hh, float* h_all, float* c_all,
                      float* acts, float* xchg, int T, int B, cudaStream_t st) {
    using S = FwdSmem<NB, X3>;
    static bool attr = false;
    auto kern = lstm_fwd_cluster_kernel<NB, X3, CNT>;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) {
            set_error("lstm_fwd_cluster: cudaFuncSetAttribute(%d B): %s", S::TOTAL, cudaGetErrorString(e));
            return (int)e;
        }
        attr = true;
    }
    kern<<<(B / NB) * CL, CNT, S::TOTAL, st>>>(P, Q, W_hh, h_all, c_all, acts, xchg, T, B);
    FHVAE_LAUNCH_CHECK("lstm_fwd_cluster");
    return 0;
}

bool lstm_cluster_supported(int B, int H) { return H == CH && B % 32 == 0; }

// NB = 16 spreads the batch over twice as many clusters (halves the per-SM pointwise + DSMEM work);
// worth it only if all B/16 clusters of 8 CTAs are co-resident (16 clusters for B = 256).
template <int NB, bool X3>
static int max_clusters_fwd() {
    using S = FwdSmem<NB, X3>;
    auto kern = lstm_fwd_cluster_kernel<NB, X3, CNT>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CL * 32);
    cfg.blockDim = dim3(CNT);
    cfg.dynamicSmemBytes = S::TOTAL;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
static int pick_nb(int B) {
    static int max16 = -1;
    if (max16 < 0) {
        max16 = max_clusters_fwd<16, true>();
        const char* e = getenv("FHVAE_LSTM_NB");
        if (getenv("FHVAE_VERBOSE"))
            fprintf(stderr, "[fhvae] max co-resident clusters of %d CTAs (NB=16, bf16x3 fwd): %d; NB=32: %d\n", CL, max16,
                    max_clusters_fwd<32, true>());
        if (e && atoi(e) == 32) max16 = 0;
        if (e && atoi(e) == 16) max16 = 1 << 20;
    }
    return (B % 16 == 0 && B / 16 <= max16) ? 16 : 32;
}

int lstm_fwd_cluster(const float* P, const float* Q, const float* W_hh, float* h_all, float* c_all,
                     float* acts, float* xchg, int T, int B, int H, int mode, cudaStream_t st) {
    const bool x3 = (mode == FHVAE_MODE_BF16X3);
    if (pick_nb(B) == 16)
        return x3 ? launch_fwd<16, true>(P, Q, W_hh, h_all, c_all, acts, xchg, T, B, st)
                  : launch_fwd<16, false>(P, Q, W_hh, h_all, c_all, acts, xchg, T, B, st);
    return x3 ? launch_fwd<32, true>(P, Q, W_hh, h_all, c_all, acts, xchg, T, B, st)
              : launch_fwd<32, false>(P, Q, W_hh, h_all, c_all, acts, xchg, T, B, st);
}

int lstm_bwd_cluster(const float* dh_all, const float* dh_last, const float* W_hh, const float* c_all,
                     const float* acts, float* dgates, float* dgsum, float* xchg, int T, int B, int H, int mode,
                     cudaStream_t st) {
    const bool x3 = (mode == FHVAE_MODE_BF16X3);
    if (pick_nb(B) == 16)
        return x3 ? launch_bwd<16, true>(dh_all, dh_last, W_hh, c_all, acts, dgates, dgsum, xchg, T, B, st)
                  : launch_bwd<16, false>(dh_all, dh_last, W_hh, c_all, acts, dgates, dgsum, xchg, T, B, st);
    return x3 ? launch_bwd<32, true>(dh_all, dh_last, W_hh, c_all, acts, dgates, dgsum, xchg, T, B, st)
              : launch_bwd<32, false>(dh_all, dh_last, W_hh, c_all, acts, dgates, dgsum, xchg, T, B, st);
}

#ifdef FHVAE_TIMELINE
extern "C" int fhvae_debug_timeline(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, g_timeline, sizeof(g_timeline));
}
#endif

}  // namespace fhvae
