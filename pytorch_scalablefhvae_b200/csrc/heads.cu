// Latent "head" stages between the LSTM stacks, as single fp32 launches (GaussianLayer, simple_fhvae.py:205-216,
// plus the hoisted time-invariant input projection of the next stack).
//
// Between two recurrences the step is a chain of tiny dependent products on the critical path
//   head = [h_L0(T-1) | h_L1(T-1)] W^T + b   (B x 2Z, K = L*H)        GaussianLayer.mulayer / logvar_layer
//   z    = mu + eps * exp(logvar / 2)                                  GaussianLayer sample (:213-216)
//   Q    = z Wq^T (+ bq)                      (B x 4H, K = Z or Z1+Z2) time-invariant part of the next stack's
//                                                                      layer-0 input projection
// and the mirror image in the backward pass.  As tensor-core GEMM launches each link costs a launch gap plus
// a ~10 us latency-bound 128x128 tile for 8 MFLOP; here a CTA owns HB batch rows and walks the whole chain in
// exact fp32 (FFMA, fixed summation order => deterministic), with the weights streamed from L2.
#include "common.cuh"

namespace fhvae {

#ifndef HEAD_HB
#define HEAD_HB 2
#endif
constexpr int HB = HEAD_HB;      // batch rows per CTA (B = 256 -> 128 CTAs).  Whole step, config 1: HB = 4: +17 us, HB = 1: +65 us
constexpr int HT = 512;          // threads per CTA
constexpr int HMAXK = 1024;      // max L*H of a head / max 4H of dgsum staged in shared memory
constexpr int HMAXZ = 128;       // max 2Z, max Kq
constexpr int HTILE = 40960;     // floats of the weight-tile staging buffer (160 KB of dynamic shared memory)
constexpr int HBWD_SMEM = (HTILE + (HT / 32) * HB * HMAXZ) * 4;   // backward: + the cross-warp partial sums

// e / d and e % d for a CTA-uniform runtime divisor: a shift when d is a power of two (every size of the
// benchmark configuration is), the ~30-instruction integer division otherwise.  (Measured: not what bounds these
// kernels -- each phase is ~2-4k cycles of cold instruction fetch + one L2 round trip; tools/head_timeline.py.)
struct FastDiv {
    int d, lg;
    __device__ __forceinline__ explicit FastDiv(int d_) : d(d_), lg(-1) {
        if ((d_ & (d_ - 1)) == 0) {
            lg = 0;
            while ((1 << lg) < d_) ++lg;
        }
    }
    __device__ __forceinline__ int div(int e) const { return lg >= 0 ? (e >> lg) : e / d; }
    __device__ __forceinline__ int mod(int e) const { return lg >= 0 ? (e & (d - 1)) : e % d; }
};

// Cooperative copy of a rows x cols tile of W (leading dim ld) into shared memory (leading dim s_ld): every
// thread has all of its loads in flight before the first store (one exposed L2 round trip per tile instead of
// one per dependent load -- the scalar version of these kernels spent 30-50 us in serialized L2 latency).
__device__ __forceinline__ void stage_tile(const float* __restrict__ W, int64_t ld, int rows, int cols,
                                           float* __restrict__ s, int s_ld) {
    const bool vec = (cols % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
    const int c4 = cols >> 2;
    if (vec && c4 <= HT && HT % c4 == 0) {
        // thread -> fixed float4 column, rows r, r + HT/c4, ...: no index arithmetic in the loop
        const int rpp = HT / c4, c = threadIdx.x % c4, r_first = threadIdx.x / c4;
        const bool svec = (s_ld % 4 == 0);
        const float4* src = reinterpret_cast<const float4*>(W) + c;
        const int64_t ld4 = ld >> 2;
        for (int r0 = r_first; r0 < rows; r0 += rpp * 8) {
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (r0 + i * rpp < rows) v[i] = __ldcg(src + (int64_t)(r0 + i * rpp) * ld4);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = r0 + i * rpp;
                if (r < rows) {
                    float* d = s + r * s_ld + 4 * c;
                    if (svec) *reinterpret_cast<float4*>(d) = v[i];
                    else { d[0] = v[i].x; d[1] = v[i].y; d[2] = v[i].z; d[3] = v[i].w; }
                }
            }
        }
    } else {
        const int total = rows * cols;
        for (int e0 = threadIdx.x; e0 < total; e0 += HT * 8) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int e = e0 + i * HT;
                if (e < total) { const int r = e / cols, c = e - r * cols; v[i] = __ldcg(W + (int64_t)r * ld + c); }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int e = e0 + i * HT;
                if (e < total) { const int r = e / cols, c = e - r * cols; s[r * s_ld + c] = v[i]; }
            }
        }
    }
}

#ifdef FHVAE_TIMELINE
__device__ long long g_head_tl[2][16];
#define HTL(k, slot) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_head_tl[k][slot] = clock64(); } while (0)
extern "C" int fhvae_debug_head_timeline(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_head_tl, sizeof(g_head_tl)); }
#else
#define HTL(k, slot) do { } while (0)
#endif

struct HeadFwdArgs {
    const float* src[2]; int64_t ld_src; int nsrc, H;
    const float* W; const float* bias; float* head; int Z;
    const float* eps; float* zcat; int64_t ld_z; int zoff;
    const float* Wq; int64_t ld_wq; const float* bias_q; int qoff, Kq; float* Q; int NQ;
    int B;
};

__global__ void __launch_bounds__(HT) head_fwd_kernel(const __grid_constant__ HeadFwdArgs a) {
    extern __shared__ __align__(16) float s_w[];              // [HTILE] weight tile, rows padded (bank mapping below)
    __shared__ float s_h[HB][HMAXK];
    __shared__ float s_head[HB][HMAXZ];
    __shared__ float s_smp[HB][HMAXZ];
    __shared__ float s_z[HB][HMAXZ];
    const int tid = threadIdx.x;
    const int b0 = blockIdx.x * HB;
    const int K = a.nsrc * a.H, Z2 = 2 * a.Z;
    const FastDiv dK(K), dH(a.H), dZ(a.Z), dZ2(Z2), dKq(a.Kq > 0 ? a.Kq : 1);
    pdl_launch_dependents();
    pdl_wait();
    HTL(0, 0);
    // every independent global read of the CTA is issued up front (they overlap the first weight tile):
    // its HB rows of final hidden states, eps, and the part of the projection input another launch produced
    float hreg[HB * HMAXK / HT];
#pragma unroll
    for (int i = 0; i < HB * HMAXK / HT; ++i) {
        const int e = tid + i * HT;
        hreg[i] = 0.f;
        if (e < HB * K) {
            const int r = dK.div(e), k = dK.mod(e), l = dH.div(k), u = dH.mod(k);
            if (b0 + r < a.B) hreg[i] = __ldcg(a.src[l] + (int64_t)(b0 + r) * a.ld_src + u);
        }
    }
    float epsreg = 0.f, zreg = 0.f;
    if (a.eps && tid < HB * a.Z) {
        const int r = dZ.div(tid), d = dZ.mod(tid);
        if (b0 + r < a.B) epsreg = __ldcg(a.eps + (int64_t)(b0 + r) * a.Z + d);
    }
    if (a.Q && tid < HB * a.Kq) {
        const int r = dKq.div(tid), col = a.qoff + dKq.mod(tid);
        const bool own = a.eps && col >= a.zoff && col < a.zoff + a.Z;
        if (!own && b0 + r < a.B) zreg = a.zcat[(int64_t)(b0 + r) * a.ld_z + col];
    }
    // ---- head: tiles of TN outputs x K; TPO threads per output, thread p takes k = p, p + TPO, ...  Row stride
    // == TPO (mod 32): the (32/TPO outputs) x (TPO slices) of a warp hit 32 different banks; the slices are summed
    // by a fixed xor-shuffle tree.
    {
        int TPO = 8, lgT = 3;
        while (TPO > 1 && Z2 * TPO > HT) { TPO >>= 1; --lgT; }
        const int sld = (K + 31) / 32 * 32 + (TPO == 1 ? 1 : TPO);
        const int TN = min(Z2, HTILE / sld);
        for (int n0 = 0; n0 < Z2; n0 += TN) {
            const int tn = min(TN, Z2 - n0);
            __syncthreads();
            HTL(0, 1);
            stage_tile(a.W + (int64_t)n0 * K, K, tn, K, s_w, sld);
            if (n0 == 0) {
#pragma unroll
                for (int i = 0; i < HB * HMAXK / HT; ++i) {
                    const int e = tid + i * HT;
                    if (e < HB * K) s_h[dK.div(e)][dK.mod(e)] = hreg[i];
                }
            }
            __syncthreads();
            HTL(0, 2);
            for (int o0 = 0; o0 < TPO * tn; o0 += HT) {        // CTA-uniform trip count (shuffles below)
                const int o = o0 + tid, n = o >> lgT, p = o & (TPO - 1);
                const bool valid = o < TPO * tn;
                float acc[HB];
#pragma unroll
                for (int r = 0; r < HB; ++r) acc[r] = 0.f;
                if (valid) {
                    const float* w = s_w + n * sld;
#pragma unroll 8
                    for (int k = p; k < K; k += TPO) {
                        const float wv = w[k];
#pragma unroll
                        for (int r = 0; r < HB; ++r) acc[r] = fmaf(wv, s_h[r][k], acc[r]);
                    }
                }
                for (int d = 1; d < TPO; d <<= 1) {
#pragma unroll
                    for (int r = 0; r < HB; ++r) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], d);
                }
                if (valid && p == 0) {
                    const float bv = a.bias ? __ldcg(a.bias + n0 + n) : 0.f;
#pragma unroll
                    for (int r = 0; r < HB; ++r) s_head[r][n0 + n] = acc[r] + bv;
                }
            }
        }
    }
    __syncthreads();
    HTL(0, 3);
    for (int e = tid; e < HB * Z2; e += HT) {
        const int r = dZ2.div(e), n = dZ2.mod(e);
        if (b0 + r < a.B) a.head[(int64_t)(b0 + r) * Z2 + n] = s_head[r][n];
    }
    // ---- sample (same expression as reparam_fwd_kernel)
    if (a.eps && tid < HB * a.Z) {
        const int r = dZ.div(tid), d = dZ.mod(tid);
        const float z = fmaf(epsreg, expf(0.5f * s_head[r][a.Z + d]), s_head[r][d]);
        s_smp[r][d] = z;
        if (b0 + r < a.B) a.zcat[(int64_t)(b0 + r) * a.ld_z + a.zoff + d] = z;
    }
    if (!a.Q) return;
    __syncthreads();
    if (tid < HB * a.Kq) {
        const int r = dKq.div(tid), j = dKq.mod(tid), col = a.qoff + j;
        const bool own = a.eps && col >= a.zoff && col < a.zoff + a.Z;
        s_z[r][j] = own ? s_smp[r][col - a.zoff] : zreg;
    }
    // ---- Q: tiles of TQ output columns x Kq (odd row stride: conflict-free); thread -> output column
    const int sld = a.Kq | 1;
    int TQ = min(a.NQ, HTILE / sld);
    if (TQ >= HT) TQ = TQ / HT * HT;
    for (int n0 = 0; n0 < a.NQ; n0 += TQ) {
        const int tn = min(TQ, a.NQ - n0);
        __syncthreads();
        HTL(0, 4);
        stage_tile(a.Wq + (int64_t)n0 * a.ld_wq, a.ld_wq, tn, a.Kq, s_w, sld);
        __syncthreads();
        HTL(0, 5);
        for (int n = tid; n < tn; n += HT) {
            const float* w = s_w + n * sld;
            float acc[HB];
#pragma unroll
            for (int r = 0; r < HB; ++r) acc[r] = 0.f;
#pragma unroll 8
            for (int j = 0; j < a.Kq; ++j) {
                const float wv = w[j];
#pragma unroll
                for (int r = 0; r < HB; ++r) acc[r] = fmaf(wv, s_z[r][j], acc[r]);
            }
            const float bv = a.bias_q ? __ldcg(a.bias_q + n0 + n) : 0.f;
#pragma unroll
            for (int r = 0; r < HB; ++r)
                if (b0 + r < a.B) a.Q[(int64_t)(b0 + r) * a.NQ + n0 + n] = acc[r] + bv;
        }
        HTL(0, 6);
    }
}

struct HeadBwdArgs {
    const float* dgsum; int NG; const float* Wq; int64_t ld_wq; int Kq;
    float* dzcat; int64_t ld_dz; int dzoff; int beta;
    const float* head; const float* eps; int Z; int roff; float* dhead; int accumulate;
    const float* W; int nsrc, H; float* dh[2];
    int B;
};

__global__ void __launch_bounds__(HT) head_bwd_kernel(const __grid_constant__ HeadBwdArgs a) {
    extern __shared__ __align__(16) float s_w[];              // [HTILE] weight tile (rows contiguous: threads walk columns)
    __shared__ float s_g[HB][HMAXK];
    float (*s_part)[HB][HMAXZ] = reinterpret_cast<float (*)[HB][HMAXZ]>(s_w + HTILE);   // [HT / 32] cross-warp partials
    __shared__ float s_dhead[HB][HMAXZ];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = blockIdx.x * HB;
    const int Z2 = 2 * a.Z;
    const FastDiv dNG(a.NG > 0 ? a.NG : 1), dKq(a.Kq > 0 ? a.Kq : 1), dZ(a.Z), dZ2(Z2), dH(a.H > 0 ? a.H : 1);
    pdl_launch_dependents();
    pdl_wait();
    // ---- dz[b][j] = sum_n dgsum[b][n] Wq[n][j]: tiles of TR rows of Wq; lanes along j, the rows of a tile are
    // split across the warps, partial sums combined across warps in a fixed order
    if (a.dgsum) {
        for (int e = tid; e < HB * a.NG; e += HT) {
            const int r = dNG.div(e), n = dNG.mod(e);
            s_g[r][n] = (b0 + r < a.B) ? __ldcg(a.dgsum + (int64_t)(b0 + r) * a.NG + n) : 0.f;
        }
        const int TR = min(a.NG, HTILE / a.Kq / 8 * 8);
        float acc[HB][4];                                      // Kq <= 128: up to 4 columns per lane
#pragma unroll
        for (int r = 0; r < HB; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[r][q] = 0.f;
        for (int r0 = 0; r0 < a.NG; r0 += TR) {
            const int tr = min(TR, a.NG - r0);
            __syncthreads();
            stage_tile(a.Wq + (int64_t)r0 * a.ld_wq, a.ld_wq, tr, a.Kq, s_w, a.Kq);
            __syncthreads();
            const int rpw = (tr + HT / 32 - 1) / (HT / 32);
            const int n0 = warp * rpw, n1 = min(tr, n0 + rpw);
            for (int n = n0; n < n1; ++n) {
                const float* w = s_w + n * a.Kq;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = lane + 32 * q;
                    if (j < a.Kq) {
                        const float wv = w[j];
#pragma unroll
                        for (int r = 0; r < HB; ++r) acc[r][q] = fmaf(s_g[r][r0 + n], wv, acc[r][q]);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < HB; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (lane + 32 * q < a.Kq) s_part[warp][r][lane + 32 * q] = acc[r][q];
        __syncthreads();
        for (int e = tid; e < HB * a.Kq; e += HT) {
            const int r = dKq.div(e), j = dKq.mod(e);
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < HT / 32; ++w) s += s_part[w][r][j];
            if (b0 + r < a.B) {
                float* o = a.dzcat + (int64_t)(b0 + r) * a.ld_dz + a.dzoff + j;
                if (a.beta) s += *o;
                *o = s;
            }
        }
        __syncthreads();                   // dzcat writes of this CTA visible below
    }
    // ---- reparameterisation backward into dhead (same expressions as reparam_bwd_kernel)
    if (a.eps) {
        for (int e = tid; e < HB * a.Z; e += HT) {
            const int r = dZ.div(e), d = dZ.mod(e);
            if (b0 + r < a.B) {
                const int64_t b = b0 + r;
                const float lv = __ldcg(a.head + b * Z2 + a.Z + d);
                const float g = a.dzcat[b * a.ld_dz + a.roff + d];
                const float gmu = g, glv = g * 0.5f * __ldcg(a.eps + b * a.Z + d) * expf(0.5f * lv);
                float* pm = a.dhead + b * Z2 + d;
                float* pl = a.dhead + b * Z2 + a.Z + d;
                if (a.accumulate) { *pm += gmu; *pl += glv; } else { *pm = gmu; *pl = glv; }
            }
        }
        __syncthreads();
    }
    if (!a.W) return;
    for (int e = tid; e < HB * Z2; e += HT) {
        const int r = dZ2.div(e), j = dZ2.mod(e);
        s_dhead[r][j] = (b0 + r < a.B) ? a.dhead[(int64_t)(b0 + r) * Z2 + j] : 0.f;
    }
    // ---- dh_last of every layer: dh[l][b][u] = sum_j dhead[b][j] W[j][l*H + u]; tiles of TJ rows of W, threads along u
    const int K = a.nsrc * a.H;
    const int TJ = min(Z2, HTILE / K);
    float acc[HB][(HMAXK + HT - 1) / HT];
#pragma unroll
    for (int r = 0; r < HB; ++r)
#pragma unroll
        for (int q = 0; q < (HMAXK + HT - 1) / HT; ++q) acc[r][q] = 0.f;
    for (int j0 = 0; j0 < Z2; j0 += TJ) {
        const int tj = min(TJ, Z2 - j0);
        __syncthreads();
        stage_tile(a.W + (int64_t)j0 * K, K, tj, K, s_w, K);
        __syncthreads();
#pragma unroll
        for (int q = 0; q < (HMAXK + HT - 1) / HT; ++q) {
            const int k = tid + q * HT;
            if (k < K) {
#pragma unroll 8
                for (int j = 0; j < tj; ++j) {
                    const float wv = s_w[j * K + k];
#pragma unroll
                    for (int r = 0; r < HB; ++r) acc[r][q] = fmaf(s_dhead[r][j0 + j], wv, acc[r][q]);
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < (HMAXK + HT - 1) / HT; ++q) {
        const int k = tid + q * HT;
        if (k < K) {
            const int l = dH.div(k), u = dH.mod(k);
#pragma unroll
            for (int r = 0; r < HB; ++r)
                if (b0 + r < a.B) a.dh[l][(int64_t)(b0 + r) * a.H + u] = acc[r][q];
        }
    }
}

// coefficients of the ELBO backward from the upstream gradients of the six outputs (rows of gout:
// 0 lb, 1 log_px, 2 nk1, 3 nk2, 4 log_pmu2, 5 log_qy):  coef = [g_px, g_nk1, g_nk2, g_pmu2] per segment
__global__ void step_coef_kernel(const float* __restrict__ gout, const int64_t* __restrict__ nsegs,
                                 float* __restrict__ coef, int detach_px, int prior_grad, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float g0 = gout[b];
    coef[b] = detach_px ? 0.f : gout[B + b] + g0;
    coef[B + b] = gout[2 * B + b] + g0;
    coef[2 * B + b] = gout[3 * B + b] + g0;
    coef[3 * B + b] = prior_grad ? g0 / (float)nsegs[b] + gout[4 * B + b] : 0.f;
}

// loss = -(1/B) sum_b (lb[b] + alpha * log_qy[b])    (train_model.py:243-251); one CTA, fixed order
__global__ void __launch_bounds__(256) loss_mean_kernel(const float* lb, const float* lqy,
                                                        float alpha, int B, float* loss) {
    __shared__ float red[8];
    float s = 0.f;
    pdl_launch_dependents();
    pdl_wait();
    for (int b = threadIdx.x; b < B; b += 256) s += lb[b] + alpha * lqy[b];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w];
        *loss = -t / (float)B;
    }
}

}  // namespace fhvae

using namespace fhvae;

extern "C" int fhvae_head_fwd(const float* src0, const float* src1, int64_t ld_src, int nsrc, int H, const float* W,
                              const float* bias, float* head, int Z, const float* eps, float* zcat, int64_t ld_z,
                              int zoff, const float* Wq, int64_t ld_wq, const float* bias_q, int qoff, int Kq,
                              float* Q, int NQ, int B, void* stream) {
    FHVAE_CHECK_ARG(src0 && W && head && B > 0 && Z > 0 && H > 0, "head_fwd: bad argument");
    FHVAE_CHECK_ARG(nsrc == 1 || (nsrc == 2 && src1), "head_fwd: 1 or 2 sources");
    FHVAE_CHECK_SUP(nsrc * H <= HMAXK && 2 * Z <= HMAXZ, "head_fwd: needs L*H <= %d and 2Z <= %d", HMAXK, HMAXZ);
    FHVAE_CHECK_ARG(!eps || zcat, "head_fwd: eps without a sample destination");
    FHVAE_CHECK_ARG(!Q || (Wq && zcat && Kq > 0 && Kq <= HMAXZ && NQ > 0), "head_fwd: bad projection arguments");
    HeadFwdArgs a{{src0, src1}, ld_src, nsrc, H, W, bias, head, Z, eps, zcat, ld_z, zoff,
                  Wq, ld_wq, bias_q, qoff, Kq, Q, NQ, B};
    static bool attr_f = false;
    if (!attr_f) {
        cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HTILE * 4);
        attr_f = true;
    }
    launch_pdl(PDL_HEADS, head_fwd_kernel, dim3(cdiv(B, HB)), dim3(HT), HTILE * 4, as_stream(stream), a);
    FHVAE_LAUNCH_CHECK("head_fwd");
    return 0;
}

extern "C" int fhvae_head_bwd(const float* dgsum, int NG, const float* Wq, int64_t ld_wq, int Kq, float* dzcat,
                              int64_t ld_dz, int dzoff, int beta, const float* head, const float* eps, int Z,
                              int roff, float* dhead, int accumulate, const float* W, int nsrc, int H, float* dh0,
                              float* dh1, int B, void* stream) {
    FHVAE_CHECK_ARG(B > 0 && Z > 0 && 2 * Z <= HMAXZ, "head_bwd: bad size");
    FHVAE_CHECK_ARG(!dgsum || (Wq && dzcat && NG > 0 && Kq > 0), "head_bwd: bad dz arguments");
    FHVAE_CHECK_SUP(!dgsum || (NG <= HMAXK && Kq <= HMAXZ), "head_bwd: needs 4H <= %d and Kq <= %d", HMAXK, HMAXZ);
    FHVAE_CHECK_ARG(!eps || (head && dzcat && dhead), "head_bwd: bad reparam arguments");
    FHVAE_CHECK_ARG(!W || (dhead && dh0 && H > 0 && (nsrc == 1 || (nsrc == 2 && dh1))), "head_bwd: bad head arguments");
    HeadBwdArgs a{dgsum, NG, Wq, ld_wq, Kq, dzcat, ld_dz, dzoff, beta, head, eps, Z, roff, dhead, accumulate,
                  W, nsrc, H, {dh0, dh1}, B};
    static bool attr_b = false;
    if (!attr_b) {
        cudaFuncSetAttribute(head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HBWD_SMEM);
        attr_b = true;
    }
    launch_pdl(PDL_HEADS, head_bwd_kernel, dim3(cdiv(B, HB)), dim3(HT), HBWD_SMEM, as_stream(stream), a);
    FHVAE_LAUNCH_CHECK("head_bwd");
    return 0;
}

extern "C" int fhvae_step_coef(const float* gout, const int64_t* nsegs, float* coef, int detach_px, int prior_grad,
                               int B, void* stream) {
    FHVAE_CHECK_ARG(gout && nsegs && coef && B > 0, "step_coef: bad argument");
    step_coef_kernel<<<cdiv(B, 256), 256, 0, as_stream(stream)>>>(gout, nsegs, coef, detach_px, prior_grad, B);
    FHVAE_LAUNCH_CHECK("step_coef");
    return 0;
}

extern "C" int fhvae_loss_mean(const float* lb, const float* log_qy, float alpha, int B, float* loss, void* stream) {
    FHVAE_CHECK_ARG(lb && log_qy && loss && B > 0, "loss_mean: bad argument");
    launch_pdl(PDL_HEADS, loss_mean_kernel, dim3(1), dim3(256), 0, as_stream(stream), lb, log_qy, alpha, B, loss);
    FHVAE_LAUNCH_CHECK("loss_mean");
    return 0;
}
