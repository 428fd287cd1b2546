// K2/K4/K6: reparameterisation and the ELBO terms (simple_fhvae.py:56-69, :106-116, :213-216).
// HBM-bound: one CTA per segment, float4 coalesced loads, warp-shuffle reductions, the mu2 gather
// fused in.  Algorithmic bytes per segment: fwd 3*T*F*4 + small, bwd 5*T*F*4 + small.
#include "common.cuh"

namespace fhvae {

__global__ void reparam_fwd_kernel(const float* __restrict__ head, int64_t ld_head,
                                   const float* __restrict__ eps, float* __restrict__ sample,
                                   int64_t ld_s, int B, int Z) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * Z) return;
    const int b = i / Z, d = i % Z;
    const float mu = head[b * ld_head + d], lv = head[b * ld_head + Z + d];
    sample[b * ld_s + d] = fmaf(eps[i], expf(0.5f * lv), mu);
}

__global__ void reparam_bwd_kernel(const float* __restrict__ head, int64_t ld_head,
                                   const float* __restrict__ eps, const float* __restrict__ ds,
                                   int64_t ld_ds, float* __restrict__ dhead, int64_t ld_dh,
                                   int accumulate, int B, int Z) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * Z) return;
    const int b = i / Z, d = i % Z;
    const float lv = head[b * ld_head + Z + d];
    const float g = ds[b * ld_ds + d];
    const float gmu = g, glv = g * 0.5f * eps[i] * expf(0.5f * lv);
    float* pm = dhead + b * ld_dh + d;
    float* pl = dhead + b * ld_dh + Z + d;
    if (accumulate) { *pm += gmu; *pl += glv; } else { *pm = gmu; *pl = glv; }
}

constexpr int ELBO_THREADS = 128;

__device__ __forceinline__ float block_sum_128(float v, float* red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) red[w] = v;
    __syncthreads();
    float r = red[0] + red[1] + red[2] + red[3];
    __syncthreads();
    return r;
}

template <bool VEC>
__global__ void __launch_bounds__(ELBO_THREADS) elbo_fwd_kernel(
    const float* __restrict__ x, const float* __restrict__ xhead, int64_t xs_b, int64_t xs_t,
    int64_t lv_off, const float* __restrict__ z1head, const float* __restrict__ z2head,
    const float* __restrict__ mu2, const int64_t* __restrict__ nsegs, float* __restrict__ out5, int* __restrict__ nan_flag,
    int B, int T, int F, int Z1, int Z2) {
    __shared__ float red[4];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int TF = T * F;
    const float* xb = x + (int64_t)b * TF;
    const float* hb = xhead + (int64_t)b * xs_b;
    float s = 0.f;
    if (VEC) {
        const int F4 = F >> 2;
        for (int e = tid; e < (TF >> 2); e += ELBO_THREADS) {
            const int t = e / F4, f = (e - t * F4) << 2;
            const float4 xv = __ldcg(reinterpret_cast<const float4*>(xb + t * F + f));
            const float4 mv = __ldcg(reinterpret_cast<const float4*>(hb + t * xs_t + f));
            const float4 lv = __ldcg(reinterpret_cast<const float4*>(hb + t * xs_t + lv_off + f));
            float d;
            d = xv.x - mv.x; s += kLog2Pi + lv.x + d * d * expf(-lv.x);
            d = xv.y - mv.y; s += kLog2Pi + lv.y + d * d * expf(-lv.y);
            d = xv.z - mv.z; s += kLog2Pi + lv.z + d * d * expf(-lv.z);
            d = xv.w - mv.w; s += kLog2Pi + lv.w + d * d * expf(-lv.w);
        }
    } else {
        for (int e = tid; e < TF; e += ELBO_THREADS) {
            const int t = e / F, f = e - t * F;
            const float xv = xb[e], mv = hb[t * xs_t + f], lv = hb[t * xs_t + lv_off + f];
            const float d = xv - mv;
            s += kLog2Pi + lv + d * d * expf(-lv);
        }
    }
    const float log_px = -0.5f * block_sum_128(s, red);
    if (tid < 32) {
        float k1 = 0.f, k2 = 0.f, pm = 0.f;
        for (int d = tid; d < Z1; d += 32) {
            const float mu = z1head[(int64_t)b * 2 * Z1 + d], lv = z1head[(int64_t)b * 2 * Z1 + Z1 + d];
            k1 += 1.f + lv - (mu * mu + expf(lv));
        }
        for (int d = tid; d < Z2; d += 32) {
            const float mu = z2head[(int64_t)b * 2 * Z2 + d], lv = z2head[(int64_t)b * 2 * Z2 + Z2 + d];
            const float m2 = __ldcg(mu2 + (int64_t)b * Z2 + d);
            const float dm = mu - m2;
            k2 += 1.f + lv - kPz2Logvar - (dm * dm + expf(lv)) * kInvS2;
            pm += kLog2Pi + m2 * m2;
        }
        k1 = 0.5f * warp_sum(k1);
        k2 = 0.5f * warp_sum(k2);
        pm = -0.5f * warp_sum(pm);
        if (tid == 0) {
            const float lb = log_px + k1 + k2 + pm / (float)nsegs[b];
            out5[0 * B + b] = lb;
            out5[1 * B + b] = log_px;
            out5[2 * B + b] = k1;
            out5[3 * B + b] = k2;
            out5[4 * B + b] = pm;
            if (nan_flag && isnan(lb)) atomicOr(nan_flag, FHVAE_FLAG_NAN);
        }
    }
}

template <bool VEC>
__global__ void __launch_bounds__(ELBO_THREADS) elbo_bwd_kernel(
    const float* __restrict__ x, const float* __restrict__ xhead, int64_t xs_b, int64_t xs_t,
    int64_t lv_off, const float* __restrict__ z1head, const float* __restrict__ z2head,
    const float* __restrict__ mu2, const float* __restrict__ coef, float* __restrict__ dxhead, float* __restrict__ dz1head,
    float* __restrict__ dz2head, float* __restrict__ dmu2, int B, int T, int F, int Z1, int Z2) {
    const int b = blockIdx.x, tid = threadIdx.x;
    const int TF = T * F;
    const float c_px = coef[0 * B + b], c_k1 = coef[1 * B + b], c_k2 = coef[2 * B + b],
                c_pm = coef[3 * B + b];
    const float* xb = x + (int64_t)b * TF;
    const float* hb = xhead + (int64_t)b * xs_b;
    float* db = dxhead + (int64_t)b * xs_b;
    if (VEC) {
        const int F4 = F >> 2;
        for (int e = tid; e < (TF >> 2); e += ELBO_THREADS) {
            const int t = e / F4, f = (e - t * F4) << 2;
            const float4 xv = __ldcg(reinterpret_cast<const float4*>(xb + t * F + f));
            const float4 mv = __ldcg(reinterpret_cast<const float4*>(hb + t * xs_t + f));
            const float4 lv = __ldcg(reinterpret_cast<const float4*>(hb + t * xs_t + lv_off + f));
            float4 gm, gl;
            float d, iv;
            d = xv.x - mv.x; iv = expf(-lv.x); gm.x = c_px * d * iv; gl.x = -0.5f * c_px * (1.f - d * d * iv);
            d = xv.y - mv.y; iv = expf(-lv.y); gm.y = c_px * d * iv; gl.y = -0.5f * c_px * (1.f - d * d * iv);
            d = xv.z - mv.z; iv = expf(-lv.z); gm.z = c_px * d * iv; gl.z = -0.5f * c_px * (1.f - d * d * iv);
            d = xv.w - mv.w; iv = expf(-lv.w); gm.w = c_px * d * iv; gl.w = -0.5f * c_px * (1.f - d * d * iv);
            *reinterpret_cast<float4*>(db + t * xs_t + f) = gm;
            *reinterpret_cast<float4*>(db + t * xs_t + lv_off + f) = gl;
        }
    } else {
        for (int e = tid; e < TF; e += ELBO_THREADS) {
            const int t = e / F, f = e - t * F;
            const float xv = xb[e], mv = hb[t * xs_t + f], lv = hb[t * xs_t + lv_off + f];
            const float d = xv - mv, iv = expf(-lv);
            db[t * xs_t + f] = c_px * d * iv;
            db[t * xs_t + lv_off + f] = -0.5f * c_px * (1.f - d * d * iv);
        }
    }
    for (int d = tid; d < Z1; d += ELBO_THREADS) {
        const float mu = z1head[(int64_t)b * 2 * Z1 + d], lv = z1head[(int64_t)b * 2 * Z1 + Z1 + d];
        dz1head[(int64_t)b * 2 * Z1 + d] = -c_k1 * mu;
        dz1head[(int64_t)b * 2 * Z1 + Z1 + d] = 0.5f * c_k1 * (1.f - expf(lv));
    }
    for (int d = tid; d < Z2; d += ELBO_THREADS) {
        const float mu = z2head[(int64_t)b * 2 * Z2 + d], lv = z2head[(int64_t)b * 2 * Z2 + Z2 + d];
        const float m2 = __ldcg(mu2 + (int64_t)b * Z2 + d);
        const float dm = mu - m2;
        dz2head[(int64_t)b * 2 * Z2 + d] = -c_k2 * dm * kInvS2;
        dz2head[(int64_t)b * 2 * Z2 + Z2 + d] = 0.5f * c_k2 * (1.f - kInvS2 * expf(lv));
        dmu2[(int64_t)b * Z2 + d] = c_k2 * dm * kInvS2 - c_pm * m2;
    }
}

// Forward + backward of the ELBO terms in ONE pass, for a step whose upstream gradients are known before the
// forward runs (the fused train step: gout is a constant of (alpha, B)).  x, x_mu, x_logvar are read once instead of
// twice and the chain elbo_fwd -> step_coef -> elbo_bwd (three launches on the critical path) becomes one.  Same
// expressions in the same order as elbo_fwd_kernel / step_coef_kernel / elbo_bwd_kernel: bit-identical results.
template <bool VEC>
__global__ void __launch_bounds__(ELBO_THREADS) elbo_fwdbwd_kernel(
    const float* x, const float* xhead, int64_t xs_b, int64_t xs_t, int64_t lv_off,
    const float* z1head, const float* z2head, const float* mu2,
    const int64_t* nsegs, const float* gout, int detach_px, int prior_grad,
    float* out5, int* nan_flag, float* dxhead, float* dz1head,
    float* dz2head, float* dmu2, int B, int T, int F, int Z1, int Z2) {
    __shared__ float red[4];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int TF = T * F;
    pdl_launch_dependents();
    pdl_wait();
    const float g0 = gout[b];
    const float c_px = detach_px ? 0.f : gout[B + b] + g0;
    const float c_k1 = gout[2 * B + b] + g0, c_k2 = gout[3 * B + b] + g0;
    const float c_pm = prior_grad ? g0 / (float)nsegs[b] + gout[4 * B + b] : 0.f;
    const float* xb = x + (int64_t)b * TF;
    const float* hb = xhead + (int64_t)b * xs_b;
    float* db = dxhead + (int64_t)b * xs_b;
    float s = 0.f;
    if (VEC) {
        const int F4 = F >> 2;
        for (int e = tid; e < (TF >> 2); e += ELBO_THREADS) {
            const int t = e / F4, f = (e - t * F4) << 2;
            const float4 xv = __ldcg(reinterpret_cast<const float4*>(xb + t * F + f));
            const float4 mv = __ldcg(reinterpret_cast<const float4*>(hb + t * xs_t + f));
            const float4 lv = __ldcg(reinterpret_cast<const float4*>(hb + t * xs_t + lv_off + f));
            float4 gm, gl;
            float d, iv;
            d = xv.x - mv.x; iv = expf(-lv.x); s += kLog2Pi + lv.x + d * d * iv; gm.x = c_px * d * iv; gl.x = -0.5f * c_px * (1.f - d * d * iv);
            d = xv.y - mv.y; iv = expf(-lv.y); s += kLog2Pi + lv.y + d * d * iv; gm.y = c_px * d * iv; gl.y = -0.5f * c_px * (1.f - d * d * iv);
            d = xv.z - mv.z; iv = expf(-lv.z); s += kLog2Pi + lv.z + d * d * iv; gm.z = c_px * d * iv; gl.z = -0.5f * c_px * (1.f - d * d * iv);
            d = xv.w - mv.w; iv = expf(-lv.w); s += kLog2Pi + lv.w + d * d * iv; gm.w = c_px * d * iv; gl.w = -0.5f * c_px * (1.f - d * d * iv);
            *reinterpret_cast<float4*>(db + t * xs_t + f) = gm;
            *reinterpret_cast<float4*>(db + t * xs_t + lv_off + f) = gl;
        }
    } else {
        for (int e = tid; e < TF; e += ELBO_THREADS) {
            const int t = e / F, f = e - t * F;
            const float xv = xb[e], mv = hb[t * xs_t + f], lv = hb[t * xs_t + lv_off + f];
            const float d = xv - mv, iv = expf(-lv);
            s += kLog2Pi + lv + d * d * iv;
            db[t * xs_t + f] = c_px * d * iv;
            db[t * xs_t + lv_off + f] = -0.5f * c_px * (1.f - d * d * iv);
        }
    }
    const float log_px = -0.5f * block_sum_128(s, red);
    if (tid < 32) {
        float k1 = 0.f, k2 = 0.f, pm = 0.f;
        for (int d = tid; d < Z1; d += 32) {
            const float mu = z1head[(int64_t)b * 2 * Z1 + d], lv = z1head[(int64_t)b * 2 * Z1 + Z1 + d];
            k1 += 1.f + lv - (mu * mu + expf(lv));
        }
        for (int d = tid; d < Z2; d += 32) {
            const float mu = z2head[(int64_t)b * 2 * Z2 + d], lv = z2head[(int64_t)b * 2 * Z2 + Z2 + d];
            const float m2 = __ldcg(mu2 + (int64_t)b * Z2 + d);
            const float dm = mu - m2;
            k2 += 1.f + lv - kPz2Logvar - (dm * dm + expf(lv)) * kInvS2;
            pm += kLog2Pi + m2 * m2;
        }
        k1 = 0.5f * warp_sum(k1);
        k2 = 0.5f * warp_sum(k2);
        pm = -0.5f * warp_sum(pm);
        if (tid == 0) {
            const float lb = log_px + k1 + k2 + pm / (float)nsegs[b];
            out5[0 * B + b] = lb;
            out5[1 * B + b] = log_px;
            out5[2 * B + b] = k1;
            out5[3 * B + b] = k2;
            out5[4 * B + b] = pm;
            if (nan_flag && isnan(lb)) atomicOr(nan_flag, FHVAE_FLAG_NAN);
        }
    }
    for (int d = tid; d < Z1; d += ELBO_THREADS) {
        const float mu = z1head[(int64_t)b * 2 * Z1 + d], lv = z1head[(int64_t)b * 2 * Z1 + Z1 + d];
        dz1head[(int64_t)b * 2 * Z1 + d] = -c_k1 * mu;
        dz1head[(int64_t)b * 2 * Z1 + Z1 + d] = 0.5f * c_k1 * (1.f - expf(lv));
    }
    for (int d = tid; d < Z2; d += ELBO_THREADS) {
        const float mu = z2head[(int64_t)b * 2 * Z2 + d], lv = z2head[(int64_t)b * 2 * Z2 + Z2 + d];
        const float m2 = __ldcg(mu2 + (int64_t)b * Z2 + d);
        const float dm = mu - m2;
        dz2head[(int64_t)b * 2 * Z2 + d] = -c_k2 * dm * kInvS2;
        dz2head[(int64_t)b * 2 * Z2 + Z2 + d] = 0.5f * c_k2 * (1.f - kInvS2 * expf(lv));
        dmu2[(int64_t)b * Z2 + d] = c_k2 * dm * kInvS2 - c_pm * m2;
    }
}

static bool vec_ok(const void* a, const void* b, const void* c, int F, int64_t xs_b, int64_t xs_t,
                   int64_t lv_off) {
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return al(a) && al(b) && (c == nullptr || al(c)) && F % 4 == 0 && xs_b % 4 == 0 && xs_t % 4 == 0 &&
           lv_off % 4 == 0;
}

}  // namespace fhvae

using namespace fhvae;

extern "C" int fhvae_reparam_fwd(const float* head, int64_t ld_head, const float* eps, float* sample,
                                 int64_t ld_s, int B, int Z, void* stream) {
    FHVAE_CHECK_ARG(head && eps && sample && B > 0 && Z > 0, "reparam_fwd: bad argument");
    reparam_fwd_kernel<<<cdiv((int64_t)B * Z, 256), 256, 0, as_stream(stream)>>>(head, ld_head, eps,
                                                                                 sample, ld_s, B, Z);
    FHVAE_LAUNCH_CHECK("reparam_fwd");
    return 0;
}

extern "C" int fhvae_reparam_bwd(const float* head, int64_t ld_head, const float* eps,
                                 const float* dsample, int64_t ld_ds, float* dhead, int64_t ld_dh,
                                 int accumulate, int B, int Z, void* stream) {
    FHVAE_CHECK_ARG(head && eps && dsample && dhead && B > 0 && Z > 0, "reparam_bwd: bad argument");
    reparam_bwd_kernel<<<cdiv((int64_t)B * Z, 256), 256, 0, as_stream(stream)>>>(
        head, ld_head, eps, dsample, ld_ds, dhead, ld_dh, accumulate, B, Z);
    FHVAE_LAUNCH_CHECK("reparam_bwd");
    return 0;
}

extern "C" int fhvae_elbo_fwd(const float* x, const float* xhead, int64_t xs_b, int64_t xs_t,
                              int64_t lv_off, const float* z1head, const float* z2head,
                              const float* mu2, const int64_t* nsegs,
                              float* out5, int* nan_flag, int B, int T, int F, int Z1, int Z2,
                              void* stream) {
    FHVAE_CHECK_ARG(x && xhead && z1head && z2head && mu2 && nsegs && out5, "elbo_fwd: null pointer");
    FHVAE_CHECK_ARG(B > 0 && T > 0 && F > 0 && Z1 > 0 && Z2 > 0, "elbo_fwd: bad size");
    if (vec_ok(x, xhead, nullptr, F, xs_b, xs_t, lv_off))
        elbo_fwd_kernel<true><<<B, ELBO_THREADS, 0, as_stream(stream)>>>(
            x, xhead, xs_b, xs_t, lv_off, z1head, z2head, mu2, nsegs, out5, nan_flag, B, T, F, Z1, Z2);
    else
        elbo_fwd_kernel<false><<<B, ELBO_THREADS, 0, as_stream(stream)>>>(
            x, xhead, xs_b, xs_t, lv_off, z1head, z2head, mu2, nsegs, out5, nan_flag, B, T, F, Z1, Z2);
    FHVAE_LAUNCH_CHECK("elbo_fwd");
    return 0;
}

extern "C" int fhvae_elbo_bwd(const float* x, const float* xhead, int64_t xs_b, int64_t xs_t,
                              int64_t lv_off, const float* z1head, const float* z2head,
                              const float* mu2, const float* coef,
                              float* dxhead, float* dz1head, float* dz2head, float* dmu2, int B,
                              int T, int F, int Z1, int Z2, void* stream) {
    FHVAE_CHECK_ARG(x && xhead && z1head && z2head && mu2 && coef && dxhead && dz1head && dz2head && dmu2,
                    "elbo_bwd: null pointer");
    FHVAE_CHECK_ARG(B > 0 && T > 0 && F > 0 && Z1 > 0 && Z2 > 0, "elbo_bwd: bad size");
    if (vec_ok(x, xhead, dxhead, F, xs_b, xs_t, lv_off))
        elbo_bwd_kernel<true><<<B, ELBO_THREADS, 0, as_stream(stream)>>>(
            x, xhead, xs_b, xs_t, lv_off, z1head, z2head, mu2, coef, dxhead, dz1head, dz2head,
            dmu2, B, T, F, Z1, Z2);
    else
        elbo_bwd_kernel<false><<<B, ELBO_THREADS, 0, as_stream(stream)>>>(
            x, xhead, xs_b, xs_t, lv_off, z1head, z2head, mu2, coef, dxhead, dz1head, dz2head,
            dmu2, B, T, F, Z1, Z2);
    FHVAE_LAUNCH_CHECK("elbo_bwd");
    return 0;
}

extern "C" int fhvae_elbo_fwd_bwd(const float* x, const float* xhead, int64_t xs_b, int64_t xs_t, int64_t lv_off,
                                  const float* z1head, const float* z2head, const float* mu2, const int64_t* nsegs,
                                  const float* gout, int detach_px, int prior_grad, float* out5, int* nan_flag,
                                  float* dxhead, float* dz1head, float* dz2head, float* dmu2, int B, int T, int F,
                                  int Z1, int Z2, void* stream) {
    FHVAE_CHECK_ARG(x && xhead && z1head && z2head && mu2 && nsegs && gout && out5 && dxhead && dz1head && dz2head && dmu2,
                    "elbo_fwd_bwd: null pointer");
    FHVAE_CHECK_ARG(B > 0 && T > 0 && F > 0 && Z1 > 0 && Z2 > 0, "elbo_fwd_bwd: bad size");
    if (vec_ok(x, xhead, dxhead, F, xs_b, xs_t, lv_off))
        launch_pdl(PDL_ELBO, elbo_fwdbwd_kernel<true>, dim3(B), dim3(ELBO_THREADS), 0, as_stream(stream),
            x, xhead, xs_b, xs_t, lv_off, z1head, z2head, mu2, nsegs, gout, detach_px, prior_grad, out5, nan_flag, dxhead,
            dz1head, dz2head, dmu2, B, T, F, Z1, Z2);
    else
        launch_pdl(PDL_ELBO, elbo_fwdbwd_kernel<false>, dim3(B), dim3(ELBO_THREADS), 0, as_stream(stream),
            x, xhead, xs_b, xs_t, lv_off, z1head, z2head, mu2, nsegs, gout, detach_px, prior_grad, out5, nan_flag, dxhead,
            dz1head, dz2head, dmu2, B, T, F, Z1, Z2);
    FHVAE_LAUNCH_CHECK("elbo_fwd_bwd");
    return 0;
}
