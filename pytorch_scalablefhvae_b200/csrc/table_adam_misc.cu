// K0 (mu2 table gather / deterministic scatter-reduce / MAP estimate / sparse row copies),
// K9 (flat Adam) and the small data-movement helpers.  All HBM-/latency-bound byte shuffling:
// coalesced, vectorised where alignment allows, deterministic (no float atomics).
#include <math.h>
#include "common.cuh"

namespace fhvae {

// ---------------------------------------------------------------- K0: table ----------------------
__global__ void mu2_gather_kernel(const float* __restrict__ table, const int64_t* __restrict__ idx,
                                  float* __restrict__ mu2, int B, int Z, int64_t N,
                                  int32_t* __restrict__ err_flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * Z) return;
    const int b = i / Z, d = i % Z;
    const int64_t row = idx[b];
    if (row < 0 || row >= N) {
        // torch.gather raises on an out-of-range row (simple_fhvae.py:53); a device-resident idx cannot raise
        // without a host sync, so the row is poisoned (-> NaN lower bound, train_model.py:464 guard) and flagged
        mu2[i] = __int_as_float(0x7fc00000);
        if (err_flag && d == 0) atomicOr(err_flag, FHVAE_FLAG_BAD_INDEX);
        return;
    }
    mu2[i] = __ldg(table + row * Z + d);           // torch.gather(table, 0, idx), simple_fhvae.py:53
}

// dst[idx[b]] += sum of src rows with the same idx, summed in ascending b by the first occurrence.
// One warp per segment b; lanes stride over d.  Duplicates inside a batch are the norm (several
// segments of one utterance), float atomics would make the update order- and run-dependent.
__global__ void scatter_reduce_kernel(const float* __restrict__ src, int64_t ld_src,
                                      const int64_t* __restrict__ idx, float* __restrict__ dst,
                                      float* __restrict__ cnt, int32_t* __restrict__ touched, int B,
                                      int Z, int64_t N, int32_t* __restrict__ err_flag) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), l = threadIdx.x & 31;
    if (b >= B) return;
    const int64_t row = idx[b];
    if (row < 0 || row >= N) {                            // never write outside the table (see mu2_gather_kernel)
        if (l == 0) {
            if (touched) touched[b] = 0;
            if (err_flag) atomicOr(err_flag, FHVAE_FLAG_BAD_INDEX);
        }
        return;
    }
    // warp-parallel match of idx[] against this segment's row, 32 segments per ballot.  An earlier segment
    // with the same row owns the sum (first occurrence); later ones are added in ascending b, so the
    // summation order -- hence the result -- is exactly that of a serial scan.
    bool dup = false;
    for (int j0 = 0; j0 < b && !dup; j0 += 32) {
        const int j = j0 + l;
        const bool hit = (j < b) && (__ldg(idx + j) == row);
        dup = __ballot_sync(0xffffffffu, hit) != 0u;
    }
    if (touched && l == 0) touched[b] = dup ? 0 : 1;
    if (dup) return;
    float n = 1.f;
    float s[4];                                           // Z <= 128 handled in registers, larger Z loops below
    const int nd = (Z + 31) >> 5;
    if (nd <= 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] = (i < nd && l + 32 * i < Z) ? src[(int64_t)b * ld_src + l + 32 * i] : 0.f;
        for (int j0 = (b + 1) & ~31; j0 < B; j0 += 32) {
            const int j = j0 + l;
            unsigned m = __ballot_sync(0xffffffffu, j > b && j < B && __ldg(idx + j) == row);
            while (m) {
                const int jj = j0 + __ffs(m) - 1;
                m &= m - 1;
                n += 1.f;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i < nd && l + 32 * i < Z) s[i] += src[(int64_t)jj * ld_src + l + 32 * i];
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < nd && l + 32 * i < Z) dst[row * Z + l + 32 * i] += s[i];
    } else {
        for (int d = l; d < Z; d += 32) {
            float t = src[(int64_t)b * ld_src + d];
            for (int j = b + 1; j < B; ++j)
                if (__ldg(idx + j) == row) t += src[(int64_t)j * ld_src + d];
            dst[row * Z + d] += t;
        }
        if (l == 0)
            for (int j = b + 1; j < B; ++j)
                if (__ldg(idx + j) == row) n += 1.f;
    }
    if (cnt && l == 0) cnt[row] += n;
}

__global__ void mu2_estimate_finish_kernel(const float* __restrict__ zsum, const float* __restrict__ cnt,
                                           float* __restrict__ table, float r, int64_t K, int Z) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * Z) return;
    const float n = cnt[i / Z];
    if (n > 0.f) table[i] = zsum[i] / (n + r);     // utils.py:58-59
}

__global__ void rows_copy_kernel(const float* __restrict__ src, const int64_t* __restrict__ src_rows,
                                 float* __restrict__ dst, const int64_t* __restrict__ dst_rows,
                                 int64_t n, int Z) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * Z) return;
    const int64_t r = i / Z;
    const int d = (int)(i % Z);
    const int64_t dr = dst_rows ? dst_rows[r] : r;
    const int64_t sr = src_rows ? src_rows[r] : r;
    if (dr < 0 || sr < 0) return;
    dst[dr * Z + d] = src[sr * Z + d];
}

// ---------------------------------------------------------------- K9: Adam -----------------------
// torch.optim.Adam (no weight decay / amsgrad): m,v EMA; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps).
__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const float* g,
                                                        float* __restrict__ m, float* __restrict__ v,
                                                        int64_t n, float lr, float beta1, float beta2,
                                                        float eps, float gscale, int32_t* step,
                                                        uint32_t* done) {
    __shared__ float s_ss, s_bc2;
    pdl_launch_dependents();
    pdl_wait();
    const int t = *reinterpret_cast<volatile int32_t*>(step) + 1;
    if (threadIdx.x == 0) {
        const double bc1 = 1.0 - pow((double)beta1, (double)t);
        const double bc2 = 1.0 - pow((double)beta2, (double)t);
        s_ss = (float)((double)lr / bc1);
        s_bc2 = (float)sqrt(bc2);
    }
    __syncthreads();
    const float ss = s_ss, bc2s = s_bc2;
    const float ob1 = 1.f - beta1, ob2 = 1.f - beta2;
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pv = reinterpret_cast<float4*>(p)[i];
        float4 gv = __ldcg(reinterpret_cast<const float4*>(g) + i);   // (not ld.global.nc: see pdl_wait)
        float4 mv = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
#define ADAM1(c)                                             \
        {                                                    \
            const float gg = gv.c * gscale;                  \
            mv.c = fmaf(beta1, mv.c, ob1 * gg);              \
            vv.c = fmaf(beta2, vv.c, ob2 * gg * gg);         \
            pv.c -= ss * mv.c / (sqrtf(vv.c) / bc2s + eps);  \
        }
        ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
        reinterpret_cast<float4*>(p)[i] = pv;
        reinterpret_cast<float4*>(m)[i] = mv;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gg = g[i] * gscale;
        const float mm = fmaf(beta1, m[i], ob1 * gg);
        const float vv = fmaf(beta2, v[i], ob2 * gg * gg);
        m[i] = mm; v[i] = vv;
        p[i] -= ss * mm / (sqrtf(vv) / bc2s + eps);
    }
    // the last CTA to finish bumps the device step counter (every CTA has read it by then)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned prev = atomicInc(done, gridDim.x - 1);
        if (prev == gridDim.x - 1) *step = t;
    }
}

// ---------------------------------------------------------------- eps ~ N(0,1) -------------------
// torch.randn_like of GaussianLayer.forward (simple_fhvae.py:214) as a kernel of this library, so that the draw is a
// node of the step's CUDA graph instead of an ATen launch in front of it.  Philox4x32-10 (counter = element group,
// key = seed; the per-launch offset lives in a DEVICE counter that the last CTA bumps, so every replay of a captured
// launch draws fresh numbers) + Box-Muller.  Not bit-compatible with torch's generator (nothing in the reference pins
// the draws; parity tests inject eps).
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}
__global__ void __launch_bounds__(256) randn_kernel(float* __restrict__ out, int64_t n, uint64_t seed,
                                                    unsigned long long* __restrict__ offset, uint32_t* __restrict__ done) {
    const unsigned long long off = *reinterpret_cast<volatile unsigned long long*>(offset);
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;           // one group of 4 normals per thread
    if (g * 4 < n) {
        const unsigned long long c = off + (unsigned long long)g;
        const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u),
                                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        const float k = 2.3283064365386963e-10f;                                 // 2^-32
        const float u0 = ((float)r.x + 0.5f) * k, u1 = ((float)r.y + 0.5f) * k;
        const float u2 = ((float)r.z + 0.5f) * k, u3 = ((float)r.w + 0.5f) * k;
        const float ra = sqrtf(-2.f * logf(fminf(u0, 0.99999994f))), rb = sqrtf(-2.f * logf(fminf(u2, 0.99999994f)));
        float s0, c0, s1, c1;
        sincospif(2.f * u1, &s0, &c0);
        sincospif(2.f * u3, &s1, &c1);
        const float v[4] = {ra * c0, ra * s0, rb * c1, rb * s1};
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (g * 4 + i < n) out[g * 4 + i] = v[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {                      // the last CTA to finish advances the stream (every CTA has read it)
        __threadfence();
        if (atomicInc(done, gridDim.x - 1) == gridDim.x - 1) *offset = off + (unsigned long long)((n + 3) / 4);
    }
}

// ---------------------------------------------------------------- helpers ------------------------
__global__ void transpose_bt_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int T,
                                    int F) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // index into dst (T,B,F)
    if (i >= (int64_t)B * T * F) return;
    const int f = (int)(i % F);
    const int64_t tb = i / F;
    const int b = (int)(tb % B), t = (int)(tb / B);
    dst[i] = __ldg(src + ((int64_t)b * T + t) * F + f);
}

// the per-step input copies (x, mu_idx, num_segs -> the plan's static buffers) as ONE launch; x is also written in
// the time-major (T,B,F) layout the LSTM kernels read (x_tm may be NULL)
__global__ void __launch_bounds__(256) load_inputs_kernel(const float4* __restrict__ xs, float4* __restrict__ xd,
                                                          float4* __restrict__ xtm, int64_t n4, int B, int T, int F4,
                                                          const int64_t* __restrict__ is, int64_t* __restrict__ id,
                                                          const int64_t* __restrict__ ns, int64_t* __restrict__ nd) {
    const int64_t i0 = (int64_t)blockIdx.x * 256 + threadIdx.x;
    for (int64_t i = i0; i < n4; i += (int64_t)gridDim.x * 256) {
        const float4 v = __ldg(xs + i);
        xd[i] = v;
        if (xtm) {
            const int f = (int)(i % F4);
            const int64_t bt = i / F4;
            const int t = (int)(bt % T), b = (int)(bt / T);
            xtm[((int64_t)t * B + b) * F4 + f] = v;
        }
    }
    if (i0 < B) {
        if (is) id[i0] = is[i0];
        if (ns) nd[i0] = ns[i0];
    }
}

__global__ void gather_segments_kernel(const float* __restrict__ feats, const int64_t* __restrict__ start,
                                       const float* __restrict__ mean, const float* __restrict__ inv_std,
                                       float* __restrict__ out, int B, int T, int F, int64_t R) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * T * F) return;
    const int f = (int)(i % F);
    const int64_t bt = i / F;
    const int t = (int)(bt % T), b = (int)(bt / T);
    const int64_t row = start[b] + t;
    float v = (row >= 0 && row < R) ? __ldg(feats + row * F + f) : 0.f;
    if (mean) v = (v - __ldg(mean + f)) * __ldg(inv_std + f);
    out[i] = v;
}

// 32 columns per CTA, 32 warps stride over rows, fixed-order reduction over warps; grouped launch
struct ColsumBatch {
    fhvae_colsum_problem p[FHVAE_COLSUM_MAX_BATCH];
    int start[FHVAE_COLSUM_MAX_BATCH + 1];
    int n;
};
__global__ void __launch_bounds__(1024) colsum_kernel(const __grid_constant__ ColsumBatch cb) {
    __shared__ float red[32][33];
    int pi = 0;
    while (pi + 1 < cb.n && (int)blockIdx.x >= cb.start[pi + 1]) ++pi;
    const fhvae_colsum_problem& P = cb.p[pi];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const int c = (blockIdx.x - cb.start[pi]) * 32 + l;
    // warp w sums rows w, w+32, ...: 8 independent loads in flight per thread (the serial version was bound by
    // one L2 round trip per row: 43 us for the 5120-row decoder-head bias).  Fixed order => deterministic.
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (c < P.C) {
        const float* p = P.in + c;
        int r = w;
        for (; r + 7 * 32 < P.R; r += 8 * 32) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __ldg(p + (int64_t)(r + i * 32) * P.ld);
            s0 += v[0]; s1 += v[1]; s2 += v[2]; s3 += v[3];
            s0 += v[4]; s1 += v[5]; s2 += v[6]; s3 += v[7];
        }
        for (; r < P.R; r += 32) s0 += __ldg(p + (int64_t)r * P.ld);
    }
    red[w][l] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (w == 0 && c < P.C) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) t += red[k][l];
        P.out[c] = t;
        if (P.out2) P.out2[c] = t;
    }
}

__global__ void add2_kernel(float* __restrict__ out, const float* __restrict__ a,
                            const float* __restrict__ b, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + b[i];
}

__global__ void relu_bwd_kernel(float* __restrict__ dout, const float* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !(out[i] > 0.f)) dout[i] = 0.f;
}

__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float a, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = fmaf(a, x[i], y[i]);
}

}  // namespace fhvae

using namespace fhvae;

extern "C" int fhvae_mu2_gather(const float* table, const int64_t* idx, float* mu2, int B, int Z,
                                int64_t N, int32_t* err_flag, void* stream) {
    FHVAE_CHECK_ARG(table && idx && mu2 && B > 0 && Z > 0 && N > 0, "mu2_gather: bad argument");
    mu2_gather_kernel<<<cdiv((int64_t)B * Z, 256), 256, 0, as_stream(stream)>>>(table, idx, mu2, B, Z, N, err_flag);
    FHVAE_LAUNCH_CHECK("mu2_gather");
    return 0;
}

extern "C" int fhvae_mu2_scatter_reduce(const float* dmu2, const int64_t* idx, float* dtable,
                                        int32_t* touched, int B, int Z, int64_t N, void* stream) {
    FHVAE_CHECK_ARG(dmu2 && idx && dtable && B > 0 && Z > 0 && N > 0, "mu2_scatter_reduce: bad argument");
    scatter_reduce_kernel<<<cdiv(B, 8), 256, 0, as_stream(stream)>>>(dmu2, Z, idx, dtable, nullptr,
                                                                     touched, B, Z, N, nullptr);
    FHVAE_LAUNCH_CHECK("mu2_scatter_reduce");
    return 0;
}

extern "C" int fhvae_mu2_accumulate(const float* z2mu, int64_t ld_z, const int64_t* idx, float* zsum,
                                    float* cnt, int B, int Z, int64_t K, int32_t* err_flag, void* stream) {
    FHVAE_CHECK_ARG(z2mu && idx && zsum && cnt && B > 0 && Z > 0 && K > 0, "mu2_accumulate: bad argument");
    scatter_reduce_kernel<<<cdiv(B, 8), 256, 0, as_stream(stream)>>>(z2mu, ld_z, idx, zsum, cnt, nullptr,
                                                                     B, Z, K, err_flag);
    FHVAE_LAUNCH_CHECK("mu2_accumulate");
    return 0;
}

extern "C" int fhvae_mu2_estimate_finish(const float* zsum, const float* cnt, float* table, float r,
                                         int64_t K, int Z, void* stream) {
    FHVAE_CHECK_ARG(zsum && cnt && table && K > 0 && Z > 0, "mu2_estimate_finish: bad argument");
    mu2_estimate_finish_kernel<<<cdiv(K * Z, 256), 256, 0, as_stream(stream)>>>(zsum, cnt, table, r, K, Z);
    FHVAE_LAUNCH_CHECK("mu2_estimate_finish");
    return 0;
}

extern "C" int fhvae_rows_copy(const float* src, const int64_t* src_rows, float* dst,
                               const int64_t* dst_rows, int64_t n, int Z, void* stream) {
    FHVAE_CHECK_ARG(src && dst && n >= 0 && Z > 0, "rows_copy: bad argument");
    if (n == 0) return 0;
    rows_copy_kernel<<<cdiv(n * Z, 256), 256, 0, as_stream(stream)>>>(src, src_rows, dst, dst_rows, n, Z);
    FHVAE_LAUNCH_CHECK("rows_copy");
    return 0;
}

extern "C" int fhvae_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr,
                               float beta1, float beta2, float eps, float grad_scale, int32_t* step,
                               uint32_t* done_counter, void* stream) {
    FHVAE_CHECK_ARG(p && g && m && v && step && done_counter && n > 0, "adam_flat: bad argument");
    FHVAE_CHECK_ARG(((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)m % 16 == 0) &&
                        ((uintptr_t)v % 16 == 0), "adam_flat: buffers must be 16-byte aligned");
    int blocks = cdiv(n >> 2, 256);
    const int cap = kNumSM * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    launch_pdl(PDL_MISC, adam_flat_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), p, g, m, v, n, lr, beta1, beta2, eps,
               grad_scale, step, done_counter);
    FHVAE_LAUNCH_CHECK("adam_flat");
    return 0;
}

extern "C" int fhvae_gather_segments(const float* feats, const int64_t* start, const float* mean,
                                     const float* inv_std, float* out, int B, int T, int F, int64_t R,
                                     void* stream) {
    FHVAE_CHECK_ARG(feats && start && out && B > 0 && T > 0 && F > 0 && R > 0, "gather_segments: bad argument");
    FHVAE_CHECK_ARG((mean == nullptr) == (inv_std == nullptr), "gather_segments: mean and inv_std go together");
    gather_segments_kernel<<<cdiv((int64_t)B * T * F, 256), 256, 0, as_stream(stream)>>>(feats, start, mean,
                                                                                          inv_std, out, B, T, F, R);
    FHVAE_LAUNCH_CHECK("gather_segments");
    return 0;
}

extern "C" int fhvae_transpose_bt(const float* src, float* dst, int B, int T, int F, void* stream) {
    FHVAE_CHECK_ARG(src && dst && B > 0 && T > 0 && F > 0, "transpose_bt: bad argument");
    transpose_bt_kernel<<<cdiv((int64_t)B * T * F, 256), 256, 0, as_stream(stream)>>>(src, dst, B, T, F);
    FHVAE_LAUNCH_CHECK("transpose_bt");
    return 0;
}

extern "C" int fhvae_load_inputs(const float* x_src, float* x_dst, float* x_tm, int B, int T, int F, const int64_t* idx_src,
                                 int64_t* idx_dst, const int64_t* nsegs_src, int64_t* nsegs_dst, void* stream) {
    FHVAE_CHECK_ARG(x_src && x_dst && B > 0 && T > 0 && F > 0 && (idx_src == nullptr || idx_dst) &&
                    (nsegs_src == nullptr || nsegs_dst), "load_inputs: bad argument");
    FHVAE_CHECK_ARG(F % 4 == 0 && ((reinterpret_cast<uintptr_t>(x_src) | reinterpret_cast<uintptr_t>(x_dst) |
                                    reinterpret_cast<uintptr_t>(x_tm)) & 15) == 0,
                    "load_inputs: x must be 16-byte aligned with F %% 4 == 0");
    const int64_t n4 = (int64_t)B * T * F / 4;
    int grid = cdiv(n4, 256);
    if (grid < cdiv(B, 256)) grid = cdiv(B, 256);
    if (grid > 8 * kNumSM) grid = 8 * kNumSM;
    load_inputs_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(x_src),
                                                            reinterpret_cast<float4*>(x_dst), reinterpret_cast<float4*>(x_tm),
                                                            n4, B, T, F / 4, idx_src, idx_dst, nsegs_src, nsegs_dst);
    FHVAE_LAUNCH_CHECK("load_inputs");
    return 0;
}

extern "C" int fhvae_colsum_batch(const fhvae_colsum_problem* problems, int n_problems, void* stream) {
    FHVAE_CHECK_ARG(problems && n_problems > 0 && n_problems <= FHVAE_COLSUM_MAX_BATCH,
                    "colsum_batch: need 1..%d problems", FHVAE_COLSUM_MAX_BATCH);
    ColsumBatch cb;
    memset(&cb, 0, sizeof(cb));
    int total = 0;
    for (int i = 0; i < n_problems; ++i) {
        FHVAE_CHECK_ARG(problems[i].in && problems[i].out && problems[i].R > 0 && problems[i].C > 0,
                        "colsum_batch: problem %d invalid", i);
        cb.p[i] = problems[i];
        cb.start[i] = total;
        total += cdiv(problems[i].C, 32);
    }
    cb.start[n_problems] = total;
    cb.n = n_problems;
    colsum_kernel<<<total, 1024, 0, as_stream(stream)>>>(cb);
    FHVAE_LAUNCH_CHECK("colsum_batch");
    return 0;
}

extern "C" int fhvae_add2(float* out, const float* a, const float* b, int64_t n, void* stream) {
    FHVAE_CHECK_ARG(out && a && b && n > 0, "add2: bad argument");
    add2_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(out, a, b, n);
    FHVAE_LAUNCH_CHECK("add2");
    return 0;
}

extern "C" int fhvae_relu_bwd(float* dout, const float* out, int64_t n, void* stream) {
    FHVAE_CHECK_ARG(dout && out && n > 0, "relu_bwd: bad argument");
    relu_bwd_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(dout, out, n);
    FHVAE_LAUNCH_CHECK("relu_bwd");
    return 0;
}

extern "C" int fhvae_axpy(float* y, const float* x, float a, int64_t n, void* stream) {
    FHVAE_CHECK_ARG(y && x && n > 0, "axpy: bad argument");
    axpy_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(y, x, a, n);
    FHVAE_LAUNCH_CHECK("axpy");
    return 0;
}

extern "C" int fhvae_randn(float* out, int64_t n, uint64_t seed, unsigned long long* offset, uint32_t* done_counter,
                           void* stream) {
    FHVAE_CHECK_ARG(out && offset && done_counter && n > 0, "randn: bad argument");
    randn_kernel<<<cdiv((n + 3) / 4, 256), 256, 0, as_stream(stream)>>>(out, n, seed, offset, done_counter);
    FHVAE_LAUNCH_CHECK("randn");
    return 0;
}
