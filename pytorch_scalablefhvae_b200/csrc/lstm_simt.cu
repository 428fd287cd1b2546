// LSTM recurrence, FHVAE_MODE_F32_SIMT: one launch per time step, fp32 FFMA.  This is the exact
// reference the persistent tensor-core recurrence (lstm_cluster.cu) is checked against on the GPU.
// Cell maths: PyTorch nn.LSTM (i,f,g,o), SURVEY.md Appendix C.
#include "common.cuh"

namespace fhvae {

constexpr int LB = 16;   // batch rows per CTA
constexpr int LU = 16;   // hidden units per CTA
constexpr int LK = 32;   // k chunk

// gates_t = P[t] + Q + h_{t-1} W_hh^T ; pointwise ; write h_t, c_t, acts_t
__global__ void __launch_bounds__(256) lstm_fwd_step_kernel(
    const float* __restrict__ Pt, const float* __restrict__ Q, const float* __restrict__ W_hh,
    const float* __restrict__ h_prev, const float* __restrict__ c_prev,
    float* __restrict__ h_out, float* __restrict__ c_out, float* __restrict__ acts_out,
    int B, int H) {
    __shared__ float hs[LB][LK + 1];
    __shared__ float ws[4][LU][LK + 1];
    const int tid = threadIdx.x, tb = tid >> 4, tu = tid & 15;
    const int b0 = blockIdx.y * LB, u0 = blockIdx.x * LU;
    const int b = b0 + tb, u = u0 + tu;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (h_prev != nullptr) {
        for (int k0 = 0; k0 < H; k0 += LK) {
            for (int e = tid; e < LB * LK; e += 256) {
                const int r = e / LK, k = e % LK;
                hs[r][k] = (b0 + r < B && k0 + k < H) ? h_prev[(int64_t)(b0 + r) * H + k0 + k] : 0.f;
            }
            for (int e = tid; e < 4 * LU * LK; e += 256) {
                const int k = e % LK, ru = (e / LK) % LU, g = e / (LK * LU);
                ws[g][ru][k] = (u0 + ru < H && k0 + k < H)
                                   ? __ldg(W_hh + (int64_t)(g * H + u0 + ru) * H + k0 + k) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < LK; ++k) {
                const float hv = hs[tb][k];
#pragma unroll
                for (int g = 0; g < 4; ++g) acc[g] = fmaf(hv, ws[g][tu][k], acc[g]);
            }
            __syncthreads();
        }
    }
    if (b >= B || u >= H) return;
    const int64_t row4 = (int64_t)b * 4 * H;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        if (Pt) acc[g] += Pt[row4 + g * H + u];
        if (Q) acc[g] += Q[row4 + g * H + u];
    }
    const float ig = sigmoidf_acc(acc[0]);
    const float fg = sigmoidf_acc(acc[1]);
    const float gg = tanhf(acc[2]);
    const float og = sigmoidf_acc(acc[3]);
    const float cp = c_prev ? c_prev[(int64_t)b * H + u] : 0.f;
    const float c = fmaf(fg, cp, ig * gg);
    const float h = og * tanhf(c);
    h_out[(int64_t)b * H + u] = h;
    c_out[(int64_t)b * H + u] = c;
    acts_out[row4 + 0 * H + u] = ig;
    acts_out[row4 + 1 * H + u] = fg;
    acts_out[row4 + 2 * H + u] = gg;
    acts_out[row4 + 3 * H + u] = og;
}

// dh = dh_all[t] + dh_last(t==T-1) + dgates[t+1] W_hh ; pointwise BPTT ; write dgates[t], dc, dgsum
__global__ void __launch_bounds__(256) lstm_bwd_step_kernel(
    const float* __restrict__ dh_t, const float* __restrict__ dh_last,
    const float* __restrict__ W_hh, const float* __restrict__ dg_next,
    const float* __restrict__ c_t, const float* __restrict__ c_prev, const float* __restrict__ acts_t,
    float* __restrict__ dg_t, float* __restrict__ dgsum, float* __restrict__ dc_state,
    int first, int B, int H) {
    __shared__ float ds[LB][LK + 1];
    __shared__ float ws[LK][LU + 1];
    const int tid = threadIdx.x, tb = tid >> 4, tu = tid & 15;
    const int b0 = blockIdx.y * LB, u0 = blockIdx.x * LU;
    const int b = b0 + tb, u = u0 + tu;
    const int H4 = 4 * H;
    float acc = 0.f;
    if (dg_next != nullptr) {
        for (int j0 = 0; j0 < H4; j0 += LK) {
            for (int e = tid; e < LB * LK; e += 256) {
                const int r = e / LK, j = e % LK;
                ds[r][j] = (b0 + r < B) ? dg_next[(int64_t)(b0 + r) * H4 + j0 + j] : 0.f;
            }
            for (int e = tid; e < LK * LU; e += 256) {
                const int ru = e % LU, j = e / LU;
                ws[j][ru] = (u0 + ru < H) ? __ldg(W_hh + (int64_t)(j0 + j) * H + u0 + ru) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < LK; ++j) acc = fmaf(ds[tb][j], ws[j][tu], acc);
            __syncthreads();
        }
    }
    if (b >= B || u >= H) return;
    const int64_t bu = (int64_t)b * H + u, row4 = (int64_t)b * H4;
    float dh = acc;
    if (dh_t) dh += dh_t[bu];
    if (dh_last) dh += dh_last[bu];
    const float ig = acts_t[row4 + u], fg = acts_t[row4 + H + u];
    const float gg = acts_t[row4 + 2 * H + u], og = acts_t[row4 + 3 * H + u];
    const float tc = tanhf(c_t[bu]);
    const float cp = c_prev ? c_prev[bu] : 0.f;
    float dc = (first ? 0.f : dc_state[bu]) + dh * og * (1.f - tc * tc);
    const float d_o = dh * tc;
    const float d_i = dc * gg, d_g = dc * ig, d_f = dc * cp;
    dc_state[bu] = dc * fg;
    const float gi = d_i * ig * (1.f - ig);
    const float gf = d_f * fg * (1.f - fg);
    const float g_g = d_g * (1.f - gg * gg);
    const float go = d_o * og * (1.f - og);
    dg_t[row4 + u] = gi;
    dg_t[row4 + H + u] = gf;
    dg_t[row4 + 2 * H + u] = g_g;
    dg_t[row4 + 3 * H + u] = go;
    if (dgsum) {
        if (first) {
            dgsum[row4 + u] = gi; dgsum[row4 + H + u] = gf;
            dgsum[row4 + 2 * H + u] = g_g; dgsum[row4 + 3 * H + u] = go;
        } else {
            dgsum[row4 + u] += gi; dgsum[row4 + H + u] += gf;
            dgsum[row4 + 2 * H + u] += g_g; dgsum[row4 + 3 * H + u] += go;
        }
    }
}

int lstm_fwd_simt(const float* P, const float* Q, const float* W_hh, float* h_all, float* c_all,
                  float* acts, int T, int B, int H, cudaStream_t st) {
    dim3 grid(cdiv(H, LU), cdiv(B, LB));
    const int64_t bh = (int64_t)B * H;
    for (int t = 0; t < T; ++t) {
        lstm_fwd_step_kernel<<<grid, 256, 0, st>>>(
            P ? P + t * 4 * bh : nullptr, Q, W_hh,
            t ? h_all + (t - 1) * bh : nullptr, t ? c_all + (t - 1) * bh : nullptr,
            h_all + t * bh, c_all + t * bh, acts + t * 4 * bh, B, H);
    }
    count_launches(T - 1);
    FHVAE_LAUNCH_CHECK("lstm_fwd_simt");
    return 0;
}

int lstm_bwd_simt(const float* dh_all, const float* dh_last, const float* W_hh, const float* c_all,
                  const float* acts, float* dgates, float* dgsum, float* dc, int T, int B, int H,
                  cudaStream_t st) {
    dim3 grid(cdiv(H, LU), cdiv(B, LB));
    const int64_t bh = (int64_t)B * H;
    for (int t = T - 1; t >= 0; --t) {
        lstm_bwd_step_kernel<<<grid, 256, 0, st>>>(
            dh_all ? dh_all + t * bh : nullptr, (t == T - 1) ? dh_last : nullptr, W_hh,
            (t == T - 1) ? nullptr : dgates + (t + 1) * 4 * bh,
            c_all + t * bh, t ? c_all + (t - 1) * bh : nullptr, acts + t * 4 * bh,
            dgates + t * 4 * bh, dgsum, dc, t == T - 1, B, H);
    }
    count_launches(T - 1);
    FHVAE_LAUNCH_CHECK("lstm_bwd_simt");
    return 0;
}

}  // namespace fhvae
