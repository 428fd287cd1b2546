// K7: discriminative log q(i | z2) over all rows of the mu2 table (simple_fhvae.py:119-122):
//   s_bn = -||z2mu_b - m_n||^2 / (2 * 0.25),  log q_b = s_{b,y_b} - logsumexp_n s_bn.
// The reference materialises (B,N,Z) and (B,N); here neither exists: table tiles stream through
// shared memory, the softmax is online (running max / running sum), the backward recomputes p_bn.
// The N range is split across CTAs (and, for a sharded table, across ranks): partial (max, sumexp)
// pairs are combined afterwards, so the same kernels serve the single- and multi-GPU paths.
// Distances are evaluated in the direct (z - m)^2 form in fp32 -- no ||z||^2+||m||^2-2zm cancellation.
#include <math.h>
#include "common.cuh"

namespace fhvae {

constexpr int D_THREADS = 256;
constexpr int D_ROWS = 128;   // table rows per shared-memory tile (4 per lane)

template <int Z>
__device__ __forceinline__ void load_table_tile(float (*ts)[D_ROWS + 1], const float* __restrict__ table,
                                                int64_t n0, int64_t n_end) {
    for (int e = threadIdx.x; e < D_ROWS * Z; e += D_THREADS) {
        const int r = e / Z, d = e % Z;
        ts[d][r] = (n0 + r < n_end) ? __ldg(table + (n0 + r) * Z + d) : 0.f;
    }
}

// ---- forward: partial (max, sumexp) over one N-split, 32 segments per CTA (4 per warp) ----------
template <int Z>
__global__ void __launch_bounds__(D_THREADS) disc_fwd_partial_kernel(
    const float* __restrict__ z2mu, int64_t ld_z, const float* __restrict__ table, int64_t N,
    float* __restrict__ part, int64_t rows_per_split, int B) {
    __shared__ float ts[Z][D_ROWS + 1];
    __shared__ float zs[32][Z];
    const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
    const int b0 = blockIdx.x * 32;
    const int64_t n_begin = (int64_t)blockIdx.y * rows_per_split;
    const int64_t n_end = min(N, n_begin + rows_per_split);
    for (int e = tid; e < 32 * Z; e += D_THREADS) {
        const int s = e / Z, d = e % Z;
        zs[s][d] = (b0 + s < B) ? z2mu[(int64_t)(b0 + s) * ld_z + d] : 0.f;
    }
    float mx[4], sm[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) { mx[s] = -INFINITY; sm[s] = 0.f; }
    for (int64_t n0 = n_begin; n0 < n_end; n0 += D_ROWS) {
        __syncthreads();
        load_table_tile<Z>(ts, table, n0, n_end);
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const int r = l + 32 * rr;
            float dist[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int d = 0; d < Z; ++d) {
                const float tv = ts[d][r];
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const float df = zs[w * 4 + s][d] - tv;
                    dist[s] = fmaf(df, df, dist[s]);
                }
            }
            if (n0 + r < n_end) {
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const float sc = -0.5f * kInvS2 * dist[s];
                    const float nm = fmaxf(mx[s], sc);
                    sm[s] = sm[s] * expf(mx[s] - nm) + expf(sc - nm);
                    mx[s] = nm;
                }
            }
        }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const float M = warp_max(mx[s]);
        const float v = (mx[s] == -INFINITY) ? 0.f : sm[s] * expf(mx[s] - M);
        const float S = warp_sum(v);
        const int b = b0 + w * 4 + s;
        if (l == 0 && b < B) {
            float* o = part + ((int64_t)blockIdx.y * B + b) * 2;
            o[0] = M;
            o[1] = S;
        }
    }
}

__global__ void disc_target_kernel(const float* __restrict__ z2mu, int64_t ld_z,
                                   const float* __restrict__ mu2, float* __restrict__ tgt, int B, int Z) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), l = threadIdx.x & 31;
    if (b >= B) return;
    float acc = 0.f;
    for (int d = l; d < Z; d += 32) {
        const float df = z2mu[(int64_t)b * ld_z + d] - mu2[(int64_t)b * Z + d];
        acc = fmaf(df, df, acc);
    }
    acc = warp_sum(acc);
    if (l == 0) tgt[b] = -0.5f * kInvS2 * acc;
}

__global__ void disc_combine_kernel(const float* __restrict__ part, int nparts,
                                    const float* __restrict__ tgt, float* __restrict__ log_qy,
                                    float* __restrict__ lse, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float M = -INFINITY;
    for (int i = 0; i < nparts; ++i) M = fmaxf(M, part[((int64_t)i * B + b) * 2]);
    float S = 0.f;
    for (int i = 0; i < nparts; ++i) {
        const float m = part[((int64_t)i * B + b) * 2], s = part[((int64_t)i * B + b) * 2 + 1];
        if (m != -INFINITY) S += s * expf(m - M);
    }
    const float L = M + logf(S);
    lse[b] = L;
    log_qy[b] = tgt[b] - L;
}

// ---- backward, dense table rows: dtable[n] = -sum_b g_b p_bn (z_b - m_n) / s2 ---------------------
// lane <-> table row (32 rows per CTA), warp <-> segment subset, fixed-order cross-warp reduction.
template <int Z>
__global__ void __launch_bounds__(D_THREADS) disc_bwd_rows_kernel(
    const float* __restrict__ z2mu, int64_t ld_z, const float* __restrict__ table, int64_t N,
    const float* __restrict__ lse, const float* __restrict__ g, float* __restrict__ dtable, int B) {
    __shared__ float zs[64][Z];
    __shared__ float ls[64], gs[64];
    __shared__ float red[8][32][Z + 1];
    const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
    const int64_t n0 = (int64_t)blockIdx.x * 32;
    // stage this CTA's 32 rows through red[0] (coalesced), then into registers
    for (int e = tid; e < 32 * Z; e += D_THREADS) {
        const int r = e / Z, d = e % Z;
        red[0][r][d] = (n0 + r < N) ? __ldg(table + (n0 + r) * Z + d) : 0.f;
    }
    __syncthreads();
    float mrow[Z], acc[Z];
#pragma unroll
    for (int d = 0; d < Z; ++d) { mrow[d] = red[0][l][d]; acc[d] = 0.f; }
    for (int bc = 0; bc < B; bc += 64) {
        __syncthreads();
        for (int e = tid; e < 64 * Z; e += D_THREADS) {
            const int s = e / Z, d = e % Z;
            zs[s][d] = (bc + s < B) ? z2mu[(int64_t)(bc + s) * ld_z + d] : 0.f;
        }
        if (tid < 64) {
            ls[tid] = (bc + tid < B) ? lse[bc + tid] : 0.f;
            gs[tid] = (bc + tid < B) ? g[bc + tid] : 0.f;
        }
        __syncthreads();
        for (int j = w; j < 64 && bc + j < B; j += 8) {
            float dist = 0.f;
#pragma unroll
            for (int d = 0; d < Z; ++d) {
                const float df = zs[j][d] - mrow[d];
                dist = fmaf(df, df, dist);
            }
            const float cf = -gs[j] * kInvS2 * expf(-0.5f * kInvS2 * dist - ls[j]);
#pragma unroll
            for (int d = 0; d < Z; ++d) acc[d] = fmaf(cf, zs[j][d] - mrow[d], acc[d]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int d = 0; d < Z; ++d) red[w][l][d] = acc[d];
    __syncthreads();
    for (int e = tid; e < 32 * Z; e += D_THREADS) {
        const int r = e / Z, d = e % Z;
        if (n0 + r < N) {
            float s = 0.f;
#pragma unroll
            for (int ww = 0; ww < 8; ++ww) s += red[ww][r][d];
            dtable[(n0 + r) * Z + d] = s;
        }
    }
}

// ---- backward, per segment: sumpm_part[split][b] = sum_{n in split} p_bn m_n ------------------------
// 16 segments per CTA (2 per warp), lanes <-> rows, butterfly reduction over lanes.
template <int Z>
__global__ void __launch_bounds__(D_THREADS) disc_bwd_segs_kernel(
    const float* __restrict__ z2mu, int64_t ld_z, const float* __restrict__ table, int64_t N,
    const float* __restrict__ lse, float* __restrict__ sumpm_part, int64_t rows_per_split, int B) {
    __shared__ float ts[Z][D_ROWS + 1];
    __shared__ float zs[16][Z];
    __shared__ float ls[16];
    const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
    const int b0 = blockIdx.x * 16;
    const int64_t n_begin = (int64_t)blockIdx.y * rows_per_split;
    const int64_t n_end = min(N, n_begin + rows_per_split);
    for (int e = tid; e < 16 * Z; e += D_THREADS) {
        const int s = e / Z, d = e % Z;
        zs[s][d] = (b0 + s < B) ? z2mu[(int64_t)(b0 + s) * ld_z + d] : 0.f;
    }
    if (tid < 16) ls[tid] = (b0 + tid < B) ? lse[b0 + tid] : 0.f;
    float acc[2][Z];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int d = 0; d < Z; ++d) acc[s][d] = 0.f;
    for (int64_t n0 = n_begin; n0 < n_end; n0 += D_ROWS) {
        __syncthreads();
        load_table_tile<Z>(ts, table, n0, n_end);
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const int r = l + 32 * rr;
            float dist[2] = {0.f, 0.f};
#pragma unroll
            for (int d = 0; d < Z; ++d) {
                const float tv = ts[d][r];
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const float df = zs[w * 2 + s][d] - tv;
                    dist[s] = fmaf(df, df, dist[s]);
                }
            }
            const bool ok = (n0 + r < n_end);
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const float p = ok ? expf(-0.5f * kInvS2 * dist[s] - ls[w * 2 + s]) : 0.f;
#pragma unroll
                for (int d = 0; d < Z; ++d) acc[s][d] = fmaf(p, ts[d][r], acc[s][d]);
            }
        }
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int b = b0 + w * 2 + s;
#pragma unroll
        for (int d = 0; d < Z; ++d) {
            const float v = warp_sum(acc[s][d]);
            if (l == (d & 31) && b < B) sumpm_part[((int64_t)blockIdx.y * B + b) * Z + d] = v;
        }
    }
}

__global__ void disc_bwd_finish_kernel(const float* __restrict__ z2mu, int64_t ld_z,
                                       const float* __restrict__ mu2,
                                       const float* __restrict__ sumpm_part, int nparts,
                                       const float* __restrict__ g, float* __restrict__ dz2mu,
                                       int64_t ld_dz, float* __restrict__ dmu2, int B, int Z) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * Z) return;
    const int b = i / Z, d = i % Z;
    float sp = 0.f;
    for (int k = 0; k < nparts; ++k) sp += sumpm_part[((int64_t)k * B + b) * Z + d];
    const float my = mu2[i], z = z2mu[(int64_t)b * ld_z + d], gb = g[b] * kInvS2;
    dz2mu[(int64_t)b * ld_dz + d] += gb * (my - sp);
    dmu2[i] += gb * (z - my);
}

// ---- sharded table (row u lives on rank u mod W at local row u / W): per-step exchange helpers -----------
// One packet row per segment: [z2_mu (Z floats) | mu_idx (int64, 2 floats) | g = dL/dlog_qy | pad] -- ONE all-gather
// per step carries everything the other ranks need to score this rank's segments against their rows.
__global__ void shard_pack_kernel(const float* __restrict__ z2mu, int64_t ld_z, const int64_t* __restrict__ idx,
                                  const float* __restrict__ g, float* __restrict__ packet, int B, int Z) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int ld = Z + 4;
    if (i >= B * ld) return;
    const int b = i / ld, d = i % ld;
    float v = 0.f;
    if (d < Z) v = z2mu[(int64_t)b * ld_z + d];
    else if (d == Z) v = __int_as_float((int)(uint32_t)(idx[b] & 0xffffffffll));
    else if (d == Z + 1) v = __int_as_float((int)(uint32_t)((uint64_t)idx[b] >> 32));
    else if (d == Z + 2) v = g[b];
    packet[i] = v;
}
// gathered packets (Bg rows) -> local row of every global segment's utterance on THIS rank (-1: owned elsewhere),
// the raw ids and the contiguous g vector.  Exact int64 arithmetic.
__global__ void shard_unpack_kernel(const float* __restrict__ packet, int Bg, int Z, int world, int rank, int64_t N_local,
                                    int64_t* __restrict__ idx_g, int64_t* __restrict__ lidx_g, float* __restrict__ g_g,
                                    int32_t* __restrict__ err_flag) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= Bg) return;
    const float* row = packet + (int64_t)b * (Z + 4);
    const uint64_t lo = (uint32_t)__float_as_int(row[Z]), hi = (uint32_t)__float_as_int(row[Z + 1]);
    const int64_t u = (int64_t)(lo | (hi << 32));
    int64_t l = -1;
    if (u >= 0 && (u % world) == rank) {
        l = u / world;
        if (l >= N_local) {                       // beyond the table: same treatment as fhvae_mu2_gather
            l = -1;
            if (err_flag) atomicOr(err_flag, FHVAE_FLAG_BAD_INDEX);
        }
    } else if (u < 0 && err_flag) {
        atomicOr(err_flag, FHVAE_FLAG_BAD_INDEX);
    }
    if (idx_g) idx_g[b] = u;
    lidx_g[b] = l;
    g_g[b] = row[Z + 2];
}
// (max, sumexp) partials of every rank and N-split, combined in FIXED (rank-major) order for ALL global segments --
// every rank computes bit-identical lse_g -- plus log q of this rank's own segments [b_off, b_off + B_local).
__global__ void disc_combine_sharded_kernel(const float* __restrict__ part, int nparts, int Bg,
                                            const float* __restrict__ tgt, int b_off, int B_local,
                                            float* __restrict__ log_qy, float* __restrict__ lse_g) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= Bg) return;
    float M = -INFINITY;
    for (int i = 0; i < nparts; ++i) M = fmaxf(M, part[((int64_t)i * Bg + b) * 2]);
    float S = 0.f;
    for (int i = 0; i < nparts; ++i) {
        const float m = part[((int64_t)i * Bg + b) * 2], s = part[((int64_t)i * Bg + b) * 2 + 1];
        if (m != -INFINITY) S += s * expf(m - M);
    }
    const float L = M + logf(S);
    lse_g[b] = L;
    const int bl = b - b_off;
    if (bl >= 0 && bl < B_local) log_qy[bl] = tgt[bl] - L;
}

static int64_t rows_per_split(int64_t N, int nsplit) {
    int64_t r = (N + nsplit - 1) / nsplit;
    return (r + D_ROWS - 1) / D_ROWS * D_ROWS;
}

}  // namespace fhvae

using namespace fhvae;

#define DISPATCH_Z(Z, ...)                                              \
    switch (Z) {                                                        \
        case 8:  { constexpr int ZC = 8;  __VA_ARGS__; } break;         \
        case 16: { constexpr int ZC = 16; __VA_ARGS__; } break;         \
        case 32: { constexpr int ZC = 32; __VA_ARGS__; } break;         \
        default: set_error("disc: z2_dim %d not in {8,16,32}", Z); return FHVAE_ENOSUP; \
    }

extern "C" int fhvae_disc_nsplit(int B, int64_t N) {
    if (B <= 0 || N <= 0) return 1;
    const int seg_ctas = cdiv(B, 32);
    int want = cdiv(2 * kNumSM, seg_ctas);
    const int maxs = cdiv(N, D_ROWS);
    if (want > maxs) want = maxs;
    if (want < 1) want = 1;
    // make sure every split owns at least one row
    while (want > 1 && rows_per_split(N, want) * (want - 1) >= N) --want;
    return want;
}

extern "C" int fhvae_disc_fwd_partial(const float* z2mu, int64_t ld_z, const float* table, int64_t N,
                                      int Z, float* part, int nsplit, int B, void* stream) {
    FHVAE_CHECK_ARG(z2mu && table && part && N > 0 && B > 0 && nsplit > 0, "disc_fwd_partial: bad argument");
    dim3 grid(cdiv(B, 32), nsplit);
    const int64_t rps = rows_per_split(N, nsplit);
    DISPATCH_Z(Z, (disc_fwd_partial_kernel<ZC><<<grid, D_THREADS, 0, as_stream(stream)>>>(
                       z2mu, ld_z, table, N, part, rps, B)));
    FHVAE_LAUNCH_CHECK("disc_fwd_partial");
    return 0;
}

extern "C" int fhvae_disc_target(const float* z2mu, int64_t ld_z, const float* mu2, float* tgt, int B,
                                 int Z, void* stream) {
    FHVAE_CHECK_ARG(z2mu && mu2 && tgt && B > 0 && Z > 0, "disc_target: bad argument");
    disc_target_kernel<<<cdiv(B, 8), 256, 0, as_stream(stream)>>>(z2mu, ld_z, mu2, tgt, B, Z);
    FHVAE_LAUNCH_CHECK("disc_target");
    return 0;
}

extern "C" int fhvae_disc_combine(const float* part, int nparts, const float* tgt, float* log_qy,
                                  float* lse, int B, void* stream) {
    FHVAE_CHECK_ARG(part && tgt && log_qy && lse && nparts > 0 && B > 0, "disc_combine: bad argument");
    disc_combine_kernel<<<cdiv(B, 128), 128, 0, as_stream(stream)>>>(part, nparts, tgt, log_qy, lse, B);
    FHVAE_LAUNCH_CHECK("disc_combine");
    return 0;
}

extern "C" int fhvae_disc_bwd_rows(const float* z2mu, int64_t ld_z, const float* table, int64_t N,
                                   int Z, const float* lse, const float* g, float* dtable, int B,
                                   void* stream) {
    FHVAE_CHECK_ARG(z2mu && table && lse && g && dtable && N > 0 && B > 0, "disc_bwd_rows: bad argument");
    DISPATCH_Z(Z, (disc_bwd_rows_kernel<ZC><<<cdiv(N, 32), D_THREADS, 0, as_stream(stream)>>>(
                       z2mu, ld_z, table, N, lse, g, dtable, B)));
    FHVAE_LAUNCH_CHECK("disc_bwd_rows");
    return 0;
}

extern "C" int fhvae_disc_bwd_segs(const float* z2mu, int64_t ld_z, const float* table, int64_t N,
                                   int Z, const float* lse, float* sumpm_part, int nsplit, int B,
                                   void* stream) {
    FHVAE_CHECK_ARG(z2mu && table && lse && sumpm_part && N > 0 && B > 0 && nsplit > 0,
                    "disc_bwd_segs: bad argument");
    dim3 grid(cdiv(B, 16), nsplit);
    const int64_t rps = rows_per_split(N, nsplit);
    DISPATCH_Z(Z, (disc_bwd_segs_kernel<ZC><<<grid, D_THREADS, 0, as_stream(stream)>>>(
                       z2mu, ld_z, table, N, lse, sumpm_part, rps, B)));
    FHVAE_LAUNCH_CHECK("disc_bwd_segs");
    return 0;
}

extern "C" int fhvae_disc_bwd_finish(const float* z2mu, int64_t ld_z, const float* mu2,
                                     const float* sumpm_part, int nparts, const float* g,
                                     float* dz2mu, int64_t ld_dz, float* dmu2, int B, int Z,
                                     void* stream) {
    FHVAE_CHECK_ARG(z2mu && mu2 && sumpm_part && g && dz2mu && dmu2 && nparts > 0 && B > 0 && Z > 0,
                    "disc_bwd_finish: bad argument");
    disc_bwd_finish_kernel<<<cdiv((int64_t)B * Z, 256), 256, 0, as_stream(stream)>>>(
        z2mu, ld_z, mu2, sumpm_part, nparts, g, dz2mu, ld_dz, dmu2, B, Z);
    FHVAE_LAUNCH_CHECK("disc_bwd_finish");
    return 0;
}


extern "C" int fhvae_shard_pack(const float* z2mu, int64_t ld_z, const int64_t* idx, const float* g, float* packet,
                                int B, int Z, void* stream) {
    FHVAE_CHECK_ARG(z2mu && idx && g && packet && B > 0 && Z > 0, "shard_pack: bad argument");
    shard_pack_kernel<<<cdiv((int64_t)B * (Z + 4), 256), 256, 0, as_stream(stream)>>>(z2mu, ld_z, idx, g, packet, B, Z);
    FHVAE_LAUNCH_CHECK("shard_pack");
    return 0;
}

extern "C" int fhvae_shard_unpack(const float* packet, int Bg, int Z, int world, int rank, int64_t N_local,
                                  int64_t* idx_g, int64_t* lidx_g, float* g_g, int32_t* err_flag, void* stream) {
    FHVAE_CHECK_ARG(packet && lidx_g && g_g && Bg > 0 && Z > 0 && world > 0 && rank >= 0 && rank < world && N_local >= 0,
                    "shard_unpack: bad argument");
    shard_unpack_kernel<<<cdiv(Bg, 256), 256, 0, as_stream(stream)>>>(packet, Bg, Z, world, rank, N_local, idx_g, lidx_g,
                                                                      g_g, err_flag);
    FHVAE_LAUNCH_CHECK("shard_unpack");
    return 0;
}

extern "C" int fhvae_disc_combine_sharded(const float* part, int nparts, int Bg, const float* tgt, int b_off,
                                          int B_local, float* log_qy, float* lse_g, void* stream) {
    FHVAE_CHECK_ARG(part && tgt && log_qy && lse_g && nparts > 0 && Bg > 0 && b_off >= 0 && B_local > 0 &&
                        b_off + B_local <= Bg, "disc_combine_sharded: bad argument");
    disc_combine_sharded_kernel<<<cdiv(Bg, 128), 128, 0, as_stream(stream)>>>(part, nparts, Bg, tgt, b_off, B_local,
                                                                              log_qy, lse_g);
    FHVAE_LAUNCH_CHECK("disc_combine_sharded");
    return 0;
}
