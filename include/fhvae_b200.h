/* fhvae_b200.h -- C ABI of libfhvae_b200.so: the sm_100a kernels behind the ScalableFHVAE
 * train / inference step.
 *
 * The reference (BurnhamG/PyTorch-ScalableFHVAE) is pure Python and has NO operator registry, FFI
 * or plugin API: its drop-in boundary is the nn.Module surface
 * (simple_fhvae.py:8-124, fhvae.py:4-14).  This header is therefore the layer *below* that
 * surface: each entry point names the reference ATen op group it replaces (file:line into the
 * reference).  The Python mirror (pytorch_scalablefhvae_b200/model.py) binds these with ctypes --
 * see INTEGRATION.md for the stub a maintainer would add to the reference.
 *
 * Conventions
 *  - plain pointers + sizes, no torch types; every pointer is DEVICE memory owned by the caller
 *    (PyTorch); the library never allocates, frees or retains a pointer past the call;
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises,
 *    nothing reads device memory on the host => every call is CUDA-graph capturable;
 *  - return 0 on success; <0 = argument error (FHVAE_E*); >0 = cudaError_t of the launch.
 *    fhvae_last_error_string() describes the last failure of the calling thread;
 *  - fp32 everywhere at the boundary; int64 indices (torch.long), exact;
 *  - LSTM-internal tensors are TIME-MAJOR: (T, B, *).
 */
#ifndef FHVAE_B200_H
#define FHVAE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FHVAE_EINVAL (-1)   /* bad size / null pointer */
#define FHVAE_ENOSUP (-2)   /* shape not supported by this kernel */

/* GEMM compute modes (fhvae_gemm_batch `mode`, fhvae_lstm_* `mode`) */
#define FHVAE_MODE_F32_SIMT 0   /* fp32 FFMA, exact-order reference kernels                  */
#define FHVAE_MODE_BF16X3   1   /* tcgen05 kind::f16, hi/lo bf16 split x3 (fp32-parity mode)   */
#define FHVAE_MODE_BF16     2   /* tcgen05 kind::f16, single bf16 pass ("bf16 input-GEMM mode") */

/* bits of the per-plan device status word (the `nan_flag` argument of fhvae_elbo_fwd, the `err_flag` of the
 * table kernels): the host reads it when it wants to (train_model.py:464-466 is a forced sync upstream) */
#define FHVAE_FLAG_NAN        1  /* a NaN lower bound was produced (train_model.py:464)            */
#define FHVAE_FLAG_BAD_INDEX  2  /* mu_idx / label outside [0, N): torch.gather would have raised  */

const char* fhvae_last_error_string(void);
int fhvae_version(void);
/* compile-time facts the host may assert on */
int fhvae_built_for_sm(void);              /* 100 */
/* kernels launched by this process through this library so far (bench.py `gpu_launches`) */
unsigned long long fhvae_launch_count(void);
/* Process-wide switch (returns the previous value).  On: every split-K launch of fhvae_gemm_batch /
 * fhvae_wgrad_planes_batch uses at most TWO partials per output tile, added into a pre-zeroed C by
 * red.global.add -- (0 + a) + b == (0 + b) + a exactly, so results no longer depend on arrival order and two runs
 * of a step are bit-identical (everything else on the path already sums in a fixed order).  Off (default): K is
 * split to fill the SMs (faster; weight gradients differ in the last bits from run to run).  Launch geometries are
 * baked into captured CUDA graphs: set it before the first step. */
int fhvae_set_deterministic(int on);
int fhvae_get_deterministic(void);

/* ---------------------------------------------------------------------------------------------
 * K1/K3/K5/K8 dense contractions.  Replaces nn.Linear addmm (simple_fhvae.py:130-134,208-212),
 * the nn.LSTM input projections of the FHVAE restatement, and their autograd dgrad/wgrad mm's.
 *   C[m,n] = sum_k A(m,k) * B(k,n) + bias[n] + beta * C[m,n];   optional ReLU.
 *   A(m,k) = A[m*sa_m + k*sa_k],  B(k,n) = B[k*sb_k + n*sb_n]  (element strides; one of each pair is 1)
 * Up to FHVAE_GEMM_MAX_BATCH independent problems per launch (grouped GEMM: one grid).
 * ------------------------------------------------------------------------------------------- */
#define FHVAE_GEMM_MAX_BATCH 24
typedef struct fhvae_gemm_problem {
    const float* A;
    const float* B;
    float*       C;
    const float* bias;          /* (N,) or NULL */
    int32_t M, N, K;
    int32_t relu;               /* 1: C = max(C, 0) */
    int64_t sa_m, sa_k;
    int64_t sb_k, sb_n;
    int64_t ldc;
    float   beta;               /* 0 or 1 (any value accepted) */
    int32_t reserved;
} fhvae_gemm_problem;
int fhvae_gemm_batch(const fhvae_gemm_problem* problems, int n_problems, int mode, void* stream);

/* ---------------------------------------------------------------------------------------------
 * LSTM recurrence (PyTorch nn.LSTM cell, gate order i,f,g,o; zero initial state).
 *   gates_t = P[t] (T,B,4H, may be NULL) + Q (B,4H, time-invariant, may be NULL) + h_{t-1} W_hh^T
 *   (biases are folded into P or Q by the projection GEMM).
 * Saves for BPTT: h_all (T,B,H), c_all (T,B,H), acts (T,B,4H) = post-activation i,f,g,o.
 * xchg: caller-allocated scratch of 16*B*H floats (contents undefined), the L2-resident exchange buffer of
 * the cluster kernels (may be NULL in FHVAE_MODE_F32_SIMT).
 * ------------------------------------------------------------------------------------------- */
int fhvae_lstm_fwd(const float* P, const float* Q, const float* W_hh,
                   float* h_all, float* c_all, float* acts, float* xchg,
                   int T, int B, int H, int mode, void* stream);
/* BPTT.  dh_all (T,B,H) = dL/dh_t from the consumer of all outputs (may be NULL);
 * dh_last (B,H) = extra dL/dh_{T-1} from the final-state consumer (may be NULL).
 * Outputs: dgates (T,B,4H) pre-activation gate gradients; dgsum (B,4H) = sum_t dgates (may be NULL).
 * Scratch: dh_rec (16,B,H) (exchange buffer of the cluster kernel) and dc (B,H), caller-allocated,
 * contents undefined on entry. */
int fhvae_lstm_bwd(const float* dh_all, const float* dh_last, const float* W_hh,
                   const float* c_all, const float* acts,
                   float* dgates, float* dgsum, float* dh_rec, float* dc,
                   int T, int B, int H, int mode, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Layer-wavefront recurrence over a stack of 1 or 2 LSTM layers in ONE launch (H in {128, 256}: groups of H/32 CTAs; B % 32 == 0,
 * T <= 63, tensor-core modes only).  Layer 1 runs one step behind layer 0 and computes its own input
 * projection  h0_t W_ih1^T + bias1  from the words layer 0 publishes, so the (T,B,4H) projection buffer and
 * the projection GEMM of layer 1 do not exist.  Replaces the nn.LSTM(num_layers=2) forward of the
 * restatement (reference fhvae.py:4-14 is a stub).
 *   layer 0: P0 (T,B,4H) and/or Q0 (B,4H) as in fhvae_lstm_fwd;  layer 1: W_ih1 (4H,H), bias1 (4H) = b_ih+b_hh.
 *   xchg: caller-allocated, ZERO-INITIALISED ONCE, fhvae_lstm_wave_xchg_bytes(T,B,H,nlayers) bytes, dedicated
 *   to one (T,B,nlayers) geometry; it holds the exchange words and per-CTA launch counters and must not be
 *   written by the caller afterwards.  fhvae_lstm_wave_supported returns 1 if the shape/mode is served.
 * ------------------------------------------------------------------------------------------- */
int fhvae_lstm_wave_supported(int T, int B, int H, int nlayers, int mode);
/* batch rows ONE launch serves on this device (groups of 32 rows x CTAs per group x layers <= SM count): 288 for two
 * layers of H = 256 on 148 SMs.  A larger batch runs as consecutive launches, so forward-only drivers (posterior
 * extraction, eval_model.py:41-59) size their batches as multiples of it.  0: shape not served. */
int fhvae_lstm_wave_rows_per_launch(int H, int nlayers);
long long fhvae_lstm_wave_xchg_bytes(int T, int B, int H, int nlayers);
int fhvae_lstm_wave_fwd(const float* P0, const float* Q0, const float* W_hh0, float* h0, float* c0, float* acts0,
                        const float* W_ih1, const float* bias1, const float* W_hh1, float* h1, float* c1,
                        float* acts1, void* xchg, int T, int B, int H, int nlayers, int mode, void* stream);

/* BPTT of the same stack in ONE launch, top layer first; the bottom layer receives
 *   dh0_t = dgates0_{t+1} W_hh0 + dgates1_t W_ih1 (+ dh_last_bot at t = T-1)
 * inside the kernel (cross-layer product accumulated into the recurrent split-K accumulator), so the layer-1 dgrad
 * GEMM and its (T,B,H) buffer do not exist.  nlayers == 1: only the *_top arguments are used.
 * Outputs as fhvae_lstm_bwd: dgates_* (T,B,4H), dgsum_* (B,4H, may be NULL).  xchg: ZERO-INITIALISED ONCE,
 * fhvae_lstm_wave_bwd_xchg_bytes bytes, dedicated to one (T,B,nlayers) geometry and to this entry point. */
long long fhvae_lstm_wave_bwd_xchg_bytes(int T, int B, int H, int nlayers);
int fhvae_lstm_wave_bwd(const float* dh_all_top, const float* dh_last_top, const float* dh_last_bot,
                        const float* W_hh_top, const float* c_top, const float* acts_top, float* dgates_top,
                        float* dgsum_top, const float* W_ih_top, const float* W_hh_bot, const float* c_bot,
                        const float* acts_bot, float* dgates_bot, float* dgsum_bot, void* xchg, int T, int B, int H,
                        int nlayers, int mode, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2/K4/K6: reparameterisation + ELBO terms (simple_fhvae.py:56-69, :106-116, :213-216).
 * ------------------------------------------------------------------------------------------- */
/* head (B, 2Z) = [mu | logvar] rows of leading dim ld_head;  sample[b*ld_s + d] = mu + eps*exp(.5 logvar) */
int fhvae_reparam_fwd(const float* head, int64_t ld_head, const float* eps,
                      float* sample, int64_t ld_s, int B, int Z, void* stream);
/* dhead[b, 0:Z] (+)= dsample ; dhead[b, Z:2Z] (+)= dsample * 0.5*eps*exp(.5 logvar).
 * accumulate=0 overwrites dhead, 1 adds to it. */
int fhvae_reparam_bwd(const float* head, int64_t ld_head, const float* eps,
                      const float* dsample, int64_t ld_ds,
                      float* dhead, int64_t ld_dh, int accumulate, int B, int Z, void* stream);

/* One CTA per segment.  x (B,T,F) contiguous.  Decoder head element (b,t,f):
 *   mu = xhead[b*xs_b + t*xs_t + f],  logvar = xhead[b*xs_b + t*xs_t + lv_off + f].
 * z1head / z2head: (B, 2Z) rows [mu | logvar].  mu2 (B,Z2) = gathered table rows (fhvae_mu2_gather),
 * nsegs (B,) int64.
 * out5 (5,B): lower_bound, log_px_z, neg_kld_z1, neg_kld_z2, log_pmu2.
 * nan_flag (may be NULL): set to 1 if any lower_bound is NaN (train_model.py:464-466 guard). */
int fhvae_elbo_fwd(const float* x, const float* xhead, int64_t xs_b, int64_t xs_t, int64_t lv_off,
                   const float* z1head, const float* z2head,
                   const float* mu2, const int64_t* nsegs,
                   float* out5, int* nan_flag,
                   int B, int T, int F, int Z1, int Z2, void* stream);
/* coef (4,B): dL/d{log_px_z, neg_kld_z1, neg_kld_z2, log_pmu2} with dL/dlower_bound already folded
 * in by the host ( c_px = g_lb + g_px, ..., c_pmu2 = g_lb/nsegs + g_pmu2 ).
 * Writes dxhead (same layout as xhead), dz1head, dz2head (B,2Z) (overwrite), dmu2 (B,Z2) (overwrite). */
int fhvae_elbo_bwd(const float* x, const float* xhead, int64_t xs_b, int64_t xs_t, int64_t lv_off,
                   const float* z1head, const float* z2head,
                   const float* mu2, const float* coef,
                   float* dxhead, float* dz1head, float* dz2head, float* dmu2,
                   int B, int T, int F, int Z1, int Z2, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K7: discriminative log q(i|z2) over all table rows (simple_fhvae.py:119-122), never
 * materialising (B,N,Z) or (B,N).  z2mu rows have leading dim ld_z (the [mu|logvar] head).
 * fwd: partial (max, sumexp) per N-split -> part (nsplit,B,2); then combine.  Z in {8,16,32,64}.
 * ------------------------------------------------------------------------------------------- */
int fhvae_disc_nsplit(int B, int64_t N);
int fhvae_disc_fwd_partial(const float* z2mu, int64_t ld_z, const float* table, int64_t N, int Z,
                           float* part, int nsplit, int B, void* stream);
/* tgt (B,) = target logit -||z_b - mu2_b||^2 / (2 s2) in direct form; mu2 (B,Z) = gathered rows. */
int fhvae_disc_target(const float* z2mu, int64_t ld_z, const float* mu2, float* tgt, int B, int Z,
                      void* stream);
/* combine `nparts` partial (max,sumexp) sets (nparts = nsplit, or nsplit*world after an all-gather).
 * Writes log_qy (B,) = tgt - lse, and lse (B,). */
int fhvae_disc_combine(const float* part, int nparts, const float* tgt,
                       float* log_qy, float* lse, int B, void* stream);
/* bwd, g (B,) = dL/dlog_qy, p_bn = exp(s_bn - lse_b):
 *   rows:   dtable (N,Z) overwritten with -sum_b g_b p_bn (z_b - m_n)/s2  (dense softmax part; owner-local)
 *   segs:   sumpm_part (nsplit,B,Z) = per-N-split partial sums of p_bn m_n
 *   finish: dz2mu[b*ld_dz + d] += g_b/s2 * (mu2_b - sum_parts sumpm);  dmu2 (B,Z) += g_b (z_b - mu2_b)/s2
 *           (the sparse target part, reduced into the table by fhvae_mu2_scatter_reduce). */
int fhvae_disc_bwd_rows(const float* z2mu, int64_t ld_z, const float* table, int64_t N, int Z,
                        const float* lse, const float* g, float* dtable, int B, void* stream);
int fhvae_disc_bwd_segs(const float* z2mu, int64_t ld_z, const float* table, int64_t N, int Z,
                        const float* lse, float* sumpm_part, int nsplit, int B, void* stream);
int fhvae_disc_bwd_finish(const float* z2mu, int64_t ld_z, const float* mu2, const float* sumpm_part,
                          int nparts, const float* g, float* dz2mu, int64_t ld_dz, float* dmu2,
                          int B, int Z, void* stream);

/* Sharded table (north star: "the mu2 table is sharded by utterance id with row gradients routed to their owner";
 * row u lives on rank u mod W at local row u / W).  The softmax of simple_fhvae.py:119-122 runs over ALL rows, so
 * per step every rank scores all B_global segments against its rows with the kernels above (B = B_global, table =
 * the local shard) and the partials meet in fhvae_disc_combine_sharded.  Helpers of that exchange:
 *   shard_pack:   packet (B, Z+4) rows [z2_mu | mu_idx as two 32-bit words | g = dL/dlog_qy | pad] -- one all-gather
 *   shard_unpack: gathered packets (Bg rows) -> idx_g (may be NULL), lidx_g = local row on THIS rank or -1, g_g;
 *                 ids < 0 or beyond the shard OR FHVAE_FLAG_BAD_INDEX into *err_flag (may be NULL)
 *   combine:      part (nparts = W*nsplit, Bg, 2) in rank-major order -> lse_g (Bg,) bit-identical on every rank,
 *                 log_qy (B_local,) = tgt - lse for this rank's segments [b_off, b_off + B_local). */
int fhvae_shard_pack(const float* z2mu, int64_t ld_z, const int64_t* idx, const float* g, float* packet, int B, int Z,
                     void* stream);
int fhvae_shard_unpack(const float* packet, int Bg, int Z, int world, int rank, int64_t N_local, int64_t* idx_g,
                       int64_t* lidx_g, float* g_g, int32_t* err_flag, void* stream);
int fhvae_disc_combine_sharded(const float* part, int nparts, int Bg, const float* tgt, int b_off, int B_local,
                               float* log_qy, float* lse_g, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K0: mu2 table (simple_fhvae.py:39-54).  Exact int64 indexing; deterministic reductions.
 * ------------------------------------------------------------------------------------------- */
/* Rows outside [0,N) (torch.gather raises, simple_fhvae.py:53): the output row is NaN (so the lower bound trips
 * the reference's own NaN guard) and FHVAE_FLAG_BAD_INDEX is OR-ed into *err_flag (may be NULL).  The scatter
 * and accumulate kernels never write outside the table: such rows are skipped. */
int fhvae_mu2_gather(const float* table, const int64_t* idx, float* mu2, int B, int Z, int64_t N,
                     int32_t* err_flag, void* stream);
/* dtable[idx[b]] += sum over duplicates (ascending b, fixed order).  touched (B,) int32: 1 at the
 * first occurrence of each distinct row, else 0 (the "rows touched" set). */
int fhvae_mu2_scatter_reduce(const float* dmu2, const int64_t* idx, float* dtable, int32_t* touched,
                             int B, int Z, int64_t N, void* stream);
/* utils.py:45-60 batched: zsum (K,Z) += z2mu rows, cnt (K,) += 1  (deterministic per row), then
 * fhvae_mu2_estimate_finish: table[k] = zsum[k] / (cnt[k] + r) where cnt>0.  */
int fhvae_mu2_accumulate(const float* z2mu, int64_t ld_z, const int64_t* idx, float* zsum, float* cnt,
                         int B, int Z, int64_t K, int32_t* err_flag, void* stream);
int fhvae_mu2_estimate_finish(const float* zsum, const float* cnt, float* table, float r,
                              int64_t K, int Z, void* stream);
/* sparse row write-back / fetch between a master shard and the active cache (hierarchical sampling):
 * dst[dst_rows[i]] = src[src_rows[i]] for i < n; rows with dst_rows[i] < 0 are skipped. */
int fhvae_rows_copy(const float* src, const int64_t* src_rows, float* dst, const int64_t* dst_rows,
                    int64_t n, int Z, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K9: Adam on a flat fp32 buffer (train_model.py:409-411,454: lr 1e-3, betas (0.95,0.999), eps 1e-8).
 * `step` is a DEVICE int32 counter (number of steps already taken); the kernel bumps it at the end
 * so the same launch can be replayed from a CUDA graph.  grad_scale multiplies g (1/world for DP).
 * ------------------------------------------------------------------------------------------- */
int fhvae_adam_flat(float* p, const float* g, float* m, float* v, int64_t n,
                    float lr, float beta1, float beta2, float eps, float grad_scale,
                    int32_t* step, uint32_t* done_counter, void* stream);

/* ---------------------------------------------------------------------------------------------
 * small data-movement helpers
 * ------------------------------------------------------------------------------------------- */
/* Device-side segment feeder (datasets.py:214-223 + :100-105 without the per-item file open / H2D copy):
 * feats (R,F) = all utterances packed row-wise in HBM; seg b = rows [start[b], start[b]+T);
 * out[b,t,f] = (feats[(start[b]+t)*F + f] - mean[f]) * inv_std[f]   (mean / inv_std may be NULL: no MVN). */
int fhvae_gather_segments(const float* feats, const int64_t* start, const float* mean, const float* inv_std,
                          float* out, int B, int T, int F, int64_t R, void* stream);
/* (B,T,F) -> (T,B,F) */
int fhvae_transpose_bt(const float* src, float* dst, int B, int T, int F, void* stream);
/* out[c] = sum_r in[r*ld + c]  (r < R); deterministic.  out2 (may be NULL) receives a second copy
 * (nn.LSTM keeps b_ih and b_hh: both get the same gradient).  Up to FHVAE_COLSUM_MAX_BATCH per launch. */
#define FHVAE_COLSUM_MAX_BATCH 16
typedef struct fhvae_colsum_problem {
    const float* in;
    float*       out;
    float*       out2;
    int64_t      ld;
    int32_t      R, C;
} fhvae_colsum_problem;
int fhvae_colsum_batch(const fhvae_colsum_problem* problems, int n_problems, void* stream);
/* out = a + b */
int fhvae_add2(float* out, const float* a, const float* b, int64_t n, void* stream);
/* dpre = dout * (out > 0), in place on dout */
int fhvae_relu_bwd(float* dout, const float* out, int64_t n, void* stream);
/* y = a*x + y */
int fhvae_axpy(float* y, const float* x, float a, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused latent-head stages between the LSTM stacks (exact fp32, one launch each; heads.cu).
 * Replaces GaussianLayer.forward (simple_fhvae.py:205-216) on the final hidden states and the hoisted
 * time-invariant input projection of the next stack, and their autograd.
 * ------------------------------------------------------------------------------------------- */
/* head (B,2Z) = [src0 | src1] (B, nsrc*H; rows of leading dim ld_src) @ W^T (2Z x nsrc*H) + bias;
 * if eps:  zcat[b*ld_z + zoff + d] = mu + eps[b,d]*exp(.5 logvar)   (d < Z);
 * if Q:    Q (B,NQ) = zcat[b, qoff : qoff+Kq] @ Wq^T (NQ rows of leading dim ld_wq) + bias_q (may be NULL). */
int fhvae_head_fwd(const float* src0, const float* src1, int64_t ld_src, int nsrc, int H, const float* W,
                   const float* bias, float* head, int Z, const float* eps, float* zcat, int64_t ld_z, int zoff,
                   const float* Wq, int64_t ld_wq, const float* bias_q, int qoff, int Kq, float* Q, int NQ, int B,
                   void* stream);
/* if dgsum: dzcat[b, dzoff : dzoff+Kq] (+= if beta) = dgsum (B,NG) @ Wq (NG rows of leading dim ld_wq);
 * if eps:   reparameterisation backward of dzcat[b, roff : roff+Z] into dhead (B,2Z) (fhvae_reparam_bwd);
 * if W:     dh_l (B,H) = dhead @ W[:, l*H : (l+1)*H]   for l < nsrc   (W: 2Z x nsrc*H). */
int fhvae_head_bwd(const float* dgsum, int NG, const float* Wq, int64_t ld_wq, int Kq, float* dzcat, int64_t ld_dz,
                   int dzoff, int beta, const float* head, const float* eps, int Z, int roff, float* dhead,
                   int accumulate, const float* W, int nsrc, int H, float* dh0, float* dh1, int B, void* stream);
/* coef (4,B) = [g_px, g_nk1, g_nk2, g_pmu2] from the upstream gradients gout (6,B) of the six forward outputs
 * (rows: lower_bound, log_px_z, neg_kld_z1, neg_kld_z2, log_pmu2, log_qy; simple_fhvae.py:106-116). */
int fhvae_step_coef(const float* gout, const int64_t* nsegs, float* coef, int detach_px, int prior_grad, int B,
                    void* stream);
/* loss = -mean(lower_bound + alpha*log_qy)   (train_model.py:243-251) */
int fhvae_loss_mean(const float* lower_bound, const float* log_qy, float alpha, int B, float* loss, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Weight-gradient GEMM from pre-split bf16 planes, TMA-fed (gemm_wgrad.cu).
 * Replaces the autograd of nn.Linear / nn.LSTM weights for the long contractions (K = T*B rows):
 *   C[m*ldc + n] = sum_k (A_hi + A_lo)[k][m] * (B_hi + B_lo)[k][n]      (hi*hi + hi*lo + lo*hi, fp32 accumulate)
 * A, B: bf16, element (plane, k, m) at plane*plane_stride + k*ld + m, plane 0 = hi = bf16(x), plane 1 = lo =
 * bf16(x - hi); 16-byte aligned, strides multiples of 8 elements.  mode = FHVAE_MODE_BF16X3 (three products) or
 * FHVAE_MODE_BF16 (hi*hi only).  C is overwritten.
 * ------------------------------------------------------------------------------------------- */
#define FHVAE_WGRAD_MAX_BATCH 8
#define FHVAE_SPLIT_MAX_BATCH 16
typedef struct fhvae_wgrad_problem {
    const void* A;
    const void* B;
    float*      C;
    int32_t     M, N, K, reserved;
    int64_t     lda, a_plane_stride, ldb, b_plane_stride, ldc;
} fhvae_wgrad_problem;
int fhvae_wgrad_planes_batch(const fhvae_wgrad_problem* problems, int n_problems, int mode, void* stream);
/* dst planes (bf16) <- src (rows x cols fp32, leading dim ld_src): hi at dst[r*ld_dst + c], lo at
 * dst[plane_stride + r*ld_dst + c].  cols % 8 == 0. */
typedef struct fhvae_split_problem {
    const float* src;
    void*        dst;
    int64_t      ld_src, ld_dst, plane_stride;
    int32_t      rows, cols;
} fhvae_split_problem;
int fhvae_split_planes_batch(const fhvae_split_problem* problems, int n_problems, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Projection GEMM on pre-split planes (csrc/gemm_proj.cu): C[M,N] = A[M,K] W[N,K]^T (+ bias[N]) -- nn.Linear forward
 * (simple_fhvae.py:130-134, 208-212) for the short-K, large-M products on the critical path of the LSTM model (layer-0
 * input projections of x over all T*B rows, the decoder's Gaussian head).  A and W are bf16 hi/lo planes
 * [2][rows][K] with K contiguous (hi at element r*ld + k, lo at plane_stride + r*ld + k), fed by TMA; K % 16 == 0,
 * leading dimensions and plane strides multiples of 8 elements, 16-byte aligned.  C (fp32, leading dim ldc) is
 * overwritten.  mode = FHVAE_MODE_BF16X3 (three products) or FHVAE_MODE_BF16 (hi*hi only).
 * ------------------------------------------------------------------------------------------- */
#define FHVAE_PROJ_MAX_BATCH 8
typedef struct fhvae_proj_problem {
    const void*  A;
    const void*  W;
    float*       C;
    const float* bias;
    int32_t      M, N, K, reserved;
    int64_t      lda, a_plane_stride, ldw, w_plane_stride, ldc;
} fhvae_proj_problem;
int fhvae_proj_planes_batch(const fhvae_proj_problem* problems, int n_problems, int mode, void* stream);

/* The wavefront recurrence with bf16 hi/lo planes as additional (fwd) or alternative (bwd) outputs: the planes are
 * the TMA operands of fhvae_wgrad_planes_batch, written by the kernel that produces the tensor instead of by a
 * separate fhvae_split_planes_batch pass.  *_planes: bf16 [2][T*B*width], hi at element i, lo at plane_stride + i
 * (may be NULL).  In the backward the fp32 dgates_* may be NULL when the corresponding planes are given.
 * packed (may be NULL): the stack's pre-packed weight operands written by fhvae_lstm_wave_pack from the SAME weights
 * and mode -- the kernels then fill TMEM / shared memory with coalesced loads and bulk copies instead of converting
 * the fp32 weights in every launch (9-15 us of each of the six recurrent launches of a step). */
long long fhvae_lstm_wave_pack_bytes(int H, int nlayers, int mode);
/* W_hh0: layer 0 (or the only layer); W_ih1, W_hh1: layer 1 of a 2-layer stack (else NULL).  packed: 128-byte aligned,
 * fhvae_lstm_wave_pack_bytes bytes.  Must be re-run whenever the weights change (once per optimizer step). */
int fhvae_lstm_wave_pack(const float* W_hh0, const float* W_ih1, const float* W_hh1, void* packed, int H, int nlayers,
                         int mode, void* stream);
int fhvae_lstm_wave_fwd_planes(const float* P0, const float* Q0, const float* W_hh0, float* h0, float* c0, float* acts0,
                               const float* W_ih1, const float* bias1, const float* W_hh1, float* h1, float* c1,
                               float* acts1, void* xchg, void* h0_planes, void* h1_planes, int64_t plane_stride,
                               const void* packed, int T, int B, int H, int nlayers, int mode, void* stream);
int fhvae_lstm_wave_bwd_planes(const float* dh_all_top, const float* dh_last_top, const float* dh_last_bot,
                               const float* W_hh_top, const float* c_top, const float* acts_top, float* dgates_top,
                               float* dgsum_top, const float* W_ih_top, const float* W_hh_bot, const float* c_bot,
                               const float* acts_bot, float* dgates_bot, float* dgsum_bot, void* xchg,
                               void* dg_top_planes, void* dg_bot_planes, int64_t plane_stride, const void* packed,
                               int T, int B, int H, int nlayers, int mode, void* stream);

/* out[0..n) ~ N(0,1): the eps draws of GaussianLayer.forward (torch.randn_like, simple_fhvae.py:214) as a kernel of this
 * library (Philox4x32-10 keyed by `seed`, Box-Muller).  `offset` is a DEVICE counter (zero-initialised by the caller)
 * that the launch advances, so replays of a captured launch draw fresh numbers; `done_counter`: zero-initialised
 * device uint32 scratch.  Not bit-compatible with torch's generator (nothing in the reference pins the draws). */
int fhvae_randn(float* out, int64_t n, uint64_t seed, unsigned long long* offset, uint32_t* done_counter, void* stream);

/* x (B,T,F), mu_idx (B int64, may be NULL), num_segs (B int64, may be NULL) -> the step's static input buffers, one
 * launch (device-to-device; the batch of train_model.py:444-445 into CUDA-graph-stable storage).  x is also written
 * time-major (T,B,F) into x_tm when x_tm != NULL (what fhvae_transpose_bt does as a separate launch). */
int fhvae_load_inputs(const float* x_src, float* x_dst, float* x_tm, int B, int T, int F, const int64_t* idx_src,
                      int64_t* idx_dst, const int64_t* nsegs_src, int64_t* nsegs_dst, void* stream);

/* fhvae_elbo_fwd + fhvae_step_coef + fhvae_elbo_bwd as one launch, for a step whose upstream gradients gout (6,B)
 * are known before the forward (the fused train step).  Bit-identical to the three separate calls. */
int fhvae_elbo_fwd_bwd(const float* x, const float* xhead, int64_t xs_b, int64_t xs_t, int64_t lv_off,
                       const float* z1head, const float* z2head, const float* mu2, const int64_t* nsegs,
                       const float* gout, int detach_px, int prior_grad, float* out5, int* nan_flag, float* dxhead,
                       float* dz1head, float* dz2head, float* dmu2, int B, int T, int F, int Z1, int Z2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FHVAE_B200_H */
