// C-ABI plumbing: error string, version, and the mode dispatch of the GEMM / LSTM entry points.
#include <stdarg.h>
#include <stdlib.h>
#include <atomic>
#include "common.cuh"

namespace fhvae {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
bool pdl_enabled(int family) {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("FHVAE_PDL");
        v = e ? atoi(e) : 31;
    }
    return (v & family) != 0;
}
static std::atomic<int> g_deterministic{0};
bool deterministic_mode() { return g_deterministic.load(std::memory_order_relaxed) != 0; }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int gemm_batch_simt(const fhvae_gemm_problem* problems, int n, cudaStream_t st);
int gemm_batch_tc(const fhvae_gemm_problem* problems, int n, int mode, cudaStream_t st);
int lstm_fwd_simt(const float* P, const float* Q, const float* W_hh, float* h_all, float* c_all,
                  float* acts, int T, int B, int H, cudaStream_t st);
int lstm_bwd_simt(const float* dh_all, const float* dh_last, const float* W_hh, const float* c_all,
                  const float* acts, float* dgates, float* dgsum, float* dc, int T, int B, int H,
                  cudaStream_t st);

bool lstm_cluster_supported(int B, int H);
int lstm_fwd_cluster(const float* P, const float* Q, const float* W_hh, float* h_all, float* c_all,
                     float* acts, float* xchg, int T, int B, int H, int mode, cudaStream_t st);

int lstm_bwd_cluster(const float* dh_all, const float* dh_last, const float* W_hh, const float* c_all,
                     const float* acts, float* dgates, float* dgsum, float* xchg, int T, int B, int H, int mode,
                     cudaStream_t st);

bool lstm_wave_supported(int T, int B, int H, int L);
size_t lstm_wave_xchg_bytes(int T, int B, int H, int L);
int lstm_wave_rows_per_launch(int H, int L);
int lstm_wave_fwd(const float* P0, const float* Q0, const float* Whh0, float* h0, float* c0, float* a0,
                  const float* Wih1, const float* b1, const float* Whh1, float* h1, float* c1, float* a1,
                  void* xchg, int T, int B, int H, int L, int mode, cudaStream_t st, void* hp0 = nullptr, void* hp1 = nullptr,
                  long long hps = 0, const void* packed = nullptr);
size_t lstm_wave_pack_bytes(int H, int L, int mode);
int lstm_wave_pack(const float* Whh_l0, const float* Wih1, const float* Whh_l1, void* out, int H, int L, int mode, cudaStream_t st);
size_t lstm_wave_bwd_xchg_bytes(int T, int B, int H, int L);
int lstm_wave_bwd(const float* dh_all, const float* dh_last1, const float* dh_last0, const float* Whh1, const float* c1,
                  const float* a1, float* dg1, float* dgsum1, const float* Wih1, const float* Whh0, const float* c0,
                  const float* a0, float* dg0, float* dgsum0, void* xchg, int T, int B, int H, int L, int mode, cudaStream_t st,
                  void* dgp1 = nullptr, void* dgp0 = nullptr, long long dgps = 0, const void* packed = nullptr);

}  // namespace fhvae

using namespace fhvae;

extern "C" const char* fhvae_last_error_string(void) { return g_err; }
extern "C" int fhvae_version(void) { return 1; }
extern "C" unsigned long long fhvae_launch_count(void) { return g_launches.load(); }
extern "C" int fhvae_built_for_sm(void) { return 100; }

extern "C" int fhvae_gemm_batch(const fhvae_gemm_problem* problems, int n_problems, int mode,
                                void* stream) {
    FHVAE_CHECK_ARG(problems && n_problems > 0 && n_problems <= FHVAE_GEMM_MAX_BATCH,
                    "gemm_batch: need 1..%d problems", FHVAE_GEMM_MAX_BATCH);
    for (int i = 0; i < n_problems; ++i) {
        const fhvae_gemm_problem& p = problems[i];
        FHVAE_CHECK_ARG(p.A && p.B && p.C && p.M >= 0 && p.N >= 0 && p.K >= 0,
                        "gemm_batch: problem %d has a null pointer or negative size", i);
        FHVAE_CHECK_ARG((p.sa_m == 1 || p.sa_k == 1) && (p.sb_k == 1 || p.sb_n == 1),
                        "gemm_batch: problem %d: one stride of A and of B must be 1", i);
    }
    if (mode == FHVAE_MODE_F32_SIMT) return gemm_batch_simt(problems, n_problems, as_stream(stream));
    if (mode == FHVAE_MODE_BF16X3 || mode == FHVAE_MODE_BF16)
        return gemm_batch_tc(problems, n_problems, mode, as_stream(stream));
    set_error("gemm_batch: unknown mode %d", mode);
    return FHVAE_EINVAL;
}

extern "C" int fhvae_lstm_fwd(const float* P, const float* Q, const float* W_hh, float* h_all,
                              float* c_all, float* acts, float* xchg, int T, int B, int H, int mode,
                              void* stream) {
    FHVAE_CHECK_ARG(W_hh && h_all && c_all && acts && (P || Q), "lstm_fwd: null pointer");
    FHVAE_CHECK_ARG(T > 0 && B > 0 && H > 0 && H % 8 == 0, "lstm_fwd: need T,B>0 and H %% 8 == 0");
    if (mode != FHVAE_MODE_F32_SIMT && lstm_cluster_supported(B, H)) {
        FHVAE_CHECK_ARG(xchg, "lstm_fwd: the cluster kernel needs the 16*B*H-float exchange scratch");
        return lstm_fwd_cluster(P, Q, W_hh, h_all, c_all, acts, xchg, T, B, H, mode, as_stream(stream));
    }
    return lstm_fwd_simt(P, Q, W_hh, h_all, c_all, acts, T, B, H, as_stream(stream));
}

extern "C" int fhvae_lstm_bwd(const float* dh_all, const float* dh_last, const float* W_hh,
                              const float* c_all, const float* acts, float* dgates, float* dgsum,
                              float* dh_rec, float* dc, int T, int B, int H, int mode, void* stream) {
    FHVAE_CHECK_ARG(W_hh && c_all && acts && dgates && dc && (dh_all || dh_last), "lstm_bwd: null pointer");
    FHVAE_CHECK_ARG(T > 0 && B > 0 && H > 0 && H % 8 == 0, "lstm_bwd: need T,B>0 and H %% 8 == 0");
    if (mode != FHVAE_MODE_F32_SIMT && lstm_cluster_supported(B, H)) {
        FHVAE_CHECK_ARG(dh_rec, "lstm_bwd: the cluster kernel needs the 16*B*H-float exchange scratch (dh_rec)");
        return lstm_bwd_cluster(dh_all, dh_last, W_hh, c_all, acts, dgates, dgsum, dh_rec, T, B, H, mode,
                                as_stream(stream));
    }
    return lstm_bwd_simt(dh_all, dh_last, W_hh, c_all, acts, dgates, dgsum, dc, T, B, H,
                         as_stream(stream));
}

extern "C" int fhvae_set_deterministic(int on) {
    const int prev = g_deterministic.exchange(on ? 1 : 0);
    return prev;
}
extern "C" int fhvae_get_deterministic() { return g_deterministic.load() ? 1 : 0; }

static constexpr int WAVE_MIN_B = 32;

extern "C" int fhvae_lstm_wave_supported(int T, int B, int H, int nlayers, int mode) {
    return (mode == FHVAE_MODE_BF16X3 || mode == FHVAE_MODE_BF16) && lstm_wave_supported(T, B, H, nlayers) ? 1 : 0;
}

extern "C" int fhvae_lstm_wave_rows_per_launch(int H, int nlayers) { return lstm_wave_rows_per_launch(H, nlayers); }

extern "C" long long fhvae_lstm_wave_xchg_bytes(int T, int B, int H, int nlayers) {
    if (!lstm_wave_supported(T, B, H, nlayers)) return 0;
    return (long long)lstm_wave_xchg_bytes(T, B, H, nlayers);
}

extern "C" int fhvae_lstm_wave_fwd(const float* P0, const float* Q0, const float* W_hh0, float* h0, float* c0,
                                   float* acts0, const float* W_ih1, const float* bias1, const float* W_hh1,
                                   float* h1, float* c1, float* acts1, void* xchg, int T, int B, int H,
                                   int nlayers, int mode, void* stream) {
    FHVAE_CHECK_SUP(fhvae_lstm_wave_supported(T, B, H, nlayers, mode),
                    "lstm_wave_fwd: needs a tensor-core mode, H in {128, 256}, B %% 32 == 0, T <= 63, 1 or 2 layers");
    FHVAE_CHECK_ARG(W_hh0 && h0 && c0 && acts0 && xchg && (P0 || Q0), "lstm_wave_fwd: null pointer (layer 0)");
    FHVAE_CHECK_ARG(nlayers == 1 || (W_ih1 && W_hh1 && h1 && c1 && acts1), "lstm_wave_fwd: null pointer (layer 1)");
    return lstm_wave_fwd(P0, Q0, W_hh0, h0, c0, acts0, W_ih1, bias1, W_hh1, h1, c1, acts1, xchg, T, B, H, nlayers,
                         mode, as_stream(stream));
}

extern "C" int fhvae_lstm_wave_fwd_planes(const float* P0, const float* Q0, const float* W_hh0, float* h0, float* c0,
                                          float* acts0, const float* W_ih1, const float* bias1, const float* W_hh1,
                                          float* h1, float* c1, float* acts1, void* xchg, void* h0_planes,
                                          void* h1_planes, int64_t plane_stride, const void* packed, int T, int B, int H,
                                          int nlayers, int mode, void* stream) {
    FHVAE_CHECK_SUP(fhvae_lstm_wave_supported(T, B, H, nlayers, mode),
                    "lstm_wave_fwd: needs a tensor-core mode, H in {128, 256}, B %% 32 == 0, T <= 63, 1 or 2 layers");
    FHVAE_CHECK_ARG(W_hh0 && h0 && c0 && acts0 && xchg && (P0 || Q0), "lstm_wave_fwd: null pointer (layer 0)");
    FHVAE_CHECK_ARG(nlayers == 1 || (W_ih1 && W_hh1 && h1 && c1 && acts1), "lstm_wave_fwd: null pointer (layer 1)");
    return lstm_wave_fwd(P0, Q0, W_hh0, h0, c0, acts0, W_ih1, bias1, W_hh1, h1, c1, acts1, xchg, T, B, H, nlayers,
                         mode, as_stream(stream), h0_planes, nlayers == 2 ? h1_planes : nullptr, plane_stride, packed);
}

extern "C" long long fhvae_lstm_wave_pack_bytes(int H, int nlayers, int mode) {
    if (!fhvae_lstm_wave_supported(1, WAVE_MIN_B, H, nlayers, mode)) return 0;
    return (long long)lstm_wave_pack_bytes(H, nlayers, mode);
}

extern "C" int fhvae_lstm_wave_pack(const float* W_hh0, const float* W_ih1, const float* W_hh1, void* packed, int H,
                                    int nlayers, int mode, void* stream) {
    FHVAE_CHECK_SUP(fhvae_lstm_wave_supported(1, WAVE_MIN_B, H, nlayers, mode),
                    "lstm_wave_pack: needs a tensor-core mode, H in {128, 256}, 1 or 2 layers");
    FHVAE_CHECK_ARG(W_hh0 && packed && (nlayers == 1 || (W_ih1 && W_hh1)), "lstm_wave_pack: null pointer");
    FHVAE_CHECK_ARG((reinterpret_cast<uintptr_t>(packed) & 127) == 0, "lstm_wave_pack: the image buffer must be 128-byte aligned");
    return lstm_wave_pack(W_hh0, W_ih1, W_hh1, packed, H, nlayers, mode, as_stream(stream));
}

extern "C" long long fhvae_lstm_wave_bwd_xchg_bytes(int T, int B, int H, int nlayers) {
    if (!lstm_wave_supported(T, B, H, nlayers)) return 0;
    return (long long)lstm_wave_bwd_xchg_bytes(T, B, H, nlayers);
}

extern "C" int fhvae_lstm_wave_bwd(const float* dh_all_top, const float* dh_last_top, const float* dh_last_bot,
                                   const float* W_hh_top, const float* c_top, const float* acts_top, float* dgates_top,
                                   float* dgsum_top, const float* W_ih_top, const float* W_hh_bot, const float* c_bot,
                                   const float* acts_bot, float* dgates_bot, float* dgsum_bot, void* xchg, int T, int B,
                                   int H, int nlayers, int mode, void* stream) {
    FHVAE_CHECK_SUP(fhvae_lstm_wave_supported(T, B, H, nlayers, mode),
                    "lstm_wave_bwd: needs a tensor-core mode, H in {128, 256}, B %% 32 == 0, T <= 63, 1 or 2 layers");
    FHVAE_CHECK_ARG(W_hh_top && c_top && acts_top && dgates_top && xchg, "lstm_wave_bwd: null pointer (top layer)");
    FHVAE_CHECK_ARG(nlayers == 1 || (W_ih_top && W_hh_bot && c_bot && acts_bot && dgates_bot),
                    "lstm_wave_bwd: null pointer (bottom layer)");
    FHVAE_CHECK_ARG(dh_all_top || dh_last_top || (nlayers == 2 && dh_last_bot), "lstm_wave_bwd: no incoming gradient");
    return lstm_wave_bwd(dh_all_top, dh_last_top, dh_last_bot, W_hh_top, c_top, acts_top, dgates_top, dgsum_top, W_ih_top,
                         W_hh_bot, c_bot, acts_bot, dgates_bot, dgsum_bot, xchg, T, B, H, nlayers, mode, as_stream(stream));
}

extern "C" int fhvae_lstm_wave_bwd_planes(const float* dh_all_top, const float* dh_last_top, const float* dh_last_bot,
                                          const float* W_hh_top, const float* c_top, const float* acts_top,
                                          float* dgates_top, float* dgsum_top, const float* W_ih_top,
                                          const float* W_hh_bot, const float* c_bot, const float* acts_bot,
                                          float* dgates_bot, float* dgsum_bot, void* xchg, void* dg_top_planes,
                                          void* dg_bot_planes, int64_t plane_stride, const void* packed, int T, int B, int H,
                                          int nlayers, int mode, void* stream) {
    FHVAE_CHECK_SUP(fhvae_lstm_wave_supported(T, B, H, nlayers, mode),
                    "lstm_wave_bwd: needs a tensor-core mode, H in {128, 256}, B %% 32 == 0, T <= 63, 1 or 2 layers");
    FHVAE_CHECK_ARG(W_hh_top && c_top && acts_top && (dgates_top || dg_top_planes) && xchg,
                    "lstm_wave_bwd: null pointer (top layer)");
    FHVAE_CHECK_ARG(nlayers == 1 || (W_ih_top && W_hh_bot && c_bot && acts_bot && (dgates_bot || dg_bot_planes)),
                    "lstm_wave_bwd: null pointer (bottom layer)");
    FHVAE_CHECK_ARG(dh_all_top || dh_last_top || (nlayers == 2 && dh_last_bot), "lstm_wave_bwd: no incoming gradient");
    return lstm_wave_bwd(dh_all_top, dh_last_top, dh_last_bot, W_hh_top, c_top, acts_top, dgates_top, dgsum_top, W_ih_top,
                         W_hh_bot, c_bot, acts_bot, dgates_bot, dgsum_bot, xchg, T, B, H, nlayers, mode, as_stream(stream),
                         dg_top_planes, nlayers == 2 ? dg_bot_planes : nullptr, plane_stride, packed);
}
