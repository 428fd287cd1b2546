// placeholder until gemm_tc.cu (tcgen05) lands: tensor-core modes fall through to an error
#include "common.cuh"
namespace fhvae {
int gemm_batch_tc(const fhvae_gemm_problem*, int, int mode, cudaStream_t) {
    set_error("gemm_batch: tensor-core mode %d not built", mode);
    return FHVAE_ENOSUP;
}
}
