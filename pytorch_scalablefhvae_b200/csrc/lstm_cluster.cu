// Persistent thread-block-cluster LSTM recurrence (W_hh resident in shared memory across the
// cluster, tcgen05 MMA, gates fused in the epilogue).  Lands after the SIMT path is parity-green.
#include "common.cuh"
