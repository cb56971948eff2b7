// Persistent W_hh-resident recurrent layer on tcgen05 (bf16 operands, fp32 TMEM
// accumulators).  Placeholder until the kernel lands: reports "shape not supported"
// so slnlp_rnn_layer_fwd uses the general fp32 CUDA path (never a CPU path).
#include "common.cuh"

namespace slnlp {
int rnn_layer_fwd_tc(int, int, int, int, int, float*, const float*, const float*, const int64_t*,
                     const float*, const float*, float*, float*, float*, cudaStream_t) {
  return -1;
}
}  // namespace slnlp
