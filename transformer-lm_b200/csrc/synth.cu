// synth.cu -- placeholder until the synthetic corpus generator lands.
#include "kernels.h"
BPE_API int bpe_synth_dev(bpe_ctx *, int, uint64_t, uint8_t *, uint64_t) { return BPE_ERR_UNSUPPORTED; }
BPE_API int bpe_synth_host(int, uint64_t, uint8_t *, uint64_t) { return BPE_ERR_UNSUPPORTED; }
