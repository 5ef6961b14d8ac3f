// encode.cu -- placeholder until the encoder lands (every entry point reports BPE_ERR_UNSUPPORTED).
#include "kernels.h"
#include "ctx.h"
struct bpe_tok { bpe_ctx *ctx; };
BPE_API int bpe_tok_create(bpe_ctx *ctx, const int32_t *, const int32_t *, const int32_t *, int, const int32_t *, int,
                           const uint8_t *, const uint64_t *, const int64_t *, int64_t, const uint8_t *, const uint32_t *,
                           const int64_t *, int, bpe_tok **out) { if (out) *out = nullptr; return bpe_set_error(ctx, BPE_ERR_UNSUPPORTED, "encoder not built"); }
BPE_API void bpe_tok_destroy(bpe_tok *) {}
BPE_API int bpe_encode(bpe_tok *, const uint8_t *, uint64_t, int, void *, uint64_t, uint64_t *, bpe_encode_stats *) { return BPE_ERR_UNSUPPORTED; }
BPE_API int bpe_encode_dev(bpe_tok *, const uint8_t *, uint64_t, int, void *, uint64_t, uint64_t *, bpe_encode_stats *) { return BPE_ERR_UNSUPPORTED; }
BPE_API int bpe_tok_key_error(bpe_tok *, uint8_t *, uint64_t, uint64_t *) { return BPE_ERR_UNSUPPORTED; }
BPE_API int bpe_tok_cache_reset(bpe_tok *) { return BPE_ERR_UNSUPPORTED; }
BPE_API int bpe_decode(bpe_tok *, const int64_t *, uint64_t, uint8_t *, uint64_t, uint64_t *) { return BPE_ERR_UNSUPPORTED; }
