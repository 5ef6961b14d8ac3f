// synth.cu -- synthetic corpus generator entry points (bench / test infrastructure).
#include "kernels.h"
#include "synth_gen.h"

__global__ void __launch_bounds__(128) k_synth(int shape, u64 seed, u64 first_block, uint8_t *__restrict__ out, u64 n, u64 n_blocks) {
    for (u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += (u64)gridDim.x * blockDim.x) {
        u64 base = b * SYNTH_BLOCK;
        u32 limit = (u32)(n - base < SYNTH_BLOCK ? n - base : SYNTH_BLOCK);
        synth_block(shape, seed, first_block + b, out + base, limit);
    }
}

BPE_API int bpe_synth_dev_at(bpe_ctx *ctx, int shape, uint64_t seed, uint64_t first_block, uint8_t *out_dev, uint64_t n) {
    if (!ctx || (!out_dev && n) || (shape != BPE_SYNTH_TINYSTORIES && shape != BPE_SYNTH_OWT)) return BPE_ERR_ARG;
    if (!n) return BPE_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    u64 nb = (n + SYNTH_BLOCK - 1) / SYNTH_BLOCK;
    u64 grid = (nb + 127) / 128;
    if (grid > (u64)ctx->sm_count * 16) grid = (u64)ctx->sm_count * 16;
    KLAUNCH(k_synth, (unsigned)grid, 128, 0, ctx->stream, shape, seed, first_block, out_dev, n, nb);
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return BPE_OK;
}

BPE_API int bpe_synth_dev(bpe_ctx *ctx, int shape, uint64_t seed, uint8_t *out_dev, uint64_t n) {
    return bpe_synth_dev_at(ctx, shape, seed, 0, out_dev, n);
}

BPE_API int bpe_synth_host_at(int shape, uint64_t seed, uint64_t first_block, uint8_t *out_host, uint64_t n) {
    if ((!out_host && n) || (shape != BPE_SYNTH_TINYSTORIES && shape != BPE_SYNTH_OWT)) return BPE_ERR_ARG;
    u64 nb = (n + SYNTH_BLOCK - 1) / SYNTH_BLOCK;
    for (u64 b = 0; b < nb; b++) {
        u64 base = b * SYNTH_BLOCK;
        u32 limit = (u32)(n - base < SYNTH_BLOCK ? n - base : SYNTH_BLOCK);
        synth_block(shape, seed, first_block + b, out_host + base, limit);
    }
    return BPE_OK;
}

BPE_API int bpe_synth_host(int shape, uint64_t seed, uint8_t *out_host, uint64_t n) { return bpe_synth_host_at(shape, seed, 0, out_host, n); }
