// batch.cu -- token-array consumers next to the tokenizer.
//
//   bpe_batch_windows_dev   replaces the per-row Python loop of load_batch (models/util.py:37-57 of the reference):
//                           inputs[row] = dataset[idx : idx + ctx], targets[row] = dataset[idx + 1 : idx + ctx + 1],
//                           both as int64 (torch.long), for `batch` start indices chosen by the caller (the reference
//                           draws them with torch.randint on the host, so the caller keeps doing that: same generator,
//                           same batches).
//
// HBM-bound: 2 x 2 B read (the second read of a token hits L1/L2) and 16 B written per element; one element per thread,
// consecutive threads on consecutive columns of a row, so reads and writes are coalesced.
#include <algorithm>
#include "kernels.h"
#include "ctx.h"

template <typename T>
__global__ void __launch_bounds__(256) k_batch_windows(const T *__restrict__ tok, const long long *__restrict__ starts, u32 batch,
                                                      u32 context, long long *__restrict__ x, long long *__restrict__ y) {
    const u64 total = (u64)batch * context;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (u64)gridDim.x * blockDim.x) {
        const u32 row = (u32)(i / context), col = (u32)(i - (u64)row * context);
        const T *p = tok + starts[row] + col;
        x[i] = (long long)p[0];
        y[i] = (long long)p[1];
    }
}

BPE_API int bpe_batch_windows_dev(bpe_ctx *ctx, const void *tokens_dev, int dtype, uint64_t n_tokens, const int64_t *starts_host,
                                  uint32_t batch, uint32_t context, int64_t *x_dev, int64_t *y_dev) {
    if (!ctx) return BPE_ERR_ARG;
    if ((!tokens_dev && n_tokens) || (!starts_host && batch) || (batch && context && (!x_dev || !y_dev)))
        return bpe_set_error(ctx, BPE_ERR_ARG, "bpe_batch_windows_dev: null argument");
    if (dtype != BPE_DTYPE_U16 && dtype != BPE_DTYPE_I32) return bpe_set_error(ctx, BPE_ERR_ARG, "bpe_batch_windows_dev: dtype must be BPE_DTYPE_U16 or BPE_DTYPE_I32");
    if (batch == 0 || context == 0) return BPE_OK;
    for (uint32_t r = 0; r < batch; r++)       // dataset[idx + 1 : idx + ctx + 1] must exist (the reference draws idx < len - ctx)
        if (starts_host[r] < 0 || (uint64_t)starts_host[r] + context + 1 > n_tokens) {
            ctx->err_detail = (int64_t)r;
            return bpe_set_error(ctx, BPE_ERR_ARG, "bpe_batch_windows_dev: window %u (start %lld, context %u) leaves the %llu tokens", r,
                                 (long long)starts_host[r], context, (unsigned long long)n_tokens);
        }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    BPE_TRY(bpe_buf_reserve(ctx, ctx->tmp0, round_up((size_t)batch * 8, 256)));
    // (pageable source: cudaMemcpyAsync stages it before returning, the caller's array may be reused at once)
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->tmp0.p, starts_host, (size_t)batch * 8, cudaMemcpyHostToDevice, st));
    const u64 total = (u64)batch * context;
    const unsigned grid = (unsigned)std::min<u64>((u64)ctx->sm_count * bpe_grid_mult(64), (total + 255) / 256);
    if (dtype == BPE_DTYPE_U16)
        KLAUNCH(k_batch_windows<uint16_t>, grid, 256, 0, st, (const uint16_t *)tokens_dev, (const long long *)ctx->tmp0.p, batch, context, (long long *)x_dev, (long long *)y_dev);
    else
        KLAUNCH(k_batch_windows<int32_t>, grid, 256, 0, st, (const int32_t *)tokens_dev, (const long long *)ctx->tmp0.p, batch, context, (long long *)x_dev, (long long *)y_dev);
    CUDA_TRY(ctx, cudaGetLastError());
    // On the caller's stream (bpe_ctx_set_stream) the result is stream-ordered with the caller's other work and nothing waits
    // here; on the context's private stream the call returns with x / y complete.
    if (st == ctx->own_stream) CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return BPE_OK;
}
