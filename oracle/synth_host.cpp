// synth_host.cpp -- host-only build of the synthetic corpus generator (bench / test infrastructure).
//
// The same header the CUDA library compiles (transformer-lm_b200/csrc/synth_gen.h: integer-only, block-addressable), built
// with g++ into a library of its own so that the CPU arms of bench.py (--impl reference, cpu_baseline) generate their
// input without loading the product's CUDA library.
#include <stdint.h>
#include "../transformer-lm_b200/csrc/synth_gen.h"

extern "C" __attribute__((visibility("default")))
int synth_host_at(int shape, uint64_t seed, uint64_t first_block, uint8_t *out, uint64_t n) {
    if ((!out && n) || (shape != 0 && shape != 1)) return -1;
    const uint64_t nb = (n + SYNTH_BLOCK - 1) / SYNTH_BLOCK;
    for (uint64_t b = 0; b < nb; b++) {
        const uint64_t base = b * SYNTH_BLOCK;
        const uint32_t limit = (uint32_t)(n - base < SYNTH_BLOCK ? n - base : SYNTH_BLOCK);
        synth_block(shape, seed, first_block + b, out + base, limit);
    }
    return 0;
}
