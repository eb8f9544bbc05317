/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Philox4x32-10 counter-based generator, shared by the CPU oracle and (as an
 * independent restatement of the same published algorithm, Salmon et al. SC'11)
 * by the CUDA kernels.  It replaces Julia's `rand()` stream, which cannot be
 * reproduced without Julia: every uniform the reference draws
 * inside the sweep (src/pmdi.jl:253, src/misc.jl:28,43, src/pmdi.jl:350,367)
 * is addressed here by (seed, iteration, kind, step, k, index) so that both
 * sides consume identical draws at any problem size.
 */
#ifndef PMDI_ORACLE_PHILOX_H
#define PMDI_ORACLE_PHILOX_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* kinds of draw (c2 high byte) */
enum {
  OR_DRAW_ALLOC   = 0, /* allocation uniform, index = particle (src/pmdi.jl:253)        */
  OR_DRAW_RESAMP  = 1, /* systematic-resampling offset, index = 0 (src/misc.jl:28)      */
  OR_DRAW_SHUFFLE = 2, /* Fisher-Yates pick for position index (src/misc.jl:43)         */
  OR_DRAW_SELECT  = 3, /* p_star uniform, index = 0 (src/pmdi.jl:350)                   */
  OR_DRAW_FEATURE = 4  /* feature-flag uniform, index = feature (src/pmdi.jl:367)       */
};

static inline void or_philox4x32_10(uint32_t c[4], const uint32_t key[2]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c[0];
    uint64_t p1 = (uint64_t)M1 * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += W0; k1 += W1;
  }
}

/* uniform in [0,1) with 53 random bits */
static inline double or_uniform(uint64_t seed, uint32_t iter, uint32_t kind,
                                uint32_t step, uint32_t k, uint32_t index) {
  uint32_t c[4] = { index, step, (kind << 16) | (k & 0xFFFFu), iter };
  uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
  or_philox4x32_10(c, key);
  uint64_t x = ((uint64_t)c[0] << 32) | c[1];
  return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

#ifdef __cplusplus
}
#endif
#endif
