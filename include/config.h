/* config.h — default problem shape of the reference (include/config.h:10-28).
 *
 * In the reference these constexprs ARE the problem: N, d_model, h and the tile sizes are
 * compiled into every kernel.  Here they are only the defaults the profile_* driver uses when
 * no --N/--d_model/--h flag is given; solve() honours its run-time arguments. */
#ifndef QMHA_CONFIG_H
#define QMHA_CONFIG_H

constexpr int N = 8192;        /* sequence length        (reference config.h:22) */
constexpr int d_model = 1024;  /* model dimension        (reference config.h:23) */
constexpr int h = 32;          /* attention heads        (reference config.h:24) */
static_assert(d_model % h == 0, "d_model must be divisible by h");
constexpr int d = d_model / h; /* per-head dimension = 32 (reference config.h:28) */

#endif
