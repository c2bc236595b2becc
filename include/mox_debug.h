/* mox_debug.h — test hooks exported by libmox.so next to the drop-in ABI of mox.h.  They have no
 * counterpart in the reference; tests/ uses them to check GPU building blocks in isolation. */
#ifndef MOX_DEBUG_H
#define MOX_DEBUG_H
#include "mox.h"
#ifdef __cplusplus
extern "C" {
#endif
/* Stable LSD radix sort of (key, value) pairs with the builder's onesweep kernels.
 * Host pointers; sorted in place. */
int mox_debug_radix_sort(mox_ctx*, uint32_t* keys, uint32_t* vals, size_t n);
#ifdef __cplusplus
}
#endif
#endif
