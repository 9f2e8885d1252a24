/*
 * notorch_b200 — debug / test-only entry points of libnotorch_b200.so. NOT part of the drop-in boundary (include/notorch_b200.h):
 * nothing on the data path calls them; they exist for the role-timeline traces under profiles/ and for the host-only tests of the
 * weight-gradient work decomposition.
 */
#ifndef NOTORCH_B200_DEBUG_H_
#define NOTORCH_B200_DEBUG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Debug only: device buffer of >= 65001 uint64 (word 0 = counter, zeroed by the caller) into which CTA 0 of the fused
 * tensor-core kernel appends role/time records; NULL switches tracing off. Not part of the data path. */
void nt_debug_set_trace_buffer(void* device_u64_buffer);
/* Host-only (no CUDA call): the work decomposition nt_layer_backward_wgrad (CTA-pair kernel) would use for E edges, hidden size d
 * and num_sms SMs. out12 = {m_units, n_tiles, n_tile, n_a, n_b, half_last, full_units, half_units, splits, splits_last,
 * k_blocks_per_split, k_blocks_per_split_last}; a K-block is 32 edges. For tests of the split logic. */
int nt_debug_wgrad_geometry(int64_t E, int64_t d, int num_sms, int64_t* out12);

/* Host-only (no CUDA call): the (row, chunk) split of the flattened item index t = row * chunks + chunk as the row kernels compute
 * it - a multiply-high by ceil(2^64 / chunks) when the item count `total` fits 32 bits (magic = 0: the general 64-bit division).
 * out3 = {magic != 0, row, chunk}. For the CPU test of that arithmetic. */
int nt_debug_split_item(int64_t total, int64_t chunks, int64_t t, int64_t* out3);

#ifdef __cplusplus
}
#endif
#endif /* NOTORCH_B200_DEBUG_H_ */
