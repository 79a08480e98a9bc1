// prims.h — hand-written device-wide primitives used by the tessellator and the binner. One kernel launch each.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

// In-place exclusive scan of `rows` (<= 16) independent rows of `n_plus_1` u32 each: on return row[i] = sum(row_in[0..i)),
// i in [0, n_plus_1); the input value in the last slot is ignored (it receives the row total). Single pass, decoupled
// look-back. `scratch` (cr_scan_scratch_words words) must have been zeroed once with cr_scan_prepare after its allocation;
// it may be shared by scans of different shapes as long as they run one after the other on the stream.
int cr_scan_exclusive(cudaStream_t stream, uint32_t* data, uint32_t n_plus_1, uint32_t rows, uint32_t* scratch);
uint32_t cr_scan_scratch_words(uint32_t n_plus_1, uint32_t rows);
int cr_scan_prepare(cudaStream_t stream, uint32_t* scratch, size_t words);

// Stable LSD radix sort of (key, value) u32 pairs on the low `key_bits` bits of the key. The number of live pairs is read
// from the device word *n_ptr (<= capacity, which sizes the grids). The sorted result is returned through *keys_out /
// *vals_out, which alias either the input or the alt buffers (ping-pong).
int cr_radix_sort_pairs(cudaStream_t stream, uint32_t* keys, uint32_t* vals, uint32_t* keys_alt, uint32_t* vals_alt, uint32_t capacity, const uint32_t* n_ptr,
                        uint32_t key_bits, uint32_t* scratch, uint32_t** keys_out, uint32_t** vals_out);
uint32_t cr_radix_scratch_words(uint32_t capacity);

// begin[t] = first index i in [0, *n_ptr] with i == *n_ptr or sorted_keys[i] >= t, for t in [0, n_keys_plus_1). `capacity` sizes the grid.
int cr_lower_bounds(cudaStream_t stream, const uint32_t* sorted_keys, uint32_t capacity, const uint32_t* n_ptr, uint32_t* begin, uint32_t n_keys_plus_1);

extern unsigned long long g_cr_kernel_launches;   // counted by every launcher of this library
