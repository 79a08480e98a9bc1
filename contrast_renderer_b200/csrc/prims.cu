// prims.cu — exclusive scan, stable LSD radix sort and sorted-key range search, hand-written for sm_100a.
// These are the "sizes and order" plumbing between the count and emit passes of the tessellator and between the
// binner and the tile rasteriser (the stable sort is what carries the reference's draw order into every tile).
#include "device_common.cuh"
#include "prims.h"

unsigned long long g_cr_kernel_launches = 0;

namespace {

#define SCAN_BLOCK 1024
__device__ __forceinline__ uint32_t block_inclusive_scan(uint32_t v, uint32_t* warp_sums) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (uint32_t)o) x += y; }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= (uint32_t)o) w += y; }
        warp_sums[lane] = w;
    }
    __syncthreads();
    const uint32_t r = x + (warp ? warp_sums[warp - 1] : 0u);
    __syncthreads();
    return r;
}
__global__ void __launch_bounds__(SCAN_BLOCK) scan_block_kernel(uint32_t* __restrict__ data, uint32_t n_plus_1, uint32_t* __restrict__ block_sums, uint32_t blocks_per_row) {
    __shared__ uint32_t warp_sums[32];
    uint32_t* row = data + (size_t)blockIdx.y * n_plus_1;
    const uint32_t i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const uint32_t v = (i + 1 < n_plus_1) ? row[i] : 0u;   // slot n is not an input
    const uint32_t incl = block_inclusive_scan(v, warp_sums);
    if (i < n_plus_1) row[i] = incl - v;
    if (threadIdx.x == SCAN_BLOCK - 1) block_sums[(size_t)blockIdx.y * blocks_per_row + blockIdx.x] = incl;
}
__global__ void __launch_bounds__(SCAN_BLOCK) scan_sums_kernel(uint32_t* __restrict__ block_sums, uint32_t blocks_per_row) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    uint32_t* row = block_sums + (size_t)blockIdx.x * blocks_per_row;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < blocks_per_row; base += SCAN_BLOCK) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < blocks_per_row ? row[i] : 0u;
        const uint32_t incl = block_inclusive_scan(v, warp_sums) + carry;
        if (i < blocks_per_row) row[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == SCAN_BLOCK - 1) carry = incl;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(SCAN_BLOCK) scan_add_kernel(uint32_t* __restrict__ data, uint32_t n_plus_1, const uint32_t* __restrict__ block_sums, uint32_t blocks_per_row) {
    const uint32_t i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    if (i < n_plus_1) data[(size_t)blockIdx.y * n_plus_1 + i] += block_sums[(size_t)blockIdx.y * blocks_per_row + blockIdx.x];
}

// ------------------------------------------------------------------------------------------------ radix sort
#define RS_THREADS 256
#define RS_ITEMS_PER_THREAD 16
#define RS_TILE (RS_THREADS * RS_ITEMS_PER_THREAD)
// hist is digit-major: hist[d * n_blocks + b]
__global__ void __launch_bounds__(RS_THREADS) radix_hist_kernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t shift, uint32_t* __restrict__ hist, uint32_t n_blocks) {
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int k = 0; k < RS_ITEMS_PER_THREAD; ++k) {
        const uint32_t i = base + k * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&sh[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * n_blocks + blockIdx.x] = sh[threadIdx.x];
}
__global__ void __launch_bounds__(RS_THREADS) radix_scatter_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t n, uint32_t shift,
                                                                   const uint32_t* __restrict__ hist_scanned, uint32_t n_blocks, uint32_t* __restrict__ keys_out,
                                                                   uint32_t* __restrict__ vals_out) {
    __shared__ uint32_t warp_count[RS_THREADS / 32][256];
    __shared__ uint32_t running[256];     // items of each digit already placed by earlier sub-tiles of this block
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    running[threadIdx.x] = hist_scanned[(size_t)threadIdx.x * n_blocks + blockIdx.x];
    const uint32_t base = blockIdx.x * RS_TILE;
    for (int k = 0; k < RS_ITEMS_PER_THREAD; ++k) {
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; ++w) warp_count[w][threadIdx.x] = 0;
        __syncthreads();
        const uint32_t i = base + k * RS_THREADS + threadIdx.x;
        const bool valid = i < n;
        const uint32_t key = valid ? keys[i] : 0xffffffffu;
        const uint32_t val = valid ? vals[i] : 0u;
        const uint32_t d = valid ? ((key >> shift) & 255u) : 256u;   // invalid lanes match only each other
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) warp_count[warp][d] = __popc(peers);
        __syncthreads();
        // thread t owns digit t: exclusive prefix over warps, then advance the running base
        {
            uint32_t acc = running[threadIdx.x];
#pragma unroll
            for (int w = 0; w < RS_THREADS / 32; ++w) { const uint32_t c = warp_count[w][threadIdx.x]; warp_count[w][threadIdx.x] = acc; acc += c; }
            running[threadIdx.x] = acc;
        }
        __syncthreads();
        if (valid) {
            const uint32_t pos = warp_count[warp][d] + rank;
            keys_out[pos] = key;
            vals_out[pos] = val;
        }
        __syncthreads();
    }
}
__global__ void lower_bounds_kernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ begin, uint32_t count) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (keys[mid] < t) lo = mid + 1; else hi = mid; }
    begin[t] = lo;
}

}  // namespace

uint32_t cr_scan_scratch_words(uint32_t n_plus_1, uint32_t rows) { return rows * ((n_plus_1 + SCAN_BLOCK - 1) / SCAN_BLOCK) + 1; }
int cr_scan_exclusive(cudaStream_t stream, uint32_t* data, uint32_t n_plus_1, uint32_t rows, uint32_t* scratch) {
    if (n_plus_1 == 0 || rows == 0) return CR_OK;
    const uint32_t blocks = (n_plus_1 + SCAN_BLOCK - 1) / SCAN_BLOCK;
    scan_block_kernel<<<dim3(blocks, rows), SCAN_BLOCK, 0, stream>>>(data, n_plus_1, scratch, blocks);
    scan_sums_kernel<<<rows, SCAN_BLOCK, 0, stream>>>(scratch, blocks);
    scan_add_kernel<<<dim3(blocks, rows), SCAN_BLOCK, 0, stream>>>(data, n_plus_1, scratch, blocks);
    g_cr_kernel_launches += 3;
    CR_CUDA_TRY(cudaGetLastError());
    return CR_OK;
}
uint32_t cr_radix_scratch_words(uint32_t n) {
    const uint32_t blocks = (n + RS_TILE - 1) / RS_TILE;
    const uint32_t hist = 256 * blocks + 1;
    return hist + cr_scan_scratch_words(hist, 1);
}
int cr_radix_sort_pairs(cudaStream_t stream, uint32_t* keys, uint32_t* vals, uint32_t* keys_alt, uint32_t* vals_alt, uint32_t n, uint32_t key_bits,
                        uint32_t* scratch, uint32_t** keys_out, uint32_t** vals_out) {
    *keys_out = keys;
    *vals_out = vals;
    if (n == 0) return CR_OK;
    const uint32_t blocks = (n + RS_TILE - 1) / RS_TILE;
    const uint32_t hist_n = 256 * blocks;
    uint32_t* hist = scratch;
    uint32_t* scan_scratch = scratch + hist_n + 1;
    uint32_t *src_k = keys, *src_v = vals, *dst_k = keys_alt, *dst_v = vals_alt;
    for (uint32_t shift = 0; shift < key_bits; shift += 8) {
        radix_hist_kernel<<<blocks, RS_THREADS, 0, stream>>>(src_k, n, shift, hist, blocks);
        g_cr_kernel_launches += 1;
        const int st = cr_scan_exclusive(stream, hist, hist_n + 1, 1, scan_scratch);
        if (st != CR_OK) return st;
        radix_scatter_kernel<<<blocks, RS_THREADS, 0, stream>>>(src_k, src_v, n, shift, hist, blocks, dst_k, dst_v);
        g_cr_kernel_launches += 1;
        uint32_t* t = src_k; src_k = dst_k; dst_k = t;
        t = src_v; src_v = dst_v; dst_v = t;
    }
    CR_CUDA_TRY(cudaGetLastError());
    *keys_out = src_k;
    *vals_out = src_v;
    return CR_OK;
}
int cr_lower_bounds(cudaStream_t stream, const uint32_t* sorted_keys, uint32_t n, uint32_t* begin, uint32_t n_keys_plus_1) {
    lower_bounds_kernel<<<(n_keys_plus_1 + 255) / 256, 256, 0, stream>>>(sorted_keys, n, begin, n_keys_plus_1);
    g_cr_kernel_launches += 1;
    CR_CUDA_TRY(cudaGetLastError());
    return CR_OK;
}
