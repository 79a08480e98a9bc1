// prims.cu — exclusive scan, stable LSD radix sort and sorted-key range search, hand-written for sm_100a.
// These are the "sizes and order" plumbing between the count and emit passes of the tessellator and between the
// binner and the tile rasteriser (the stable sort is what carries the reference's draw order into every tile).
//
// Every primitive is ONE launch and takes its element count either from the host or from a device word, so that the
// render pass can be enqueued without the host ever learning the sizes (api.cu, "optimistic submit").
#include <algorithm>
#include <atomic>
#include "device_common.cuh"
#include "prims.h"

unsigned long long g_cr_kernel_launches = 0;

namespace {

// ------------------------------------------------------------------------------------------------------ scan
// Single-pass exclusive scan with decoupled look-back. A block takes a ticket (so that lower-numbered tiles are always
// running or done), scans its tile, publishes its aggregate, and sums its predecessors' published values 32 at a time until
// it meets one that already carries an inclusive prefix. Status word: epoch (30 bits) | state (2 bits) | value (32 bits); the
// epoch changes with every call, so the scratch array is never cleared.
#define SCAN_THREADS 512
#define SCAN_ITEMS 16
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)
#define SCAN_AGGREGATE 1ull
#define SCAN_INCLUSIVE 2ull

__device__ __forceinline__ unsigned long long scan_pack(uint32_t epoch, unsigned long long state, uint32_t value) {
    return ((unsigned long long)epoch << 34) | (state << 32) | value;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// scratch: [SCAN_MAX_ROWS ticket counters (u64)] [rows x blocks status words]. The tickets sit in front so that calls with
// different shapes, which share one scratch buffer, can never leave a status word where another call keeps a ticket.
#define SCAN_MAX_ROWS 16
__global__ void __launch_bounds__(SCAN_THREADS) scan_lookback_kernel(uint32_t* __restrict__ data, uint32_t n_plus_1, unsigned long long* __restrict__ scratch,
                                                                     uint32_t blocks_per_row, uint32_t epoch) {
    __shared__ uint32_t sh_warp[SCAN_THREADS / 32];
    __shared__ uint32_t sh_ticket, sh_prefix;
    uint32_t* row = data + (size_t)blockIdx.y * n_plus_1;
    unsigned long long* state = scratch + blockIdx.y;
    unsigned long long* status = scratch + SCAN_MAX_ROWS + (size_t)blockIdx.y * blocks_per_row;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        const uint32_t t = (uint32_t)atomicAdd(state, 1ull);
        if (t + 1 == blocks_per_row) st_relaxed_u64(state, 0ull);   // every ticket of this call is taken: ready for the next call
        sh_ticket = t;
    }
    __syncthreads();
    const uint32_t tile = sh_ticket;
    const uint32_t base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t sum = 0;
    const bool aligned = (reinterpret_cast<uintptr_t>(row) & 15u) == 0;   // a thread's 16 items are 64 bytes: four 128-bit accesses when the row allows
    if (aligned && base + SCAN_ITEMS < n_plus_1) {
        const uint4* src = reinterpret_cast<const uint4*>(row + base);
#pragma unroll
        for (int q = 0; q < SCAN_ITEMS / 4; ++q) {
            const uint4 t = src[q];
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) sum += v[k];
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            const uint32_t i = base + k;
            v[k] = (i + 1 < n_plus_1) ? row[i] : 0u;   // slot n is not an input
            sum += v[k];
        }
    }
    // block exclusive scan of the per-thread sums
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += y; }
    if (lane == 31) sh_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < SCAN_THREADS / 32 ? sh_warp[lane] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= (uint32_t)o) w += y; }
        if (lane < SCAN_THREADS / 32) sh_warp[lane] = w;
    }
    __syncthreads();
    const uint32_t block_total = sh_warp[SCAN_THREADS / 32 - 1];
    const uint32_t thread_excl = (incl - sum) + (warp ? sh_warp[warp - 1] : 0u);
    // publish, look back
    if (warp == 0) {
        if (tile == 0) {
            if (lane == 0) { st_relaxed_u64(&status[0], scan_pack(epoch, SCAN_INCLUSIVE, block_total)); sh_prefix = 0; }
        } else {
            if (lane == 0) st_relaxed_u64(&status[tile], scan_pack(epoch, SCAN_AGGREGATE, block_total));
            uint32_t prefix = 0;
            int look = (int)tile - 1;
            for (;;) {
                const int idx = look - (int)lane;
                unsigned long long w = scan_pack(epoch, SCAN_INCLUSIVE, 0u);   // below tile 0: an inclusive prefix of zero
                if (idx >= 0) {
                    do { w = ld_relaxed_u64(&status[idx]); } while ((uint32_t)(w >> 34) != (epoch & 0x3fffffffu) || ((w >> 32) & 3ull) == 0ull);
                }
                const uint32_t inclusive = __ballot_sync(0xffffffffu, ((w >> 32) & 3ull) == SCAN_INCLUSIVE);
                const int first = __ffs((int)inclusive) - 1;   // nearest predecessor that carries an inclusive prefix (-1: none in this window)
                const uint32_t take = (first < 0 || (int)lane <= first) ? (uint32_t)w : 0u;
                prefix += __reduce_add_sync(0xffffffffu, take);
                if (first >= 0) break;
                look -= 32;
            }
            if (lane == 0) { st_relaxed_u64(&status[tile], scan_pack(epoch, SCAN_INCLUSIVE, prefix + block_total)); sh_prefix = prefix; }
        }
    }
    __syncthreads();
    uint32_t running = sh_prefix + thread_excl;
    if (aligned && base + SCAN_ITEMS <= n_plus_1) {
        uint4* dst = reinterpret_cast<uint4*>(row + base);
#pragma unroll
        for (int q = 0; q < SCAN_ITEMS / 4; ++q) {
            uint4 t;
            t.x = running; running += v[4 * q];
            t.y = running; running += v[4 * q + 1];
            t.z = running; running += v[4 * q + 2];
            t.w = running; running += v[4 * q + 3];
            dst[q] = t;
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            const uint32_t i = base + k;
            if (i < n_plus_1) row[i] = running;
            running += v[k];
        }
    }
}

// ------------------------------------------------------------------------------------------------ radix sort
#define RS_THREADS 256
#define RS_ITEMS_PER_THREAD 16
#define RS_TILE (RS_THREADS * RS_ITEMS_PER_THREAD)
// hist is digit-major: hist[d * n_blocks + b]. Blocks beyond the live count write zero histograms.
__global__ void __launch_bounds__(RS_THREADS) radix_hist_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ n_ptr, uint32_t shift,
                                                                uint32_t* __restrict__ hist, uint32_t n_blocks) {
    __shared__ uint32_t sh[256];
    const uint32_t n = *n_ptr;
    sh[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
    if (base < n) {
#pragma unroll 4
        for (int k = 0; k < RS_ITEMS_PER_THREAD; ++k) {
            const uint32_t i = base + k * RS_THREADS + threadIdx.x;
            if (i < n) atomicAdd(&sh[(keys[i] >> shift) & 255u], 1u);
        }
        __syncthreads();
    }
    hist[(size_t)threadIdx.x * n_blocks + blockIdx.x] = sh[threadIdx.x];
}
__global__ void __launch_bounds__(RS_THREADS) radix_scatter_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ n_ptr,
                                                                   uint32_t shift, const uint32_t* __restrict__ hist_scanned, uint32_t n_blocks,
                                                                   uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    __shared__ uint32_t warp_count[RS_THREADS / 32][256];
    __shared__ uint32_t running[256];     // items of each digit already placed by earlier sub-tiles of this block
    const uint32_t n = *n_ptr;
    const uint32_t base = blockIdx.x * RS_TILE;
    if (base >= n) return;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    running[threadIdx.x] = hist_scanned[(size_t)threadIdx.x * n_blocks + blockIdx.x];
    for (int k = 0; k < RS_ITEMS_PER_THREAD; ++k) {
        if (base + k * RS_THREADS >= n) break;   // block-uniform
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
// begin[t] = first i with keys[i] >= t: element i owns the keys in (keys[i - 1], keys[i]] (all keys up to keys[0] for i = 0) and
// writes i there; the thread behind the last element owns the rest. One coalesced pass instead of a binary search per key.
__global__ void __launch_bounds__(256) key_ranges_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ n_ptr, uint32_t* __restrict__ begin, uint32_t count) {
    const uint32_t n = *n_ptr;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += stride) {
        const uint32_t lo = i == 0 ? 0u : min(keys[i - 1] + 1u, count);
        const uint32_t hi = i == n ? count : min(keys[i] + 1u, count);
        for (uint32_t t = lo; t < hi; ++t) begin[t] = i;
    }
}

}  // namespace

uint32_t cr_scan_scratch_words(uint32_t n_plus_1, uint32_t rows) { return 2u * (SCAN_MAX_ROWS + rows * ((n_plus_1 + SCAN_TILE - 1) / SCAN_TILE)) + 2u; }
int cr_scan_exclusive(cudaStream_t stream, uint32_t* data, uint32_t n_plus_1, uint32_t rows, uint32_t* scratch) {
    if (n_plus_1 == 0 || rows == 0) return CR_OK;
    if (rows > SCAN_MAX_ROWS) { cr_set_error_message("cr_scan_exclusive: %u rows > %d", rows, SCAN_MAX_ROWS); return CR_ERR_INVALID_ARGUMENT; }
    static std::atomic<uint32_t> counter{0};
    uint32_t epoch = (counter.fetch_add(1u) + 1u) & 0x3fffffffu;   // unique per call; scratch that cr_scan_prepare zeroed is in epoch 0, which is never live
    if (epoch == 0) epoch = (counter.fetch_add(1u) + 1u) & 0x3fffffffu;
    const uint32_t blocks = (n_plus_1 + SCAN_TILE - 1) / SCAN_TILE;
    unsigned long long* s64 = reinterpret_cast<unsigned long long*>((reinterpret_cast<uintptr_t>(scratch) + 7u) & ~(uintptr_t)7u);
    scan_lookback_kernel<<<dim3(blocks, rows), SCAN_THREADS, 0, stream>>>(data, n_plus_1, s64, blocks, epoch);
    g_cr_kernel_launches += 1;
    CR_CUDA_TRY(cudaGetLastError());
    return CR_OK;
}
int cr_scan_prepare(cudaStream_t stream, uint32_t* scratch, size_t words) {
    CR_CUDA_TRY(cudaMemsetAsync(scratch, 0, words * 4, stream));   // tickets at zero, status words in epoch 0 (never live)
    return CR_OK;
}
uint32_t cr_radix_scratch_words(uint32_t n) {
    const uint32_t blocks = (n + RS_TILE - 1) / RS_TILE;
    const uint32_t hist = 256 * blocks + 1;
    return hist + 2u + cr_scan_scratch_words(hist, 1);
}
int cr_radix_sort_pairs(cudaStream_t stream, uint32_t* keys, uint32_t* vals, uint32_t* keys_alt, uint32_t* vals_alt, uint32_t capacity, const uint32_t* n_ptr,
                        uint32_t key_bits, uint32_t* scratch, uint32_t** keys_out, uint32_t** vals_out) {
    *keys_out = keys;
    *vals_out = vals;
    if (capacity == 0) return CR_OK;
    const uint32_t blocks = (capacity + RS_TILE - 1) / RS_TILE;
    const uint32_t hist_n = 256 * blocks;
    uint32_t* hist = scratch;
    uint32_t* scan_scratch = scratch + hist_n + 2;
    uint32_t *src_k = keys, *src_v = vals, *dst_k = keys_alt, *dst_v = vals_alt;
    for (uint32_t shift = 0; shift < key_bits; shift += 8) {
        radix_hist_kernel<<<blocks, RS_THREADS, 0, stream>>>(src_k, n_ptr, shift, hist, blocks);
        g_cr_kernel_launches += 1;
        const int st = cr_scan_exclusive(stream, hist, hist_n + 1, 1, scan_scratch);
        if (st != CR_OK) return st;
        radix_scatter_kernel<<<blocks, RS_THREADS, 0, stream>>>(src_k, src_v, n_ptr, shift, hist, blocks, dst_k, dst_v);
        g_cr_kernel_launches += 1;
        uint32_t* t = src_k; src_k = dst_k; dst_k = t;
        t = src_v; src_v = dst_v; dst_v = t;
    }
    CR_CUDA_TRY(cudaGetLastError());
    *keys_out = src_k;
    *vals_out = src_v;
    return CR_OK;
}
int cr_lower_bounds(cudaStream_t stream, const uint32_t* sorted_keys, uint32_t capacity, const uint32_t* n_ptr, uint32_t* begin, uint32_t n_keys_plus_1) {
    const uint32_t blocks = std::min<uint32_t>(148u * 8u, capacity / 256u + 1u);
    key_ranges_kernel<<<blocks, 256, 0, stream>>>(sorted_keys, n_ptr, begin, n_keys_plus_1);
    g_cr_kernel_launches += 1;
    CR_CUDA_TRY(cudaGetLastError());
    return CR_OK;
}
