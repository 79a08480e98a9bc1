// raster.h — host-visible interface of raster.cu (primitive setup, binner, K3 tile rasteriser).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/contrast_b200.h"

#define CR_MAX_PEERS 7   // tile sharding: up to 8 GPUs of one node
#define CR_TILE 16   // pixels per tile edge — one CTA of 256 threads owns one tile, one thread per pixel (all of its samples)

// One tessellated cr_shape_batch as the rasteriser sees it.
struct DeviceBatch {
    const void* vtx[7];           // CAT_LINE .. CAT_RC
    const float2* hull;           // hull strips, slice of shape s starts at cat_begin[CNT_PROTO][s]
    const uint32_t* idx[3];       // line, joint, solid
    const uint32_t* cat_begin;    // [CNT_COUNT][n_shapes + 1]
    const uint32_t* hull_count;   // [n_shapes]
    const void* stroke;           // DynamicStrokeDescriptor (48 B) x n_groups
    uint32_t n_shapes;
    uint32_t n_groups;
};

// One recorded Shape::render call (src/renderer.rs:267) with the pass state current at record time, expanded on the
// host (the slice tables are mirrored there) so that device code needs no further table walks. A "candidate" is one
// index slot of a strip, one triangle of a list, or one triangle of the hull strip, times the instance count; the
// candidates of a command are numbered category by category in the draw order of src/renderer.rs:275-354.
struct DeviceCommand {
    uint32_t cat_end[8];          // cumulative candidate end of each vertex category, relative to the command's first candidate
    uint32_t slots[8];            // candidates per instance in each category (0 if the category is not drawn)
    uint32_t vbase[8];            // first vertex of the shape's slice in the batch-wide category array ([7]: hull slice)
    uint32_t ibase[3];            // first index slot of the shape's slice (line, joint, solid)
    uint32_t batch;
    uint32_t instance_begin, instance_count;
    uint32_t operation;           // cr_render_operation
    uint32_t ref;                 // stencil reference = clip_depth << winding_counter_bits (src/renderer.rs:936)
    uint32_t layers;              // save_layer | restore_layer << 16
    uint32_t _pad[3];
};
static_assert(sizeof(DeviceCommand) == 144, "DeviceCommand layout");

// One candidate after the vertex stage: transformed, snapped to 1/256 px, oriented clockwise on screen. 64 bytes; the
// binner only reads the first 32.
struct PrimRecord {
    int X[3], Y[3];
    uint32_t meta;                // pipe (bits 0-3) | front << 4 | swapped << 5 | valid << 6 | big << 7 | category << 8
    uint32_t cmd;
    uint32_t instance;
    uint32_t v[3];                // vertex numbers in the batch-wide category array, in submission order
    uint32_t ref;                 // the command's stencil reference, alpha layers and batch (so K3 never reads the command)
    uint32_t layers;
    uint32_t batch;
    uint32_t _pad;
};
static_assert(sizeof(PrimRecord) == 64, "PrimRecord layout");

struct RasterTarget {
    void* color;                  // [height][width][samples] premultiplied: float4 (RGBA32F) or one packed unorm8 texel (u32) per sample
    uint8_t* stencil;             // [height][width][samples]
    void* alpha_layers;           // [layer][height][width][samples]: f32 (RGBA32F targets) or R8Unorm (8-bit targets, src/renderer.rs:783,898)
    float* depth;                 // [height][width][samples] f32, or null when the configuration has no depth test / write
    uint32_t depth_compare;       // cr_compare_function of the colour cover, 0 / 8 = always
    uint32_t depth_write;
    uint32_t clear_depth;         // the pass clears the depth aspect to depth_clear_value
    float depth_clear_value;
    uint32_t color_format;        // cr_color_format
    uint32_t width, height, tiles_x, tiles_y;
    uint32_t samples;             // 1 (pixel centre) or 4 (WebGPU standard pattern (6,2),(14,6),(2,10),(10,14) / 16)
    int sample_lo, sample_hi;     // smallest / largest sample offset inside a pixel in 1/256 px: 128,128 or 32,224
    uint32_t wmask, cmask;        // winding_counter_mask / clip_nesting_counter_mask (src/renderer.rs:565-566)
    uint32_t blending, cull_mode;
    uint32_t pixel_run_max;       // runs of at most this many primitives execute in K3's pixel mode
    uint32_t clear_color, clear_stencil;   // the pass clears the attachment: K3 starts from zero instead of loading, and writes EVERY tile
    // One render target spanning several GPUs (tile sharding): tile (tx, ty) is owned by rank (tx + ty) % world.
    // A rank bins and rasterises only the tiles it owns and K3 stores every finished tile into its own attachments AND,
    // over NVLink, into the peer-mapped attachments of the other ranks, so each rank ends up with the complete frame.
    uint32_t shard_world, shard_rank;             // world <= 1: not sharded
    void* peer_color[CR_MAX_PEERS];               // [world - 1] attachments of the other ranks (peer-mapped), or null
    uint8_t* peer_stencil[CR_MAX_PEERS];
};
__host__ __device__ inline bool cr_tile_owned(const RasterTarget& tg, int tx, int ty) {
    return tg.shard_world <= 1u || (uint32_t)(tx + ty) % tg.shard_world == tg.shard_rank;
}

struct RasterScene {
    const DeviceBatch* batches;
    const DeviceCommand* commands;
    const uint32_t* cmd_cand_begin;   // [n_commands + 1] first candidate of each command
    uint32_t n_commands;
    const float* transforms;          // [n_instances][16]
    const float* colors;              // [n_instances][4] or null
};

// Vertex stage + tile counting: fills records[0..n) and cand_tiles[0..n). big_list: n + 1 words of scratch (candidates
// whose tile box is large are listed there and binned one warp each).
// *pair_total (device, zeroed here) receives the 64-bit number of (tile, candidate) pairs: the placing scan is 32-bit.
int cr_raster_setup(cudaStream_t stream, const RasterScene& scene, const RasterTarget& target, uint32_t n_candidates, PrimRecord* records, uint32_t* cand_tiles,
                    uint32_t* big_list, unsigned long long* pair_total);
// (tile, candidate) pairs of every valid record, at cand_pair_begin[candidate].
int cr_raster_bin_emit(cudaStream_t stream, const RasterTarget& target, uint32_t n_candidates, const PrimRecord* records, const uint32_t* cand_pair_begin,
                       const uint32_t* big_list, uint32_t* pair_tile, uint32_t* pair_cand);
int cr_raster_tiles(cudaStream_t stream, const RasterScene& scene, const RasterTarget& target, const PrimRecord* records, const uint32_t* tile_begin,
                    const uint32_t* pair_cand, unsigned long long* covered_samples);
