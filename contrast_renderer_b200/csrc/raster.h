// raster.h — host-visible interface of raster.cu (binner + K3 tile rasteriser).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/contrast_b200.h"

#define CR_TILE 16   // pixels per tile edge (1 sample per pixel) — one CTA of 256 threads owns one tile

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

// One recorded Shape::render call (src/renderer.rs:267) with the pass state current at record time.
struct DeviceCommand {
    uint32_t batch;
    uint32_t shape;
    uint32_t instance_begin, instance_end;
    uint32_t operation;           // cr_render_operation
    uint32_t ref;                 // stencil reference = clip_depth << winding_counter_bits (src/renderer.rs:936)
    uint32_t save_layer, restore_layer;
};

struct RasterTarget {
    float4* color;                // [height][width] premultiplied RGBA32F
    uint8_t* stencil;             // [height][width]
    float* alpha_layers;          // [layer][height][width]
    uint32_t width, height, tiles_x, tiles_y;
    uint32_t wmask, cmask;        // winding_counter_mask / clip_nesting_counter_mask (src/renderer.rs:565-566)
    uint32_t blending, cull_mode;
};

struct RasterScene {
    const DeviceBatch* batches;
    const DeviceCommand* commands;
    const uint32_t* cmd_cand_begin;   // [n_commands + 1] exclusive scan of candidate primitives per command
    uint32_t n_commands;
    const float* transforms;          // [n_instances][16]
    const float* colors;              // [n_instances][4] or null
};

int cr_raster_count_candidates(cudaStream_t stream, const DeviceBatch* batches, const DeviceCommand* commands, uint32_t n_commands, uint32_t* cmd_cands);
int cr_raster_bin_count(cudaStream_t stream, const RasterScene& scene, const RasterTarget& target, uint32_t n_candidates, uint32_t* cand_tiles);
int cr_raster_bin_emit(cudaStream_t stream, const RasterScene& scene, const RasterTarget& target, uint32_t n_candidates, const uint32_t* cand_pair_begin,
                       uint32_t* pair_tile, uint32_t* pair_cand);
int cr_raster_tiles(cudaStream_t stream, const RasterScene& scene, const RasterTarget& target, const uint32_t* tile_begin, const uint32_t* pair_cand,
                    unsigned long long* covered_samples);
