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

// One recorded Shape::render call (src/renderer.rs:267) with the pass state current at record time, as the host records it:
// 32 bytes, no table look-ups (the slice tables of a batch live on the device and may still be in flight).
struct CompactCommand {
    uint32_t batch, shape;
    uint32_t instance_begin, instance_count;
    uint32_t operation;           // cr_render_operation
    uint32_t ref;                 // stencil reference = clip_depth << winding_counter_bits (src/renderer.rs:936)
    uint32_t layers;              // save_layer | restore_layer << 16
    uint32_t _pad;
};
static_assert(sizeof(CompactCommand) == 32, "CompactCommand layout");

// Device-side sizes of one submitted pass. The host sizes grids and buffers from CAPACITIES (last frame's sizes plus slack) and
// only reads these words back after the pass has been enqueued; when a capacity does not suffice the kernels that would overrun
// it do nothing, the tile kernel does not run (the attachments stay untouched) and the host re-submits with larger buffers.
#define CR_PASS_OVERFLOW_CANDS 1u
#define CR_PASS_OVERFLOW_PAIRS 2u
#define CR_PASS_OVERFLOW_CLIP 4u
struct PassCounters {
    unsigned long long cand_total;    // candidates of the pass, 64-bit (the numbering itself is 32-bit)
    unsigned long long pair_total;    // (tile, candidate) pairs, 64-bit
    unsigned long long covered;       // samples that passed the stencil test of a colour cover
    uint32_t n_pairs_live;            // pair_total if it fits the pair capacity and 32 bits, else 0 (published by the bin-emit kernel)
    uint32_t flags;                   // CR_PASS_OVERFLOW_*
    uint32_t clip_total;              // triangles produced by frustum clipping (they live behind the candidates in the record array)
    uint32_t _pad;
};
static_assert(sizeof(PassCounters) == 40, "PassCounters is read back into 10 pinned words");

// Frustum clipping (WebGPU clips primitives against the view volume; the reference's demo places instances in perspective,
// examples/showcase/main.rs:163-201). A triangle with a vertex on or behind the eye plane (w <= 0) or further than 2^21 px from
// the origin cannot be snapped; it is clipped in clip space (Sutherland-Hodgman, planes in this order: w >= CR_CLIP_W_MIN,
// G w - x >= 0, G w + x >= 0, G w - y >= 0, G w + y >= 0 with the guard band G = cr_guard_band(width, height); an edge's
// intersection point is always interpolated from its INSIDE end, so that two triangles sharing the edge get the same point),
// the polygon is cut into a fan and every fan triangle takes the place of the original in the draw order. Depth (z) is not
// clipped (as with WebGPU's unclipped-depth control). Attributes, w and z of the new vertices are interpolated linearly in
// clip space; flat attributes stay those of the original first vertex.
#define CR_CLIP_W_MIN 1.0e-6f
#define CR_CLIP_MAX_TRIANGLES 6   // a triangle cut by five planes has at most 8 corners
__host__ __device__ inline float cr_guard_band(uint32_t width, uint32_t height) {
    const uint32_t m = width > height ? width : height;
    const uint32_t g = 1048576u / (m ? m : 1u);
    return (float)(g ? g : 1u);
}
// Per clipped triangle: what its three corners carry instead of vertex-array attributes (in fan order, before orientation).
struct ClipAttr {
    float invw[3];
    float attr[3][4];   // interpolated attributes (strokes: [1][3] = flat float, see TilePrim); colour covers with a depth test: [i][0] = z / w
    uint32_t flat_u;
};
static_assert(sizeof(ClipAttr) == 64, "ClipAttr layout");

// The expanded form of a command (built on the device by cr_raster_expand from the batch's slice tables), so that the
// vertex stage needs no further table walks. A "candidate" is one
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
    // One render target composed from DRAW-ORDER slices (batch sharding into one target): rank r tessellates and bins
    // only its contiguous slice of the draw order. Per tile, the ranks whose slices touch it form a chain in rank order; a rank
    // waits for its predecessor's tile state (colour + stencil, stored into ITS attachments over NVLink), continues rasterising
    // on top of it — the same operations in the same order as one GPU would execute — and hands the tile to its successor; the
    // last rank of the chain stores the finished tile into every rank's attachments. Which ranks touch which tile is exchanged
    // as bitmaps through the peer-mapped `exchange` buffers; all flags carry the pass's epoch, so nothing is ever cleared.
    uint32_t order_world, order_rank;             // world <= 1: off
    uint32_t order_epoch;
    uint32_t order_mask_words;                    // words per rank bitmap = ceil(n_tiles / 32)
    uint32_t* exchange;                           // own: [CR_MAX_PEERS + 1 ready words][n_tiles tile flags][(CR_MAX_PEERS + 1) x mask_words bitmaps]
    uint32_t* peer_exchange[CR_MAX_PEERS];        // the other ranks' (peer-mapped), slot = rank - (rank > order_rank)
};
__host__ __device__ inline uint32_t cr_exchange_flag_offset() { return CR_MAX_PEERS + 1; }
__host__ __device__ inline uint32_t cr_exchange_mask_offset(uint32_t n_tiles) { return CR_MAX_PEERS + 1 + n_tiles; }
__host__ __device__ inline size_t cr_exchange_words(uint32_t n_tiles) { return (size_t)CR_MAX_PEERS + 1 + n_tiles + (size_t)(CR_MAX_PEERS + 1) * ((n_tiles + 31) / 32); }
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
    const ClipAttr* clip_attrs;       // [clip capacity] corners of the triangles frustum clipping produced (PrimRecord::v[0] indexes it)
};

// Commands -> DeviceCommands + the exclusive scan of their candidate counts (cmd_cand_begin, n_commands + 1 words; one launch
// for up to CR_EXPAND_FUSED_MAX commands, else expand + cr_scan_exclusive with `scan_scratch`). Zeroes *counters first.
#define CR_EXPAND_FUSED_MAX 1024u
int cr_raster_expand(cudaStream_t stream, const CompactCommand* compact, uint32_t n_commands, const DeviceBatch* batches, DeviceCommand* commands,
                     uint32_t* cmd_cand_begin, PassCounters* counters, uint32_t* scan_scratch);
// Vertex stage + tile counting over the candidate CAPACITY: fills records and cand_tiles[0..cand_capacity) (zero beyond the
// live count). big_list: cand_capacity + 1 words of scratch (candidates whose tile box is large are listed there and binned
// one warp each). counters->pair_total receives the 64-bit number of (tile, candidate) pairs: the placing scan is 32-bit.
// clip_list: cand_capacity + 1 words of scratch (candidates that need frustum clipping); their fan triangles go to
// records[cand_capacity ...) / clip_attrs[0 ... clip_capacity).
int cr_raster_setup(cudaStream_t stream, const RasterScene& scene, const RasterTarget& target, uint32_t cand_capacity, PrimRecord* records, uint32_t* cand_tiles,
                    uint32_t* big_list, uint32_t* clip_list, ClipAttr* clip_attrs, uint32_t clip_capacity, PassCounters* counters);
// (tile, candidate) pairs of every valid record, at cand_pair_begin[candidate]; publishes counters->n_pairs_live / flags.
int cr_raster_bin_emit(cudaStream_t stream, const RasterTarget& target, uint32_t cand_capacity, uint32_t pair_capacity, const PrimRecord* records,
                       const uint32_t* cand_pair_begin, const uint32_t* big_list, const uint32_t* clip_list, uint32_t* pair_tile, uint32_t* pair_cand,
                       PassCounters* counters);
// Draw-order sharding: publishes which tiles this rank's slice touches (bitmap from tile_begin) to every rank, then the ready flag.
int cr_raster_publish_touched_tiles(cudaStream_t stream, const RasterTarget& target, const uint32_t* tile_begin);
// counters may be null (clear-only launch on an empty tile table).
// Tile-ordered primitive stream: one cr_tile_prim_bytes()-byte record per sorted (tile, candidate) pair, set up for its tile.
size_t cr_tile_prim_bytes();
int cr_raster_tile_prims(cudaStream_t stream, const RasterScene& scene, const RasterTarget& target, const PrimRecord* records, const uint32_t* pair_tile, const uint32_t* pair_cand,
                         uint32_t pair_capacity, void* tile_prims, const PassCounters* counters);
int cr_raster_tiles(cudaStream_t stream, const RasterScene& scene, const RasterTarget& target, const void* tile_prims, const uint32_t* tile_begin, PassCounters* counters);
