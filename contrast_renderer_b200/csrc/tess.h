// tess.h — host-visible interface of tess.cu (K1/K2 tessellation, scans, hull).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/contrast_b200.h"

// cr_path_soa with every pointer in device memory.
struct DevicePaths {
    uint32_t n_paths;
    uint32_t n_segments;
    const float* start;
    const uint32_t* segment_begin;
    const uint8_t* segment_types;
    const uint32_t* type_begin;        // [5][n_paths + 1]
    const float* seg[5];               // line, integral quadratic, integral cubic, rational quadratic, rational cubic
    const cr_stroke_options* stroke_options;
};

// Destination arrays of the emit pass (whole batch).
struct TessOutput {
    void* vtx[7];        // CAT_LINE .. CAT_RC, packed exactly like src/vertex.rs
    float2* proto;       // proto_hull points
    uint32_t* idx[3];    // line, joint, solid: (shape-relative vertex index << 1) | strip parity, CR_RESTART between strips
};

// Element capacities of the output arrays of an earlier build, indexed like the counters (device_common.cuh CNT_*). An
// optimistic rebuild enqueues the emit pass before the host knows the new sizes; shape_bounds_kernel compares them with these
// and raises CR_DEVERR_CAPACITY (emit and hull then do nothing) when one does not suffice.
struct TessCapacity { uint32_t v[11]; };

// has_cubics: the batch holds integral or rational cubic segments (false selects the kernels without the cubic fill builder)
int cr_tess_count(cudaStream_t stream, const DevicePaths& paths, uint32_t n_groups, uint32_t* counts, uint32_t* err_flag, bool has_cubics = true);
// Also stores max over shapes of the proto-hull point count into *max_proto (device word, zeroed by the caller).
int cr_tess_shape_bounds(cudaStream_t stream, const uint32_t* offsets, uint32_t n_paths, const uint32_t* shape_path_begin, uint32_t n_shapes,
                         uint32_t* cat_begin, uint32_t* max_proto, const TessCapacity& caps, uint32_t* err_flag);
int cr_tess_emit(cudaStream_t stream, const DevicePaths& paths, const uint32_t* offsets, const uint32_t* shape_path_begin, uint32_t n_shapes,
                 const TessOutput& out, uint32_t* err_flag, bool has_cubics = true);
int cr_tess_hull(cudaStream_t stream, float2* proto, float2* scratch_a, float2* scratch_b, const uint32_t* proto_begin, uint32_t n_shapes,
                 float2* hull_out, uint32_t* hull_count, uint32_t max_points, const uint32_t* err_flag, cudaEvent_t after_sort = nullptr);   // after_sort: optional event recorded between the sort and the chain kernel
