// Pieces of the x4 CEM fast paths shared by cem.cu (two-launch streaming kernels) and cem_fused.cu (single launch).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/esr_b200.h"

namespace esr {

struct CemTab {            // polyphase tap tables, [phase][cell offset -2..2]
    float down_h[4][5];    // down: weight of element e of cell j+c for output column j
    float down_v[4][5];    // down: weight of HR row 4I+q for output row I-m, index [q][m+2]
    float up[4][5];        // up: weight of cell j+c for HR phase phi (same table for rows)
    // packed-fp32 (FFMA2) operand forms of the same numbers
    float2 down_v2[4][5];  // (down_v, down_v)
    float2 up_v2[4][5];    // (up, up)
    float2 up_h01[5];      // (up[0][c], up[1][c])
    float2 up_h23[5];      // (up[2][c], up[3][c])
};

CemTab make_tab(const esr_cem_filters& f);
// fp32 [planes][rows][cols] tensor map with a box of box_cols x box_rows x 1
int make_plane_map(CUtensorMap* tm, const float* base, int planes, int rows, int cols, int box_cols, int box_rows);
int num_sms_cached();
// single-launch x4 projection (cem_fused.cu): ESR_OK when it ran, 1 when the shape is not its business, < 0 on errors
bool cem_fused_enabled();
int cem_project4f(const esr_cem_filters& f, const float* y, const float* x, int planes, int H, int W, int crop, float* out,
                  float* workspace, cudaStream_t s);

}  // namespace esr
