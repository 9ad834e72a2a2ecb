// Tensor maps (TMA descriptors) that make cp.async.bulk.tensor write the shared-memory layout the tcgen05
// descriptors of the attention kernels expect: SWIZZLE_NONE "interleaved" 8x8 core matrices,
//   offset(row, chunk) = (row / 8) * (chunks_per_row * 128) + chunk * 128 + (row % 8) * 16.
// A row-major [rows, cols] 16-bit tile is described to the TMA unit as a 5-D tensor
//   d0 = col % 8 (one 16-byte chunk)   d1 = row % 8   d2 = col / 8 (+ head offset)   d3 = row / 8   d4 = batch
// and fetched with the box {8, 8, chunks, row_groups, 1}: the box is written to shared memory with d0 fastest,
// which is exactly the core-matrix order - one instruction from one thread replaces 1024-2048 LDGSTS.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ub {

// [B, L, H, 64] 16-bit tensor addressed through element strides (sb, sl) with heads contiguous (stride 64):
// box = one 128-row x 64-column tile of one (batch, head); coordinates {0, 0, head * 8, row0 / 8, batch}.
bool make_head_tile_map(CUtensorMap* map, const void* base, bool bf16, int B, int L, int H, long long sb, long long sl,
                        int box_rows);
// bias [NB, Lq, Lk] (NB = bias_batch * H) 16-bit contiguous: box = 128 x 128 tile; coordinates
// {0, 0, key0 / 8, q0 / 8, nb}.
bool make_bias_tile_map(CUtensorMap* map, const void* base, bool bf16, int NB, int Lq, int Lk);

// ---- 128-byte-swizzled variants (CU_TENSOR_MAP_SWIZZLE_128B) ------------------------------------------------------
// The box rows are whole 128-byte lines (64 16-bit elements): one TMA request per row instead of one per 16-byte
// chunk (measured: the 16-byte-row boxes above cost ~1.7 cycles per chunk, 3500 cycles for a K + V tile pair, which
// bounded the attention kernels).  Shared-memory layout: row r at r * 128, its 16-byte chunk c at ((c ^ (r & 7)) * 16)
// - the canonical UMMA SWIZZLE_128B layout (K-major operands, and MN-major operands whose MN extent is 64).
// [B, L, H, 64] tensor, box = box_rows x 64 of one (batch, head); coordinates {0, head, row0, batch}.
bool make_head_tile_map_sw128(CUtensorMap* map, const void* base, bool bf16, int B, int L, int H, long long sb,
                              long long sl, long long sh, int box_rows);
// bias [NB, Lq, Lk]: box = 128 rows x 64 columns (a 128 x 128 tile is two boxes); coordinates {key0, q0, nb}.
bool make_bias_tile_map_sw128(CUtensorMap* map, const void* base, bool bf16, int NB, int Lq, int Lk);

}  // namespace ub
