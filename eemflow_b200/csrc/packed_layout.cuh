// Layout of the fp16 "working pyramid" (EEM_CORR_TF32_F16): what the correlation GEMM writes and the window lookup
// reads when the caller does not need the reference's f32 `corr_pyramid` tensors (model/corr.py:13-27) themselves.
//
//   packed[b*P + i][row]  fp16, one ROW per source position i holding ALL levels back to back:
//     level l starts at element off[l]; its h_l x w_l map is stored as 4x4-pixel tiles, tile-row-major
//     (ty * tiles_x + tx), row-major inside a tile: element ((ty*tiles_x + tx)*16 + (y%4)*4 + (x%4)).
//   A tile is 16 fp16 = 32 bytes = one L2 sector, so the (2r+2)^2 = 10x10 tap window of a lookup touches whole
//   sectors only (3.25 x 3.25 tiles on average = 338 B for 200 B of taps; the f32 row-major volume fetches
//   16 sectors = 512 B for 400 B).  Map cells beyond h_l / w_l inside the last tile row / column are ZERO (the
//   GEMM's operand columns there are zero), which is also what an out-of-map tap reads in the reference
//   (grid_sample zeros padding), so only whole tiles outside the map need a bounds check.
//   Every level is padded to a multiple of 32 elements: rows are multiples of 64 B and a 16-element (32 B) store
//   never straddles a level.
#pragma once
#include <cstdint>

namespace eem {

constexpr int kPackedMaxLevels = 8;
constexpr int kPackedTile = 4;                       // 4x4 pixels
constexpr int kPackedTileElems = kPackedTile * kPackedTile;

struct PackedLayout {
  int L;
  int h[kPackedMaxLevels], w[kPackedMaxLevels];
  int tx[kPackedMaxLevels], ty[kPackedMaxLevels];    // tiles per row / per column
  int off[kPackedMaxLevels];                         // first element of level l inside a row
  int len[kPackedMaxLevels];                         // padded elements of level l (multiple of 32; 0 for an empty level)
  int row;                                           // elements per source position (multiple of 32)
};

inline PackedLayout packed_layout(int H, int W, int L) {
  PackedLayout pl{};
  pl.L = L;
  int h = H, w = W, off = 0;
  for (int l = 0; l < L && l < kPackedMaxLevels; ++l) {
    pl.h[l] = h;
    pl.w[l] = w;
    pl.tx[l] = (w + kPackedTile - 1) / kPackedTile;
    pl.ty[l] = (h + kPackedTile - 1) / kPackedTile;
    const int n = (h > 0 && w > 0) ? pl.tx[l] * pl.ty[l] * kPackedTileElems : 0;
    pl.len[l] = (n + 31) / 32 * 32;
    pl.off[l] = off;
    off += pl.len[l];
    h /= 2;
    w /= 2;
  }
  pl.row = off;
  return pl;
}

// element index of map cell (y, x) of a level inside that level's block
__host__ __device__ __forceinline__ int packed_cell(int y, int x, int tiles_x) {
  return (((y >> 2) * tiles_x + (x >> 2)) << 4) + ((y & 3) << 2) + (x & 3);
}

}  // namespace eem
