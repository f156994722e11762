// Tile rasterisation of the persistent GEMM (shared by gemm.cu and a host-side unit test: tests/test_tile_coords_host.py).
#pragma once
#if defined(__CUDACC__)
#define CLIPK_HD __host__ __device__ __forceinline__
#else
#define CLIPK_HD inline
#endif

namespace clipk {

// Tile id -> (row tile, column tile).  group_m == 1: column tiles fastest (the CTAs running together share a few A row tiles and all
// of a small B).  group_m > 1 (B far larger than L2, e.g. the retrieval gallery): ids sweep a [group_m x n_tiles] band column by
// column, so every B tile is fetched from HBM once per band instead of once per row tile.  Every (mi, ni) with mi < m_tiles,
// ni < n_tiles is produced by exactly one mn in [0, m_tiles * n_tiles), also when group_m does not divide m_tiles.
CLIPK_HD void tile_coords_raw(int m_tiles, int n_tiles, int group_m, int mn, int& mi, int& ni) {
  if (group_m <= 1) { mi = mn / n_tiles; ni = mn - mi * n_tiles; return; }
  const int band = group_m * n_tiles;
  const int g = mn / band;
  const int first = g * group_m;
  const int left = m_tiles - first;
  const int gm = group_m < left ? group_m : left;
  const int rem = mn - g * band;
  ni = rem / gm;
  mi = first + (rem - ni * gm);
}

}  // namespace clipk
