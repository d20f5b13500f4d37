// b2n_tiling.cuh -- tiled ordering of base grid cells shared by the plan builder and
// the interpolation kernels.  Points are binned by the tile of their (wrapped) base
// cell so that one CTA can stage a tile (+ J-1 halo) of the grid in shared memory.
#pragma once
#include "b2n_common.cuh"

namespace b2n {

struct Tiling {
  int ndim;
  int T[B2N_MAX_DIMS];   // tile edge per dim (1 for unused dims)
  int nt[B2N_MAX_DIMS];  // tiles per dim (1 for unused dims)
  int TT;                // cells per tile
  int64_t n_tiles;       // tiles per trajectory
  int64_t n_cells;       // n_tiles * TT
};

static inline Tiling make_tiling(int ndim, const int64_t *K) {
  Tiling t;
  t.ndim = ndim;
  const int edge = ndim == 1 ? 64 : (ndim == 2 ? 16 : 8);
  t.TT = 1;
  t.n_tiles = 1;
  for (int d = 0; d < B2N_MAX_DIMS; ++d) {
    t.T[d] = d < ndim ? edge : 1;
    t.nt[d] = d < ndim ? (int)((K[d] + edge - 1) / edge) : 1;
    t.TT *= t.T[d];
    t.n_tiles *= t.nt[d];
  }
  t.n_cells = t.n_tiles * t.TT;
  return t;
}

// tiled index of a wrapped cell (g[d] in [0, K_d)); dims beyond ndim must be 0
B2N_HD int64_t tiled_cell(const Tiling &t, const int64_t *g) {
  int64_t tile = 0, local = 0;
  for (int d = 0; d < t.ndim; ++d) {
    const int64_t td = g[d] / t.T[d];
    tile = tile * t.nt[d] + td;
    local = local * t.T[d] + (g[d] - td * t.T[d]);
  }
  return tile * t.TT + local;
}

}  // namespace b2n
