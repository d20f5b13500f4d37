// b2n_interp.cuh -- argument block shared by the generic and the tiled interpolation kernels.
#pragma once
#include "b2n_common.cuh"
#include "b2n_tiling.cuh"

namespace b2n {

int validate_geom(const b2n_geom *g, bool need_tables);
extern long long *g_trace_buffer;
extern int64_t g_trace_capacity;

template <typename T> struct InterpArgs {
  int J[B2N_MAX_DIMS];
  int coef_off[B2N_MAX_DIMS];
  int coef_stride;
  int64_t K[B2N_MAX_DIMS];
  int64_t Kprod;
  int64_t M, n_traj, B, C;
  const int32_t *perm;
  const int32_t *inv_perm;
  const int32_t *base;
  const cplx<T> *coef;
  const cplx<T> *phase;
  const int32_t *cell_start;
  const int32_t *sub_tile, *sub_start, *sub_count, *n_sub;
  const int32_t *sub_slot, *tile_sub_start;
  int64_t n_sub_max;
  int sub_cap;
  Tiling tiling;
  long long *trace;  // optional per-CTA timeline (b2n_set_trace_buffer), NULL in production
  int64_t trace_cap;
};

template <typename T>
static inline int make_args(const b2n_geom *g, const b2n_points *p, int64_t B, int64_t C, InterpArgs<T> *a) {
  int rc = validate_geom(g, false);
  if (rc) return rc;
  if (!p || !p->perm || !p->base || !p->coef || !p->phase || !p->cell_start)
    return fail_arg(B2N_E_ARG, "points plan is NULL or not built");
  if (p->ndim != g->ndim || p->dtype != g->dtype) return fail_arg(B2N_E_ARG, "plan/geometry mismatch");
  if (B < 1 || C < 1) return fail_arg(B2N_E_ARG, "n_batch=%lld n_coils=%lld", (long long)B, (long long)C);
  if (p->n_traj != 1 && p->n_traj != B)
    return fail_arg(B2N_E_ARG, "plan has %lld trajectories but n_batch=%lld", (long long)p->n_traj, (long long)B);
  int off = 0;
  a->Kprod = 1;
  for (int d = 0; d < B2N_MAX_DIMS; ++d) {
    a->J[d] = d < g->ndim ? g->numpoints[d] : 1;
    a->K[d] = d < g->ndim ? g->grid_size[d] : 1;
    a->coef_off[d] = off;
    if (d < g->ndim) {
      off += g->numpoints[d];
      a->Kprod *= g->grid_size[d];
    }
  }
  a->coef_stride = off;
  if (off != p->coef_stride) return fail_arg(B2N_E_ARG, "plan coef_stride mismatch");
  a->M = p->n_points;
  a->n_traj = p->n_traj;
  a->B = B;
  a->C = C;
  a->perm = p->perm;
  a->inv_perm = p->inv_perm;
  a->base = p->base;
  a->coef = (const cplx<T> *)p->coef;
  a->phase = (const cplx<T> *)p->phase;
  a->cell_start = p->cell_start;
  a->sub_tile = p->sub_tile;
  a->sub_start = p->sub_start;
  a->sub_count = p->sub_count;
  a->n_sub = p->n_sub;
  a->sub_slot = p->sub_slot;
  a->tile_sub_start = p->tile_sub_start;
  a->n_sub_max = p->n_sub_max;
  a->sub_cap = p->sub_cap;
  a->trace = g_trace_buffer;
  a->trace_cap = g_trace_capacity;
  a->tiling = make_tiling(g->ndim, g->grid_size);
  for (int d = 0; d < B2N_MAX_DIMS; ++d)
    if (a->tiling.T[d] != p->tile[d] || a->tiling.nt[d] != p->n_tiles[d])
      return fail_arg(B2N_E_ARG, "plan tiling does not match this library build");
  return 0;
}


}  // namespace b2n
