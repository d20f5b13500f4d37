/* b200nufft.h -- C ABI of libb200nufft.so, the sm_100a table-interpolation
 * NUFFT engine behind torchkbnufft's table-mode API.
 *
 * Boundary rules
 *  - extern "C", plain pointers and sizes only (no torch / C++ types);
 *  - every `*_dev` / grid / kdata pointer is DEVICE memory owned by the caller;
 *    the library never allocates or frees caller-visible memory -- plans live
 *    in a caller-provided workspace sized by b2n_points_workspace_bytes();
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and
 *    no call synchronises the device (CUDA-graph capturable) except where noted;
 *  - return value: 0 = OK, < 0 = argument error (B2N_E_*), > 0 = cudaError_t;
 *    b2n_last_error() returns a thread-local message for the last failure;
 *  - complex data is interleaved (re, im) float (B2N_C64) or double (B2N_C128);
 *  - no CPU fallback exists: without a CUDA device every compute entry fails.
 *
 * Each entry point names the reference interface it replaces; paths are
 * relative to the reference checkout (mmuckley/torchkbnufft).
 */
#ifndef B200NUFFT_H
#define B200NUFFT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B2N_API __attribute__((visibility("default")))
#else
#define B2N_API
#endif

#define B2N_ABI_VERSION 4
#define B2N_MAX_DIMS 3
#define B2N_MAX_NUMPOINTS 16 /* max neighbours J per dimension */

enum b2n_dtype { B2N_C64 = 0, B2N_C128 = 1 };
/* grid memory layout: (B, C, *K) as in the reference, or channel-last (B, *K, C) */
enum b2n_layout { B2N_COIL_MAJOR = 0, B2N_CHANNEL_LAST = 1 };
/* adjoint accumulation mode */
enum b2n_adjoint_mode {
  B2N_ADJ_ATOMIC = 0, /* scatter with L2 reductions; order not reproducible (like index_add_ on CUDA) */
  B2N_ADJ_SORTED = 1  /* gather per grid cell over the cell-sorted point list; bit-reproducible */
};

enum b2n_error {
  B2N_OK = 0,
  B2N_E_ARG = -1,       /* bad argument (null pointer, bad enum, size <= 0) */
  B2N_E_RANGE = -2,     /* size exceeds an engine limit (dims, J, 32-bit cell keys) */
  B2N_E_WORKSPACE = -3, /* workspace too small or misaligned */
  B2N_E_UNSUPPORTED = -4
};

/* Geometry of one NUFFT operator: what KbModule registers as buffers
 * (torchkbnufft/modules/_kbmodule.py:53-62; built by init_fn, _nufft/utils.py:257-358). */
typedef struct b2n_geom {
  int32_t ndim;                         /* 1..3 */
  int32_t dtype;                        /* enum b2n_dtype */
  int64_t grid_size[B2N_MAX_DIMS];      /* K_d, oversampled grid */
  int32_t numpoints[B2N_MAX_DIMS];      /* J_d */
  int32_t table_oversamp[B2N_MAX_DIMS]; /* L_d */
  int64_t table_len[B2N_MAX_DIMS];      /* J_d*L_d + 1 */
  const void *table_dev[B2N_MAX_DIMS];  /* device, interleaved complex of `dtype` */
  double n_shift[B2N_MAX_DIMS];         /* fftshift phase offsets (values of the real-dtype buffer) */
  /* Optional: the tables the reference builds are a REAL kernel times a linear phase,
   * table_d[i] = r_d[i] * exp(-1i * table_phase[d] * x_i), x_i = i / L_d - J_d / 2, table_phase[d] = pi (N_d - 1) / K_d
   * (torchkbnufft/_nufft/utils.py:160-204, Fessler's column trick).  When the caller has verified that for its
   * tables it passes r_d (device, real of the same precision, table_len[d] entries) and the slopes; the spread then
   * works with real weights (half the arithmetic): the phase factors into a per-point, a per-grid-cell and a wrap-sign
   * part.  NULL pointers = tables are used as they are. */
  const void *rtable_dev[B2N_MAX_DIMS];
  double table_phase[B2N_MAX_DIMS];
} b2n_geom;

/* A trajectory plan: points sorted by the TILE of their wrapped base grid cell (then by
 * cell inside the tile, then by original order) with everything that depends on omega
 * alone precomputed once.  The reference recomputes all of it on every call: tm /
 * base_offset at _nufft/interp.py:171-177 and :663-670, the per-offset table lookups of
 * calc_coef_and_indices :129-148, sort_data :566-584, the phase :200-203.
 * Tiled cell index of base cell g: tile_id(g) * prod(tile) + local(g), row-major in both.
 * Work is cut into sub-problems of at most sub_cap consecutive points of one tile (the
 * dense centre of a radial trajectory yields many sub-problems, the periphery one).
 * All pointers point into the caller's workspace. */
typedef struct b2n_points {
  int64_t n_points;    /* M, points per trajectory */
  int64_t n_traj;      /* 1 (shared trajectory) or B (one per batch element) */
  int32_t ndim;
  int32_t dtype;
  int32_t coef_stride; /* complex entries per point record = sum_d J_d */
  int32_t sub_cap;     /* max points per sub-problem */
  int32_t tile[B2N_MAX_DIMS];    /* tile edge (in base cells) per dimension */
  int32_t n_tiles[B2N_MAX_DIMS]; /* tiles per dimension = ceil(K_d / tile_d) */
  int64_t n_cells;     /* per trajectory: prod(n_tiles) * prod(tile)  (>= prod K) */
  int64_t n_sub_max;   /* capacity of the sub_* arrays (upper bound on *n_sub) */
  int32_t *perm;       /* [n_traj*M]        sorted slot -> point index inside its trajectory */
  int32_t *inv_perm;   /* [n_traj*M]        traj*M + point index -> sorted slot */
  int32_t *base;       /* [n_traj*M][ndim]  wrapped base cell, in [0, K_d) */
  void *coef;          /* [n_traj*M][coef_stride] complex weights T_d[t_d(j)], dims concatenated; the
                          fftshift phase of the point is folded into the dimension-0 entries */
  void *phase;         /* [n_traj*M]        complex exp(i sum_d omega_d n_shift_d) */
  int32_t *cell_start; /* [n_traj*n_cells+1] CSR offsets of the sorted list per tiled base cell */
  uint32_t *keys;      /* [n_traj*M]        sorted keys (traj*n_cells + tiled base cell) */
  int32_t *sub_tile;   /* [n_sub_max]       traj*prod(n_tiles) + tile id of each sub-problem */
  int32_t *sub_start;  /* [n_sub_max]       first sorted slot */
  int32_t *sub_count;  /* [n_sub_max]       number of points (<= sub_cap) */
  int32_t *n_sub;      /* device scalar: number of sub-problems actually used */
  int32_t *sub_slot;   /* [n_sub_max]       tile-major rank of each sub-problem: the sub-problems of (traj, tile) t have
                          the consecutive ranks [tile_sub_start[t], tile_sub_start[t+1]) (the sub_* arrays themselves
                          are ordered longest-first for load balance) */
  int32_t *tile_sub_start; /* [n_traj*prod(n_tiles)] first rank of each tile; the last tile ends at *n_sub */
  /* Owner-tile visit lists for the output-stationary spread (b2n_interp_adjoint_ordered; 2-D / 3-D complex64 J = 6
   * plans on grids with every K_d >= 16 whose geometry carries rtable_dev and power-of-two L_d, so that the table index
   * of neighbour j is exactly that of neighbour 0 minus j L_d; own_tile == 0 and NULL pointers otherwise).  The grid is
   * cut into OUTPUT tiles of 4 x 8 cells (2-D) or 4 x 4 x 8 cells (3-D; 8 along the last, contiguous axis); a "visit" is
   * one (point, output tile) pair whose footprint intersects the tile (3.66 per point on average in 2-D, 8.2 in 3-D);
   * the visits of a tile are listed in a fixed order (window cell
   * row-major, then sorted slot) and cut into work items of at most own_cap visits.  Every output tile has at
   * least one item (an empty one writes zeros), so the spread needs no zero-initialised grid and no atomics. */
  int32_t own_tile;          /* rows of an output tile (4; 8 columns), or 0 when the lists were not built */
  int32_t own_cap;           /* max visits per work item */
  int32_t n_own_tiles[3];    /* output tiles per dimension: ceil(K_d / own_tile), ceil(K_last / 8) along the last one */
  int32_t own_pad_;
  int64_t n_own_items_max;   /* capacity of own_items (upper bound on own_counts[0]) */
  void *own_visits;          /* 2-D: 64-byte records [<= 6*n_traj*M, see above]: float hy[4] (row weights r_y[row - ry]),
                                float hx[8] (column weights +-r_x[column - rx]; zero outside the footprint; the sign
                                is negative when the footprint wrapped around the grid and exp(1i table_phase K) =
                                -1), int32 sample index inside its trajectory, 12 unused bytes; (ry, rx) = base
                                cell minus tile origin.
                                3-D: int32x4 index records [<= 18*n_traj*M]: {sorted slot, sample index,
                                (r0+16) | (r1+16) << 8 | (r2+16) << 16, sign flag}; the kernel forms the window
                                weights from own_hw while staging */
  void *own_items;           /* int32x4 [n_own_items_max]: {traj*prod(n_own_tiles) + tile, first visit, visits | chunk
                                index inside the tile << 12, tile row << 16 | tile column (2-D) or the tile's row-major
                                index (3-D)}, longest first */
  void *own_tiles;           /* int32x4 [n_traj*prod(n_own_tiles)]: {visits, first visit, chunks, first partial-sum slot
                                (-1 for single-chunk tiles)} */
  int32_t *own_counts;       /* device: [0] = items in use, [1] = partial-sum slots in use, [2] = exception points,
                                [3] = entries of own_xv in use */
  float *own_hw;             /* [n_traj*M][12] (2-D): r_y[0..5], r_x[0..5] of the point's neighbours; [n_traj*M][24] (3-D):
                                r_0, r_1, r_2, -r_2 */
  void *own_fac;             /* complex64 [n_traj*M]: conj of the point's phase factor (adjoint form): fftshift phase
                                times exp(-1i sum_d table_phase[d] (x_d(neighbour 0) + wrapped base_d)) */
  void *own_q;               /* complex64 [sum_d K_d]: exp(-1i table_phase[d] cell), the per-cell factor of the adjoint */
  /* Exception points: the factored phase needs table_index(neighbour j) == table_index(neighbour 0) - j L_d.  That
   * holds whenever tm - (base + j) is exact in the trajectory's precision; it can fail by one table step for a
   * point within J cells of the k-space origin whose distance to a neighbour is a rounding tie (radial trajectories
   * sampled on exact multiples of the grid spacing have a few thousand of them around the centre).  Such points get
   * zero weights in own_hw / own_visits and are spread afterwards with their complex records (coef) by a fix-up kernel
   * that is output-stationary as well: every output tile has the list of the exception points that reach it, in
   * ascending slot order, and adds their contributions to its cells once -- deterministic, and the result keeps the
   * reference's table indices for every neighbour. */
  int32_t *own_exc;          /* [n_own_exc_max] sorted slots of the exception points, ascending */
  int64_t n_own_exc_max;     /* capacity of own_exc = n_traj*M; a caller that has read own_counts[2] back may lower it
                                (0 = no fix-up launch) */
  void *own_xt;              /* int32x2 [n_traj*prod(n_own_tiles)]: {first entry, entries} of the tile in own_xv */
  void *own_xv;              /* int32x2 [n_own_xv_max]: {sorted slot, (r0+16) | (r1+16) << 8 | (r2+16) << 16} per
                                (exception point, output tile) pair, r = footprint origin minus tile origin */
  int64_t n_own_xv_max;      /* capacity of own_xv (own_counts[3] = entries in use; the fix-up traps beyond it) */
} b2n_points;

/* engine options (process-wide; for A/B measurements and tests) */
enum b2n_option {
  B2N_OPT_TILED_KERNELS = 0, /* 1 (default): shared-memory tiled kernels where they apply and pay off (not for a single
                                2-D (batch, coil) row); 2: wherever they apply; 0: per-point kernels only */
  B2N_OPT_ADJ_ROW_OWNERSHIP = 1, /* tiled 2-D adjoint variant: 0 (default) auto = warp-owned tile rows for 16-coil CTAs,
                                    warp-private 8-coil tiles otherwise; 1 warp-owned rows, 2 warp-owned coils,
                                    3 warp-private tiles, 4 / 5 warp-owned rows with 4 / 2 warps per CTA, 6 taps-per-lane
                                    kernel up to 4 coils (auto uses it for 1-2 coils); kept for A/B measurements */
  B2N_OPT_FWD_COIL_CHUNK = 2, /* tiled forward for C > 8: 0 (default) one 16-coil CTA per sub-problem, 1 persistent
                                 triple-buffered kernel (measured slower, kept for A/B), 8 two 8-coil CTAs */
  B2N_OPT_ADJ_COIL_CHUNK = 3, /* 0 (default): 16 coils per CTA in the tiled adjoint; 8: two 8-coil CTAs */
  B2N_OPT_FAST_FFT = 4, /* 1 (default): fused FFT passes use the compile-time planned kernels for the lengths that
                           have a plan (64, 72, 96, 120, 128, 144, 160, 192, 200, 224, 240, 256, 288, 320, 360, 384, 400, 448, 480, 512, 576, 600, 640, 720, 768, 800, 896, 960, 1024, 1152, 1200, 1280, 1440, 1536, 1600, 1920, 2048); 0: run-time passes only */
  B2N_OPT_PDL = 5, /* 1 (default): the FFT passes, the gathers and the tiled 2-D spread are launched with programmatic
                      dependent launch (their prologues overlap the tail of the preceding kernel; the spread's adjoint
                      grid is zeroed by a kernel it overlaps with); 2: the same but the grid is zeroed by
                      cudaMemsetAsync; 3: the same as 1 but the coil-sum counters are zeroed right before their kernel (A/B);
                      0: plain stream order */
  B2N_OPT_FFT_PREFETCH = 6, /* planned FFT passes whose CTAs pull the operand rows of the CTA one wave ahead into L2, as a
                               mask: 1 forward rows, 2 forward columns, 4 inverse columns, 8 inverse rows, 16 Toeplitz
                               columns; a pass must also read >= 32 MB unless 32 is set.  Default 19. */
  B2N_OPT_ADJ_OWNED = 7, /* 1 (default): b2n_interp_adjoint_ordered uses the output-stationary owner-tile spread where
                            the plan carries visit lists (2-D complex64 J = 6); 0: the scratch-tile + merge kernels */
  B2N_OPT_OWN_CAP = 8, /* visits per work item of the owner-tile spread, read when a plan is built (0 = default: 128 in
                          2-D, 1024 in 3-D; at most 4095) */
  B2N_OPT_FFT_STREAM = 9, /* planned FFT passes that run as persistent kernels whose CTAs keep the operands of their next
                             tile in flight (asynchronous copies into shared memory) while they transform the current
                             one, as a mask: 1 forward columns, 2 inverse columns (16 / 32: the same with half as many
                             columns per CTA, for A/B).  Default 1: -3 % on the 384^2 x 32-coil forward, neutral where
                             the input is L2-resident; the inverse pass holds twice the operands per tile and loses a
                             resident CTA (profiles/r02_fft_stream_ab.log).  Results are bit-identical either way. */
  B2N_OPT_PEER_FORM = 10, /* b2n_peer_allreduce_sum: 0 (default) one-shot or two-shot exchange by message size and number
                             of ranks, 1 always one-shot, 2 always two-shot (reduce-scatter + all-gather through the same
                             windows; needs float4-aligned operands) */
  B2N_OPT_COUNT
};
B2N_API int b2n_set_option(int option, int value);
B2N_API int b2n_get_option(int option);
/* Development aid: when a device buffer of `capacity` records (6 x int64 each) is set, every CTA
 * of the tiled kernels writes {SM id, points, t_start, t_staged, t_done, flags} (globaltimer ns)
 * at its linear block index.  Pass NULL to switch tracing off (the default). */
B2N_API int b2n_set_trace_buffer(void *records_dev, int64_t capacity);

B2N_API int b2n_abi_version(void);
/* sizeof(b2n_geom) / sizeof(b2n_points) as compiled into the library: a binding checks its own struct mirrors */
B2N_API int b2n_struct_sizes(size_t *geom_bytes, size_t *points_bytes);
B2N_API const char *b2n_last_error(void);
/* Number of kernels of this library launched by the process so far (every launch of an own kernel counts; memsets and
 * cuFFT do not). */
B2N_API long long b2n_launch_count(void);
/* number of CUDA devices visible; 0 when there is none or the driver is unusable
 * (b2n_last_error() then says why) */
B2N_API int b2n_device_count(void);

/* ---- trajectory plan -------------------------------------------------------- */
/* Bytes of device workspace b2n_points_build needs (plan arrays + sort scratch). */
B2N_API int b2n_points_workspace_bytes(const b2n_geom *geom, int64_t n_points, int64_t n_traj, size_t *bytes);

/* Build the plan for `omega_dev` ([n_traj][ndim][M], real dtype of geom->dtype,
 * radians/voxel).  Replaces, once per trajectory, the per-call coordinate work of
 * table_interp_one_batch (_nufft/interp.py:171-177), calc_coef_and_indices
 * (:129-148) and sort_one_batch (:552-562; integer cell keys, stable). */
B2N_API int b2n_points_build(const b2n_geom *geom, const void *omega_dev, int64_t n_points, int64_t n_traj,
                     void *workspace_dev, size_t workspace_bytes, b2n_points *out, void *stream);

/* Debug / parity export of the reference's integer indices for every neighbour
 * offset w (row-major): arr_ind[W][M] (flat wrapped grid index, int64) and
 * tab_idx[W][ndim][M] (table index incl. centre, int32; may be NULL).
 * reference: calc_coef_and_indices, _nufft/interp.py:89-150.  Single trajectory. */
B2N_API int b2n_export_indices(const b2n_geom *geom, const void *omega_dev, int64_t n_points, int64_t *arr_ind_dev,
                       int32_t *tab_idx_dev, void *stream);

/* ---- table interpolation ---------------------------------------------------- */
/* Forward gather, grid -> points.  grid: (B, C, *K) or (B, *K, C) per `grid_layout`;
 * kdata out: (B, C, M).  reference: table_interp, _nufft/interp.py:315-403. */
B2N_API int b2n_interp_forward(const b2n_geom *geom, const b2n_points *pts, const void *grid_dev, int64_t n_batch,
                               int64_t n_coils, int grid_layout, void *kdata_dev, void *stream);

/* Adjoint spread, points -> grid (grid fully overwritten).
 * reference: table_interp_adjoint, _nufft/interp.py:587-726 (+ accum_tensor_index_add :407-419). */
B2N_API int b2n_interp_adjoint(const b2n_geom *geom, const b2n_points *pts, const void *kdata_dev, int64_t n_batch,
                               int64_t n_coils, int grid_layout, int mode, void *grid_dev, void *stream);

/* ---- fused steps around the FFT --------------------------------------------- */
/* grid = zero_pad_end( image * smaps * scaling ) * scale.
 *   image (B, Ci, *N) with Ci == C, or Ci == 1 broadcast over coils (SENSE);
 *   smaps (Bs, C, *N) or NULL, Bs in {1, B};  scaling (*N) complex or NULL.
 * reference: SENSE multiply modules/kbnufft.py:182-183 + fft_and_scale
 * _nufft/fft.py:66-76 (multiply, F.pad); `scale` folds the 'ortho' factor. */
B2N_API int b2n_apod_pad(int ndim, int dtype, const int64_t *im_size, const int64_t *grid_size, int64_t n_batch,
                 int64_t n_coils, const void *image_dev, int64_t image_coils, const void *smaps_dev,
                 int64_t smaps_batch, const void *scaling_dev, double scale, int grid_layout, void *grid_dev,
                 void *stream);

/* image = sum_c? ( crop_first_N(grid) * conj(scaling) * conj(smaps) ) * scale.
 *   with smaps: out (B, 1, *N) (coil combine); without: out (B, C, *N).
 * reference: ifft_and_scale _nufft/fft.py:113-118 (crop_dims :25-32) + coil
 * combine modules/kbnufft.py:404-405. */
B2N_API int b2n_crop_apod_coilsum(int ndim, int dtype, const int64_t *im_size, const int64_t *grid_size, int64_t n_batch,
                          int64_t n_coils, const void *grid_dev, int grid_layout, const void *smaps_dev,
                          int64_t smaps_batch, const void *scaling_dev, double scale, void *image_dev,
                          void *stream);

/* spectrum[b][c][k] *= kernel[bk][k] * scale, in place; kernel_batch in {1, B}.
 * reference: the Toeplitz filter multiply of fft_filter, _nufft/fft.py:164-173. */
B2N_API int b2n_spectrum_mul(int dtype, void *spectrum_dev, const void *kernel_dev, int64_t n_batch, int64_t n_coils,
                     int64_t n_grid, int64_t kernel_batch, int grid_layout, double scale, void *stream);

/* Deterministic adjoint on the tiled kernels (complex64, J = 6, coil-major; 2-D with every K_d >= 21 or 3-D with every
 * K_d >= 13): each sub-problem
 * accumulates its tile in shared memory in a fixed order and writes it to its own slot of `scratch_dev`; a second
 * kernel adds, for every grid cell, the slots that cover it in a fixed order.  Bit-reproducible run to run, ~10x
 * faster than B2N_ADJ_SORTED (which remains the fallback for every other case).
 * b2n_interp_adjoint_ordered_bytes: scratch size in bytes, 0 when this path does not apply. */
B2N_API int b2n_interp_adjoint_ordered_bytes(const b2n_geom *geom, const b2n_points *pts, int64_t n_batch, int64_t n_coils,
                                             int grid_layout, size_t *bytes);
/* The same query plus `zero_bytes`: the leading bytes of the scratch that hold arrival counters.  They must be zero
 * when a scratch buffer is first handed to b2n_interp_adjoint_ordered and are left zero by every completed call, so a
 * caller that keeps the buffer between calls zeroes it once (0 for the paths without counters).  With
 * n_items / n_slots > 0 (the plan's own_counts read back by the caller) the size is exact instead of an upper bound. */
B2N_API int b2n_interp_adjoint_ordered_layout(const b2n_geom *geom, const b2n_points *pts, int64_t n_batch,
                                              int64_t n_coils, int grid_layout, int64_t n_slots, size_t *bytes,
                                              size_t *zero_bytes);
B2N_API int b2n_interp_adjoint_ordered(const b2n_geom *geom, const b2n_points *pts, const void *kdata_dev, int64_t n_batch,
                                       int64_t n_coils, int grid_layout, void *scratch_dev, size_t scratch_bytes,
                                       void *grid_dev, void *stream);

/* ---- pruned, fused FFT passes (complex64, coil-major) --------------------------------
 * Own shared-memory Stockham transforms, one dimension per pass, which skip the zero-padded
 * inputs / cropped outputs and fuse the element-wise work of the steps above:
 *   forward = b2n_apod_pad + fftn(norm=None) in ndim passes,
 *   adjoint = ifftn(norm="forward") + b2n_crop_apod_coilsum (+ optional Toeplitz kernel
 *             multiply on the way in) in ndim passes.
 * b2n_fft_supported(n): 0 = not handled; 1 = length n factors into {2,3,5,7,11,13} and its ping-pong line buffers
 * fit in shared memory (n <= 1528: run-time Stockham passes); 2 = n has a compile-time plan (register-resident
 * passes, the fast path).
 * work_dev: device scratch of b2n_fft_work_bytes() bytes (intermediate, partially transformed arrays, per-coil-group
 * partial rows and arrival counters of the fused coil sum); the forward does not need it for ndim == 1.
 * reference: fft_and_scale / ifft_and_scale / fft_filter, _nufft/fft.py:36-173. */
B2N_API int b2n_fft_supported(int64_t n);
/* Fill a caller buffer of 2*n complex64 entries: [0, n) = exp(-2 pi i t / n) (computed in double), [n, 2n) = the
 * per-stage twiddle tables of the compile-time plan for length n (unused when n has none). */
B2N_API int b2n_fft_twiddles(int64_t n, void *twiddle_dev, void *stream);
B2N_API int b2n_fft_work_bytes(int ndim, const int64_t *im_size, const int64_t *grid_size, int64_t n_batch,
                               int64_t n_coils, size_t *bytes);
/* twiddle_dev: ndim device pointers, entry d = table of b2n_fft_twiddles(grid_size[d]) */
B2N_API int b2n_fft_forward_fused(int ndim, const int64_t *im_size, const int64_t *grid_size, int64_t n_batch,
                                  int64_t n_coils, const void *image_dev, int64_t image_coils,
                                  const void *smaps_dev, int64_t smaps_batch, const void *scaling_dev, double scale,
                                  const void *const *twiddle_dev, void *grid_dev, void *work_dev, void *stream);
B2N_API int b2n_fft_adjoint_fused(int ndim, const int64_t *im_size, const int64_t *grid_size, int64_t n_batch,
                                  int64_t n_coils, const void *grid_dev, const void *kernel_dev,
                                  int64_t kernel_batch, const void *smaps_dev, int64_t smaps_batch,
                                  const void *scaling_dev, double scale, const void *const *twiddle_dev,
                                  void *image_dev, void *work_dev, void *stream);
/* Toeplitz normal operator  out[b] = scale * sum_c conj(S_c) crop(IFFT(kernel * FFT(zero_pad(S_c * image[b]))))
 * (unnormalised transforms) in THREE passes: forward rows, then one column pass that transforms, multiplies by the
 * kernel spectrum and transforms back inside the CTA, then inverse rows + coil sum -- the full-size spectrum is never
 * written to memory.  2-D complex64, grid_size[0] a compile-time planned length (b2n_fft_supported == 2) with
 * 2*im_size[0] <= grid_size[0]; other shapes return B2N_E_UNSUPPORTED and take b2n_fft_forward_fused +
 * b2n_fft_adjoint_fused (kernel_dev set).  kernel_dev: (kernel_batch, *grid_size), kernel_batch 1 or n_batch;
 * work_dev: b2n_fft_work_bytes.  reference: the per-batch loop of ToepNufft.forward (modules/kbnufft.py:441-484)
 * around fft_filter (_nufft/fft.py:121-173). */
B2N_API int b2n_fft_toeplitz_fused(int ndim, const int64_t *im_size, const int64_t *grid_size, int64_t n_batch,
                                   int64_t n_coils, const void *image_dev, int64_t image_coils,
                                   const void *smaps_dev, int64_t smaps_batch, const void *kernel_dev,
                                   int64_t kernel_batch, double scale, const void *const *twiddle_dev, void *out_dev,
                                   void *work_dev, void *stream);

/* ---- sum all-reduce over NVLink peer memory (coil-sharded SENSE adjoint; one process per GPU on one node) ------------
 * The reference has no multi-GPU code; the coupling point this replaces is the coil sum of the SENSE adjoint
 * (modules/kbnufft.py:404-405) when the coils are split over GPUs.  Each rank creates a *window* (library-owned device
 * memory -- the one exception to "the library never allocates": CUDA IPC needs the base of an allocation), hands its
 * 64-byte handle to the other ranks (any host channel: torch.distributed, MPI, a pipe), opens theirs, and from then on
 * one kernel per rank and call pushes the partial result into every peer's window, flags it, waits for the peers' and
 * adds the slots in rank order (bit-identical results on all ranks).  Every rank must issue the same sequence of
 * b2n_peer_allreduce_sum calls with the same n_floats; a window serves one stream at a time.  */
#define B2N_PEER_MAX_RANKS 16
#define B2N_PEER_HANDLE_BYTES 64
typedef struct b2n_peer_comm {
  int32_t rank, world;               /* this process, processes (<= B2N_PEER_MAX_RANKS) */
  int64_t max_floats;                /* capacity the windows were sized for (b2n_peer_window_bytes) */
  void *window[B2N_PEER_MAX_RANKS];  /* [world] device pointers in THIS process: own window, peers' mappings */
} b2n_peer_comm;
B2N_API int b2n_peer_window_bytes(int world, int64_t max_floats, size_t *bytes);
/* cudaMalloc + zero-fill + cudaIpcGetMemHandle; synchronises the device.  handle_out: B2N_PEER_HANDLE_BYTES bytes */
B2N_API int b2n_peer_window_create(size_t bytes, void **window_dev, void *handle_out);
B2N_API int b2n_peer_window_open(const void *handle, void **window_dev);  /* a peer's window, mapped here */
B2N_API int b2n_peer_window_close(void *window_dev);                      /* unmap a peer's window */
B2N_API int b2n_peer_window_destroy(void *window_dev);                    /* free this rank's own window */
/* out = sum over ranks of in (float32 units: a complex64 image is 2 floats per value); in place allowed; enqueued on
 * `stream`, CUDA-graph capturable (the call counter lives in the window).  One kernel: a one-shot exchange (every rank
 * pushes its whole message to every peer) for latency-bound sizes, a two-shot exchange (reduce-scatter + all-gather of
 * shards) where that moves enough fewer bytes to pay for its second link latency (B2N_OPT_PEER_FORM forces a form). */
B2N_API int b2n_peer_allreduce_sum(const b2n_peer_comm *comm, const void *in_dev, void *out_dev, int64_t n_floats,
                                   void *stream);
/* b2n_fft_adjoint_fused followed by the sum all-reduce of image_dev over the ranks of `comm` (every rank calls it
 * with its own coils; all end up with the full coil sum).  Where the last pass is the row pass with the fused coil
 * combination and its grid is resident as a whole, each finished image row goes straight from that kernel into the
 * peers' windows and comes back summed -- compute and collective in one launch; otherwise the stand-alone all-reduce
 * kernel runs behind the last pass.  CTAs of the fused pass wait for the same rows of the peers: work that co-runs on
 * other streams must terminate on its own (a kernel that holds its SMs until this call finishes can stall the exchange;
 * a wait of more than ~20 s traps instead of hanging the device).  Same arguments as b2n_fft_adjoint_fused plus `comm`.
 * reference coupling point:
 * the coil sum of the SENSE adjoint (modules/kbnufft.py:404-405) with the coils split over GPUs. */
B2N_API int b2n_fft_adjoint_fused_allreduce(int ndim, const int64_t *im_size, const int64_t *grid_size,
                                            int64_t n_batch, int64_t n_coils, const void *grid_dev,
                                            const void *kernel_dev, int64_t kernel_batch, const void *smaps_dev,
                                            int64_t smaps_batch, const void *scaling_dev, double scale,
                                            const void *const *twiddle_dev, void *image_dev, void *work_dev,
                                            const b2n_peer_comm *comm, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200NUFFT_H */
