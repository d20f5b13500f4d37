// b2n_fft_args.cuh -- argument blocks shared by the run-time (b2n_fft.cu) and the compile-time planned
// (b2n_fft_fast_kernels.cuh, b2n_fft_plans_*.cu) FFT passes.
#pragma once
#include "b2n_common.cuh"
#include "b2n_fft_core.cuh"
#include "b2n_peer.cuh"

namespace b2n {

struct FftStages {
  int n, n_stages, radix[kFftMaxStages];
};

enum RowMode { ROW_PLAIN = 0, ROW_FWD_FIRST = 1 };

struct RowArgs {
  FftStages st;
  int n_in, n_out;       // nonzero inputs / kept outputs per line
  int64_t lines;         // number of lines
  int64_t rows_per_img;  // image rows per (batch, coil) = prod of the slower image dims
  int C, Ci, Bs;         // coils, image coils (1 or C), smaps batch (1 or B)
  const float2 *in;      // ROW_PLAIN: [lines][n_in]
  float2 *out;           // [lines][n_out]
  const float2 *image, *smaps, *scaling;  // ROW_FWD_FIRST operands
  const float2 *tw;      // twiddle table exp(-2 pi i t / n), t < n (b2n_fft_twiddles)
  float scale;
  // k_fft_rows_sense only: coil groups per image row, per-group partial rows, per-row arrival counters
  int coil_groups;
  float2 *partial;
  unsigned int *counter;
  int prefetch;          // CTAs ahead whose operand rows this CTA pulls into L2 (0: off)
  // k_fft_rows_sense only: sum all-reduce of the coil-combined image over peer memory, fused into the pass (each
  // finished row goes straight into the peers' windows, b2n_peer.cuh).  The caller sets peer.world > 1 to ask for it;
  // the launcher clears it when the pass cannot carry the exchange and reports what happened in peer_fused.
  PeerArgs peer;
  int64_t peer_floats;   // floats of the whole image (recorded as the size of this call's slot generation)
  int peer_fused;        // host side, out: 1 when the launched kernel performed the all-reduce
};

struct ColArgs {
  FftStages st;
  int n_in, n_out;   // rows read / rows written along the transformed dimension
  int64_t A, X;      // outer count, inner (contiguous) extent
  const float2 *in;  // [A][n_in][X]
  float2 *out;       // [A][n_out][X]
  const float2 *mul; // optional [mul_batch][n][X] factor applied to the inputs (Toeplitz kernel)
  int64_t a_per_mul; // outer indices per mul batch entry (0: single kernel)
  const float2 *tw;  // twiddle table of length n
  float scale;
  int prefetch;      // CTAs ahead whose input columns this CTA pulls into L2 (0: off)
};

}  // namespace b2n
