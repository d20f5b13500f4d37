/* kbnufft_oracle.c -- CPU ORACLE, TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's table-interpolation algorithm
 * (mmuckley/torchkbnufft, torchkbnufft/_nufft/interp.py).  It exists to CHECK
 * the CUDA engine in torchkbnufft_b200/ and to serve as the CPU baseline that
 * bench.py times; nothing in the product path may import, link or call it.
 * Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline /
 * --impl reference) use it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this oracle against
 *  (i) the reference's own golden vectors (tests/data/interp_data.pkl and
 *      nufft_data.pkl, re-saved as tests/golden/ref_*.npz), and
 *  (ii) outputs and integer indices of the reference itself, generated in the
 *      build container by oracle/make_golden.py (committed with the vectors).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define REAL float
#define FN(name) CAT(name, _f32)
#define FLOOR floorf
#define RINT rintf
#define COS cosf
#define SIN sinf
#define FMOD fmodf
#include "kbnufft_oracle_body.inc"
#undef REAL
#undef FN
#undef FLOOR
#undef RINT
#undef COS
#undef SIN
#undef FMOD

#define REAL double
#define FN(name) CAT(name, _f64)
#define FLOOR floor
#define RINT rint
#define COS cos
#define SIN sin
#define FMOD fmod
#include "kbnufft_oracle_body.inc"

int orc_abi_version(void) { return 1; }
