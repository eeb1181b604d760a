/* CPU restatement of the stencil hot path (plain C + OpenMP) -- TEST INFRASTRUCTURE ONLY.
 * See oracle/__init__.py for the rules and the parity status; oracle_body.inc holds the loop
 * nests with their reference citations.  Built by oracle/Makefile into oracle/_build/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define REAL double
#define INT int64_t
#define EXP exp
#define FN(name) orc_##name##_f64
#include "oracle_body.inc"
#undef REAL
#undef INT
#undef EXP
#undef FN

#define REAL float
#define INT int32_t
#define EXP expf
#define FN(name) orc_##name##_f32
#include "oracle_body.inc"
#undef REAL
#undef INT
#undef EXP
#undef FN

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
