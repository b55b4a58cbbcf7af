/* oracle/trig_check.c -- TEST INFRASTRUCTURE.  Compares oracle.c's restatement of glibc's atanf / sinf / cosf (trig_mode 1,
 * the algorithm the CUDA path runs in Engine::disrupt) with the libm of this machine over ALL 2^32 fp32 arguments.
 *   gcc -O2 -ffp-contract=off -o _ref/trig_check trig_check.c -L. -loracle -lm -Wl,-rpath,'$ORIGIN/..' && _ref/trig_check [step]
 * (`make -C oracle trig_check`; about ten minutes on one core for step 1).  Result on the build box (glibc 2.39, Xeon with
 * FMA): "atanf 0, sinf 0, cosf 0 mismatches of 4278190082 arguments".  sinf / cosf of |x| >= 120 (glibc: reduce_large) are
 * outside Engine::disrupt's range and NaN in the restatement: skipped. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
void oracle_trig_array(const float* in, float* out, int n, int which, int mode);
int main(int argc, char** argv) {
  const uint64_t step = argc > 1 ? strtoull(argv[1], 0, 10) : 1;
  enum { B = 1 << 16 };
  static float in[B], a[B], b[B];
  uint64_t bad[3] = {0, 0, 0}, tot = 0;
  for (uint64_t u0 = 0; u0 < 0x100000000ULL; u0 += step * B) {
    int n = 0;
    for (uint64_t u = u0; u < 0x100000000ULL && n < B; u += step) { uint32_t v = (uint32_t)u; float f; memcpy(&f, &v, 4); if (f == f) in[n++] = f; }
    tot += n;
    for (int which = 0; which < 3; which++) {
      oracle_trig_array(in, a, n, which, 0);
      oracle_trig_array(in, b, n, which, 1);
      for (int i = 0; i < n; i++) {
        if (which && !(fabsf(in[i]) < 120.0f)) continue;
        if (memcmp(&a[i], &b[i], 4) && !(a[i] != a[i] && b[i] != b[i])) {
          if (bad[which]++ < 4) printf("%s(%a): libm %a restatement %a\n", which == 0 ? "atanf" : which == 1 ? "sinf" : "cosf", in[i], a[i], b[i]);
        }
      }
    }
  }
  printf("atanf %llu, sinf %llu, cosf %llu mismatches of %llu arguments\n", (unsigned long long)bad[0], (unsigned long long)bad[1],
         (unsigned long long)bad[2], (unsigned long long)tot);
  return (bad[0] | bad[1] | bad[2]) != 0;
}
