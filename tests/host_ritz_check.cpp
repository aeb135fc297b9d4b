// Test driver for ptz-calib_b200/csrc/cg_ritz.hpp (host part of the CG deflation): reads m, kmax and m rows of
// (alpha, beta, gamma) from stdin, prints the kept Ritz values and the coefficient matrix Y.
#include <stdio.h>

#include "../ptz-calib_b200/csrc/cg_ritz.hpp"

int main() {
  int m = 0, kmax = 0;
  if (scanf("%d %d", &m, &kmax) != 2) return 1;
  std::vector<double> abg(3 * (size_t)m);
  for (int j = 0; j < 3 * m; ++j)
    if (scanf("%lf", &abg[j]) != 1) return 1;
  std::vector<double> Y, val;
  const int kd = ptz::lowest_ritz_vectors(abg.data(), m, kmax, kmax, Y, &val);
  printf("%d\n", kd);
  for (int c = 0; c < kd; ++c) printf("%.17g ", val[c]);
  printf("\n");
  for (int j = 0; j < m; ++j) {
    for (int c = 0; c < kmax; ++c) printf("%.17g ", Y[(size_t)j * kmax + c]);
    printf("\n");
  }
  return 0;
}
