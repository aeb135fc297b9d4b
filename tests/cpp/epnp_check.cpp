// Runs ptzcalib::epnp::solve_pnp_epnp (include/ptzcalib_epnp.hpp) on the cases of tests/golden/epnp_kat.npz, re-packed as a flat
// binary by tests/test_epnp.py; for every case the 8 sign choices of the control-point axes are written out (choice 0 = default).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/ptzcalib_epnp.hpp"

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  FILE* g = fopen(argv[2], "wb");
  if (!f || !g) return 2;
  int cases = 0;
  if (fread(&cases, 4, 1, f) != 1) return 2;
  for (int c = 0; c < cases; ++c) {
    int n = 0;
    double K[9], d[5];
    if (fread(&n, 4, 1, f) != 1 || fread(K, 8, 9, f) != 9 || fread(d, 8, 5, f) != 5) return 2;
    std::vector<double> obj(3 * (size_t)n);
    std::vector<float> pix(2 * (size_t)n);
    if (fread(obj.data(), 8, obj.size(), f) != obj.size() || fread(pix.data(), 4, pix.size(), f) != pix.size()) return 2;
    for (int sc = 0; sc < 8; ++sc) {
      const double sg[3] = {sc & 1 ? -1.0 : 1.0, sc & 2 ? -1.0 : 1.0, sc & 4 ? -1.0 : 1.0};
      double out[13] = {0};
      out[0] = ptzcalib::epnp::solve_pnp_epnp(n, obj.data(), pix.data(), K, d, out + 1, out + 10, sg) ? 1.0 : 0.0;
      fwrite(out, 8, 13, g);
    }
  }
  fclose(f);
  fclose(g);
  return 0;
}
