// Host build of mellon_b200/csrc/mb_math.cuh (the lean exp / sqrt of the K1 / K7 epilogue): accuracy against libm.
// Compiled and run by tests/test_mb_math_host.py; prints "max_rel_exp max_rel_sqrt exp0 exp709 clamp_neg clamp_zero clamp_pos".
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>

#include "../../mellon_b200/csrc/mb_math.cuh"

int main() {
  static const double tab[64] = MB_EXP2_TABLE_INIT;
  std::mt19937_64 g(1);
  double maxe = 0, maxs = 0;
  for (int i = 0; i < 4000000; i++) {
    const double u = (g() >> 11) * (1.0 / 9007199254740992.0);
    const double r = (i % 3 == 0) ? u * 700 : (i % 3 == 1 ? u * 40 : u * 1e-3);
    const double a = mbmath::exp_neg(r, tab), b = std::exp(-r);
    maxe = std::fmax(maxe, std::fabs(a - b) / b);
    const double s = std::ldexp(1.0 + u, (int)(g() % 1200) - 600);
    const double q = mbmath::sqrt_pos(s), q0 = std::sqrt(s);
    maxs = std::fmax(maxs, std::fabs(q - q0) / q0);
  }
  std::printf("%.6e %.6e %.17g %.17g %.17g %.17g %.17g\n", maxe, maxs, mbmath::exp_neg(0.0, tab), mbmath::exp_neg(709.0, tab),
              mbmath::clamp_tiny(-1.0), mbmath::clamp_tiny(0.0), mbmath::clamp_tiny(2e-300));
  return 0;
}
