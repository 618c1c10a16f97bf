// Host-side check of the index algebra of csrc/fft_regs.cuh against a naive DFT (no GPU needed):
//   nvcc -I cheetah_b200/csrc -o /tmp/fft_regs_test tests/native/fft_regs_test.cu && /tmp/fft_regs_test
#include <cmath>
#include <complex>
#include <cstdio>
#include <vector>

#include "fft_regs.cuh"

using ch::fftr::C;

template <int LEN, bool INV, bool LAYOUT_B = false>
double check() {
  constexpr int N2 = ch::fftr::Plan<LEN>::N2;
  std::vector<C> tw(LEN), column(LAYOUT_B ? ch::fftr::PlanB<LEN>::PITCH : ch::fftr::Plan<LEN>::PITCH);
  for (int k = 0; k < LEN; ++k)
    tw[k] = C{static_cast<float>(std::cos(-2.0 * M_PI * k / LEN)),
              static_cast<float>(std::sin(-2.0 * M_PI * k / LEN))};
  std::vector<std::complex<double>> x(LEN), truth(LEN);
  unsigned state = 12345u + LEN;
  auto rnd = [&] { state = state * 1664525u + 1013904223u; return (state >> 8) / 16777216.0 - 0.5; };
  for (auto& v : x) v = {rnd(), rnd()};
  for (int k = 0; k < LEN; ++k) {
    std::complex<double> acc = 0;
    for (int n = 0; n < LEN; ++n)
      acc += x[n] * std::polar(1.0, (INV ? 2.0 : -2.0) * M_PI * n * k / LEN);
    truth[k] = acc;
  }
  std::vector<std::vector<C>> regs(N2, std::vector<C>(16));
  for (int n2 = 0; n2 < N2; ++n2) {
    C v[16];
    for (int n1 = 0; n1 < 16; ++n1)
      v[n1] = C{static_cast<float>(x[N2 * n1 + n2].real()), static_cast<float>(x[N2 * n1 + n2].imag())};
    if (LAYOUT_B) ch::fftr::transform_scatter_b<LEN, INV>(v, column.data(), n2, tw.data());
    else ch::fftr::transform_scatter<LEN, INV>(v, column.data(), n2, tw.data());
  }
  double worst = 0, scale = 0;
  for (int n2 = 0; n2 < N2; ++n2) {
    C v[16];
    if (LAYOUT_B) ch::fftr::transform_gather_b<LEN, INV>(v, column.data(), n2);
    else ch::fftr::transform_gather<LEN, INV>(v, column.data(), n2);
    for (int j = 0; j < 16; ++j) {
      const auto t = truth[n2 + N2 * j];
      worst = std::fmax(worst, std::abs(std::complex<double>(v[j].x, v[j].y) - t));
      scale = std::fmax(scale, std::abs(t));
    }
  }
  return worst / scale;
}

int main() {
  double errs[] = {check<32, false>(),  check<32, true>(),  check<64, false>(),  check<64, true>(),
                   check<128, false>(), check<128, true>(), check<256, false>(), check<256, true>(),
                   check<32, false, true>(), check<64, true, true>(), check<128, false, true>(),
                   check<128, true, true>(), check<256, false, true>()};
  int bad = 0;
  for (double e : errs) {
    std::printf("%.3e\n", e);
    bad += !(e < 5e-6);
  }
  std::printf(bad ? "FAILED\n" : "ok\n");
  return bad;
}
