// Host-logic check (CPU only): trg_b200::json_number (what TRG::saveGraph writes for a float) against
// nlohmann::json::dump — the reference's writer (trg.cpp:130-177) — on special values, powers of ten around the
// fixed / scientific switch points and a few hundred thousand random floats.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>

#include <nlohmann/json.hpp>

namespace trg_b200 { std::string json_number(float f); }

static int fails = 0;
static long same = 0, near = 0;
static void check(float f) {
  if (!std::isfinite(f)) return;
  const std::string want = nlohmann::json((double)f).dump();
  const std::string got = trg_b200::json_number(f);
  if (want == got) { ++same; return; }
  // nlohmann's Grisu2 is not always shortest / correctly rounded in the 17th digit (2 % of random floats): the two
  // strings must then still use the same notation, parse to the same double and differ only behind 13 digits
  const bool same_notation = (want.find('e') == std::string::npos) == (got.find('e') == std::string::npos);
  const bool same_value = std::strtod(want.c_str(), nullptr) == std::strtod(got.c_str(), nullptr) &&
                          std::strtod(got.c_str(), nullptr) == (double)f;
  size_t common = 0;
  while (common < want.size() && common < got.size() && want[common] == got[common]) ++common;
  if (!same_notation || !same_value || common < 13) {
    if (fails < 10) std::printf("FAIL %.9g: nlohmann '%s' here '%s'\n", (double)f, want.c_str(), got.c_str());
    ++fails;
  }
  ++near;
}

int main() {
  const float special[] = {0.f, -0.f, 1.f, -1.f, 0.5f, 0.1f, 1e-4f, 9.9999e-5f, 1e-5f, 1e5f, 1e15f, 9.99e14f, 1e16f, 123456.f, 100000.f,
                           0.0001f, 0.00012345f, 3.4e38f, 1.17549435e-38f, 1e-45f, 16777216.f, 0.3f, 2.5f, 1234.5678f, 1e7f, 1e-3f};
  for (float f : special) { check(f); check(-f); }
  for (int e = -45; e <= 38; ++e) {
    const float p = std::pow(10.0f, (float)e);
    check(p); check(std::nextafter(p, 0.f)); check(std::nextafter(p, INFINITY)); check(1.5f * p); check(p / 3.f);
  }
  std::mt19937 gen(7);
  for (int i = 0; i < 300000; ++i) {
    uint32_t bits = gen();
    float f;
    std::memcpy(&f, &bits, 4);
    check(f);
  }
  std::uniform_real_distribution<float> U(-400.f, 400.f), W(0.f, 1.f);
  for (int i = 0; i < 200000; ++i) { check(U(gen)); check(W(gen)); }
  std::printf("%s (%d failures; %ld identical strings, %ld equal values with another last digit)\n", fails ? "FAILED" : "ok", fails, same, near);
  return fails ? 1 : 0;
}
