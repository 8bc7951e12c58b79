// Host-logic check (CPU only): trg_b200::sequential_map_order against the real libstdc++ container, for fresh maps
// (bucket count 1) and for maps that were used and clear()ed before (clear keeps the bucket array — the state
// TRG::initGraph finds when a graph is rebuilt), sizes around every rehash point up to a few hundred thousand keys.
#include <cstdio>
#include <unordered_map>
#include <vector>

#include "map_order.h"

static int fails = 0;

static void check(size_t n, size_t prefill) {
  std::unordered_map<int, int> m;
  for (size_t k = 0; k < prefill; ++k) m[(int)k] = 0;
  m.clear();
  const size_t buckets = m.bucket_count();
  for (size_t k = 0; k < n; ++k) m[(int)k] = (int)k;
  const std::vector<int> want = trg_b200::sequential_map_order(n, buckets);
  size_t i = 0;
  bool ok = want.size() == m.size();
  for (auto& kv : m) {
    if (!ok) break;
    ok = want[i++] == kv.first;
  }
  if (!ok) {
    if (fails < 10) std::printf("FAIL n=%zu prefill=%zu buckets=%zu\n", n, prefill, buckets);
    ++fails;
  }
}

int main() {
  const size_t sizes[] = {0, 1, 2, 12, 13, 14, 29, 30, 59, 60, 127, 128, 541, 542, 1109, 5000, 62233, 62234, 130000, 544771};
  const size_t prefills[] = {0, 1, 13, 100, 5000, 70000, 600000};
  for (size_t n : sizes)
    for (size_t p : prefills) check(n, p);
  for (size_t n = 0; n < 3000; ++n) check(n, 0);
  std::printf("%s (%d failures)\n", fails ? "FAILED" : "ok", fails);
  return fails ? 1 : 0;
}
