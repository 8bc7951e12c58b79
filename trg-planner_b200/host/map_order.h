// Iteration order of a libstdc++ std::unordered_map filled with sequential integer keys — what TRG::cleanGraph
// (trg.cpp:491-535) exposes twice: the survivors get their new ids in the iteration order of the map the build
// filled with ids 0 .. n-1 (:497-504), and node_tree is rebuilt in the iteration order of the renumbered map
// (:528-530). Computing the order instead of walking half a million hash nodes is what lets the device build
// (trg_device_build.cpp) materialise the cleaned graph in parallel. Checked against the real container by
// tests/host/map_order_check.cpp.
#pragma once
#include <algorithm>
#include <cstddef>
#include <unordered_map>
#include <vector>

namespace trg_b200 {

// Iteration order of a std::unordered_map<int, T> (libstdc++, identity hash, max load factor 1) into
// which the keys 0 .. n-1 are inserted in ascending order, starting from `bucket_count` buckets
// (1 = never used; otherwise what clear() left behind). A key whose bucket is empty becomes the new
// head of the element list (hashtable.h _M_insert_bucket_begin), and keys below the bucket count all
// have their own bucket; a rehash re-links the elements in list order, each again at the head
// (_M_rehash_aux), i.e. reverses the list. Rehash points come from libstdc++'s own policy object.
// Keys >= bucket count cannot occur before a rehash (the policy grows first), so no chain is shared.
inline std::vector<int> sequential_map_order(size_t n, size_t bucket_count) {
  std::__detail::_Prime_rehash_policy pol(1.0f);
  size_t bkt = bucket_count;
  pol._M_next_resize = bkt <= 1 ? 0 : (size_t)__builtin_floor((double)bkt * 1.0);
  // deque with a direction flag: push at the logical front, reverse = flip
  std::vector<int> buf(2 * n + 2);
  size_t lo = n + 1, hi = n + 1;  // elements in [lo, hi)
  bool flipped = false;           // logical front is at hi when flipped
  for (size_t k = 0; k < n; ++k) {
    const auto r = pol._M_need_rehash(bkt, k, 1);
    if (r.first) {
      bkt = r.second;
      flipped = !flipped;
    }
    if (!flipped) buf[--lo] = (int)k; else buf[hi++] = (int)k;
  }
  std::vector<int> out(n);
  if (!flipped) std::copy(buf.begin() + lo, buf.begin() + hi, out.begin());
  else std::reverse_copy(buf.begin() + lo, buf.begin() + hi, out.begin());
  return out;
}

}  // namespace trg_b200
