// tests/cpp/order_check.cpp -- host check of csrc/ecb_order.h (the order in which the hit children of a BVH4 node are visited)
// against a literal restatement of the reference: push the hit children in index order, sort the last two / three / four
// stack items with the networks of kernels/common/stack_item.h:39-60, pop the top (kernels/bvh/bvh_traverser1.h:436-506).
// Every hit mask x every assignment of distances from a small set (so every pattern of equal distances occurs), plus the
// bit patterns whose unsigned order differs from their float order (-0.0f, negative values) -- the networks compare the BITS.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../embree_compressed_b200/csrc/ecb_order.h"

struct Item { uint32_t ref, dist; };
static void xchg(Item& a, Item& b) { std::swap(a, b); }
static uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

// the reference, literally: returns cur, leaves the pushed items on `st`
static uint32_t reference_order(unsigned mask, const uint32_t child[4], const float tNear[4], std::vector<Item>& st) {
  int c[4], nh = 0;
  for (int r = 0; r < 4; r++)
    if (mask & (1u << r)) c[nh++] = r;
  if (nh == 1) return child[c[0]];
  const uint32_t d0 = bits(tNear[c[0]]), d1 = bits(tNear[c[1]]);
  if (nh == 2) {
    if (d0 < d1) { st.push_back({child[c[1]], d1}); return child[c[0]]; }
    st.push_back({child[c[0]], d0});
    return child[c[1]];
  }
  st.push_back({child[c[0]], d0});
  st.push_back({child[c[1]], d1});
  st.push_back({child[c[2]], bits(tNear[c[2]])});
  size_t sp = st.size();
  if (nh == 3) {
    Item &s1 = st[sp - 1], &s2 = st[sp - 2], &s3 = st[sp - 3];
    if (s2.dist < s1.dist) xchg(s2, s1);
    if (s3.dist < s2.dist) xchg(s3, s2);
    if (s2.dist < s1.dist) xchg(s2, s1);
  } else {
    st.push_back({child[c[3]], bits(tNear[c[3]])});
    sp = st.size();
    Item &s1 = st[sp - 1], &s2 = st[sp - 2], &s3 = st[sp - 3], &s4 = st[sp - 4];
    if (s2.dist < s1.dist) xchg(s2, s1);
    if (s4.dist < s3.dist) xchg(s4, s3);
    if (s3.dist < s1.dist) xchg(s3, s1);
    if (s4.dist < s2.dist) xchg(s4, s2);
    if (s3.dist < s2.dist) xchg(s3, s2);
  }
  const uint32_t cur = st.back().ref;
  st.pop_back();
  return cur;
}

int main() {
  const float values[] = {0.0f, 0.5f, 1.0f, 2.0f, -0.0f, -1.0f, 3.4e38f};
  const int nv = (int)(sizeof(values) / sizeof(values[0]));
  const uint32_t child[4] = {0x100u, 0x201u, 0x302u, 0x403u};
  long cases = 0, bad = 0;
  for (unsigned mask = 1; mask < 16; mask++)
    for (int a = 0; a < nv; a++)
      for (int b = 0; b < nv; b++)
        for (int c = 0; c < nv; c++)
          for (int d = 0; d < nv; d++) {
            const float tNear[4] = {values[a], values[b], values[c], values[d]};
            std::vector<Item> want_stack = {{0xAAAAu, 7u}};  // something already on the stack stays untouched
            const uint32_t want = reference_order(mask, child, tNear, want_stack);
            uint32_t ref[8] = {0xAAAAu, 0, 0, 0, 0, 0, 0, 0}, dist[8] = {7u, 0, 0, 0, 0, 0, 0, 0};
            int sp = 1;
            const uint32_t got = ecb_order_hit_children(mask, child, tNear, ref, dist, sp);
            cases++;
            bool ok = got == want && sp == (int)want_stack.size();
            for (int k = 0; ok && k < sp; k++) ok = ref[k] == want_stack[k].ref && dist[k] == want_stack[k].dist;
            if (!ok && bad++ < 5) fprintf(stderr, "mismatch: mask %u tNear %g %g %g %g: cur %x (want %x), sp %d (want %zu)\n", mask, tNear[0], tNear[1],
                                          tNear[2], tNear[3], got, want, sp, want_stack.size());
          }
  printf("%ld cases, %ld mismatches\n", cases, bad);
  return bad ? 1 : 0;
}
