// ecb_order.h -- the order in which the hit children of a BVH4 node are visited.
//
// Restates traverseClosestHit (kernels/bvh/bvh_traverser1.h:436-506) with the sort networks of
// kernels/common/stack_item.h:39-60: the hit children are pushed in index order {child, tNear bits}, the last two / three /
// four stack items are sorted by `if (b.dist < a.dist) xchg(b, a)` steps on the unsigned distance bits, and the top item is
// popped again as the node to descend into.  Here nothing is pushed that would be popped at once, and the hit children are
// picked by selects on the hit mask instead of a compaction loop with a dynamic index (which lived in local memory) --
// the comparisons, their order and so the behaviour on equal distances are the reference's.
// Host-compilable: tests/cpp/order_check.cpp holds it against a literal push-and-sort restatement for every hit mask and
// every pattern of equal distances.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define ECB_ORDER_FN __host__ __device__ __forceinline__
#else
#define ECB_ORDER_FN inline
#endif

ECB_ORDER_FN uint32_t ecb_order_bits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#endif
}

// if (b.dist < a.dist) xchg(b,a)   (stack_item.h:40-60)
ECB_ORDER_FN void ecb_order_cswap(uint32_t& ra, uint32_t& da, uint32_t& rb, uint32_t& db) {
  if (db < da) {
    const uint32_t tr = ra, td = da;
    ra = rb; da = db;
    rb = tr; db = td;
  }
}

ECB_ORDER_FN int ecb_order_popc(unsigned m) {
#if defined(__CUDA_ARCH__)
  return __popc(m);
#else
  return __builtin_popcount(m);
#endif
}

// mask != 0: bit r set = child r is hit at distance tNear[r].  Pushes all hit children but the one to visit first onto
// {stack_ref, stack_dist} at sp (the farthest deepest) and returns that one.
ECB_ORDER_FN uint32_t ecb_order_hit_children(unsigned mask, const uint32_t child[4], const float tNear[4], uint32_t* stack_ref,
                                             uint32_t* stack_dist, int& sp) {
  if ((mask & (mask - 1u)) == 0u)  // one child hit (the common case): no compaction, no distances
    return (mask & 1u) ? child[0] : (mask & 2u) ? child[1] : (mask & 4u) ? child[2] : child[3];
  if (ecb_order_popc(mask) == 2) {  // two: the lower-index one (a) and the higher-index one (b), picked without the compaction loop
    const bool b0 = mask & 1u, b1 = mask & 2u, b2 = mask & 4u, b3 = mask & 8u;
    const uint32_t ca = b0 ? child[0] : b1 ? child[1] : child[2], da = ecb_order_bits(b0 ? tNear[0] : b1 ? tNear[1] : tNear[2]);
    const uint32_t cb = b3 ? child[3] : b2 ? child[2] : child[1], db = ecb_order_bits(b3 ? tNear[3] : b2 ? tNear[2] : tNear[1]);
    const bool a_first = da < db;  // bvh_traverser1.h:452-456: the farther one is pushed, ties keep the higher index on top
    stack_ref[sp] = a_first ? cb : ca;
    stack_dist[sp] = a_first ? db : da;
    sp++;
    return a_first ? ca : cb;
  }
  if (mask != 15u) {  // three: the hit children in index order are {0 or 1, 1 or 2, 2 or 3}
    const bool b0 = mask & 1u, b01 = (mask & 3u) == 3u, b3 = mask & 8u;
    uint32_t r3 = b0 ? child[0] : child[1], d3 = ecb_order_bits(b0 ? tNear[0] : tNear[1]);
    uint32_t r2 = b01 ? child[1] : child[2], d2 = ecb_order_bits(b01 ? tNear[1] : tNear[2]);
    uint32_t r1 = b3 ? child[3] : child[2], d1 = ecb_order_bits(b3 ? tNear[3] : tNear[2]);
    // items pushed in order c0,c1,c2 ; sort(s1=top(c2), s2=c1, s3=c0)  (stack_item.h:45-50)
    ecb_order_cswap(r1, d1, r2, d2);
    ecb_order_cswap(r2, d2, r3, d3);
    ecb_order_cswap(r1, d1, r2, d2);
    stack_ref[sp] = r3; stack_dist[sp] = d3; sp++;
    stack_ref[sp] = r2; stack_dist[sp] = d2; sp++;
    return r1;
  }
  // all four (stack_item.h:53-60)
  uint32_t r1 = child[3], d1 = ecb_order_bits(tNear[3]), r2 = child[2], d2 = ecb_order_bits(tNear[2]);
  uint32_t r3 = child[1], d3 = ecb_order_bits(tNear[1]), r4 = child[0], d4 = ecb_order_bits(tNear[0]);
  ecb_order_cswap(r1, d1, r2, d2);
  ecb_order_cswap(r3, d3, r4, d4);
  ecb_order_cswap(r1, d1, r3, d3);
  ecb_order_cswap(r2, d2, r4, d4);
  ecb_order_cswap(r2, d2, r3, d3);
  stack_ref[sp] = r4; stack_dist[sp] = d4; sp++;
  stack_ref[sp] = r3; stack_dist[sp] = d3; sp++;
  stack_ref[sp] = r2; stack_dist[sp] = d2; sp++;
  return r1;
}
