// ecb_kernels.cu -- sm_100a traversal kernels for the compressed subdivision/displacement BVH.
//
// Per-ray semantics restate (paths relative to the reference tree):
//   top level   BVHNIntersector1<4,BVH_AN1,robust>::intersect / occluded   kernels/bvh/bvh_intersector1.cpp:40-209
//               TravRay / intersectNodeRobust                              kernels/bvh/node_intersector1.h:113-139,334-368
//               traverseClosestHit / traverseAnyHit + sort networks        kernels/bvh/bvh_traverser1.h:436-537, kernels/common/stack_item.h:39-60
//   block       CompressedBVHIntersector1::intersect                       kernels/geometry/compressed.h:454-752
//               Node::getNode + tables                                     kernels/geometry/compressed_node.h:22-39,261-295,488-510
//               leafStorage<4>::getZ                                       kernels/geometry/compressed_leaf.h:93-111
//               project / frustum / patch / triangles / Morton             kernels/geometry/compressed_help.h:19-50,86-308
//
// Arithmetic contract: fp32, compiled with -fmad=false; the only fused multiply-adds are the explicit __fmaf_rn in
// xfm() where the reference's AVX2 build has explicit FMAs (common/math/linearspace3.h:168-169).  Division and sqrt
// are IEEE (the reference uses rcpss/rsqrtss + one Newton step, common/math/math.h:60-97; results agree to ~1 ulp).
// With this contract the kernels are bit-identical to the scalar oracle (oracle/cbvh_port.c).
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "ecb_format.h"
#include "ecb_kernels.h"
#include "ecb_order.h"

// ---------------------------------------------------------------------------------------------------------------
// -DECB_BOUNDS_CHECK: every index the kernels derive from scene data (node / block / item / instance references, stack
// positions, node-word and leaf offsets) is checked against the array it indexes; a violation is counted per class in
// g_ecb_check[] (read back with ecbBoundsCheckFailures).  compute-sanitizer is not available on the B200 pool, so the
// golden and parity tests are run against a library built with this flag instead (tests/test_bounds_check_gpu.py).
// The shipping build compiles the checks away.
// ---------------------------------------------------------------------------------------------------------------
#ifdef ECB_BOUNDS_CHECK
__device__ unsigned long long g_ecb_check[8];
#define ECB_CHECK(cond, slot)                          \
  do {                                                 \
    if (!(cond)) atomicAdd(&g_ecb_check[slot], 1ull);  \
  } while (0)
#else
#define ECB_CHECK(cond, slot) \
  do {                        \
  } while (0)
#endif
// slots: 0 top-level stack overflow, 1 top-level node index, 2 block index, 3 block stack overflow, 4 node word / leaf
// offset inside a block, 5 instance item / instance / object index, 6 triangle block index, 7 ray index

namespace {

// ---------------------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------------------
// _mm_max_ps(a,b) = a > b ? a : b (second operand on NaN / equal zeros).  In the slab tests no operand can be NaN for a
// finite ray ((plane - org) * rdir with |rdir| <= 1e18 and planes finite or +-inf) and the sign of a zero result never
// reaches a bit-compare (tNear >= +0 always, see trace_closest), so the single-instruction IEEE min/max is identical.
__device__ __forceinline__ float sse_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ float sse_min(float a, float b) { return fminf(a, b); }
// exact operand-order form, used where a NaN or a signed zero could make a difference (ray set-up)
__device__ __forceinline__ float sse_max_exact(float a, float b) { return a > b ? a : b; }
__device__ __forceinline__ float zero_fix(float a) { return fabsf(a) < 1E-18f ? 1E-18f : a; }

// Ray records are read with ordinary (coherent) loads, never ld.global.nc: the in-place entry points write hits into the
// very records they read (L.in == L.out), and the non-coherent path requires data that nothing modifies while the kernel runs.
// ld.global.cg: coherent at L2, 128 bits wide, and the records are read exactly once, so bypassing L1 costs nothing.
__device__ __forceinline__ float4 ld_ray4(const uint8_t* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float ld_ray1(const uint8_t* p) { return __ldcg(reinterpret_cast<const float*>(p)); }

__device__ __forceinline__ uint32_t compact1by1(uint32_t x) {
  x &= 0x55555555u;
  x = (x ^ (x >> 1)) & 0x33333333u;
  x = (x ^ (x >> 2)) & 0x0f0f0f0fu;
  x = (x ^ (x >> 4)) & 0x00ff00ffu;
  x = (x ^ (x >> 8)) & 0x0000ffffu;
  return x;
}

// quantization tables (compressed_node.h:22-39): [0] border, [1] mid, [2] 1-mid, [3] 1-border
__device__ __forceinline__ void init_lut(float* lut) {
  const float border[8] = {0.000f, 0.005f, 0.010f, 0.050f, 0.100f, 0.200f, 0.400f, 0.600f};
  const float mid[8] = {0.00f, 0.40f, 0.48f, 0.49f, 0.50f, 0.51f, 0.52f, 0.60f};
  const int t = threadIdx.x;
  if (t < 8) {
    lut[t] = border[t];
    lut[8 + t] = mid[t];
    lut[16 + t] = 1.f - mid[t];
    lut[24 + t] = 1.f - border[t];
  }
}

// ray i of a launch: from the AoS records (RTCRayHit / RTCRay / packed) or from the eight SoA arrays (EcbLaunch::soa)
__device__ __forceinline__ void load_ray(const EcbLaunch& L, uint64_t i, float4& r0, float4& r1, float& tfar) {
  if (L.soa[0]) {
    r0 = make_float4(__ldcg(L.soa[0] + i), __ldcg(L.soa[1] + i), __ldcg(L.soa[2] + i), __ldcg(L.soa[3] + i));
    r1 = make_float4(__ldcg(L.soa[4] + i), __ldcg(L.soa[5] + i), __ldcg(L.soa[6] + i), 0.f);
    tfar = __ldcg(L.soa[7] + i);
  } else {
    const uint8_t* rp = L.in + i * L.in_stride;
    r0 = ld_ray4(rp);       // org.xyz, tnear
    r1 = ld_ray4(rp + 16);  // dir.xyz, time
    tfar = ld_ray1(rp + L.tfar_offset);  // 32, or 28 in packed records
  }
}

struct TravRay {
  float org[3], rdn[3], rdf[3];
  float tnear, tfar;
  int nu[3];  // 1 if the near plane along that axis is the upper bound
};

__device__ __forceinline__ void travray_init(TravRay& r, const float org[3], const float dir[3], float tnear, float tfar) {
  const float ulp3 = 1.0f + 3.0f * FLT_EPSILON;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    r.org[k] = org[k];
    r.rdn[k] = 1.0f / zero_fix(dir[k]);
    r.rdf[k] = r.rdn[k] * ulp3;
    r.nu[k] = r.rdn[k] >= 0.0f ? 0 : 1;
  }
  r.tnear = tnear;
  r.tfar = tfar;
}

__device__ __forceinline__ void project3(const float* m /*row-major 3x3*/, const float a[3], float out[3]) {
  const float X = m[0] * a[0] + m[1] * a[1] + m[2] * 1.0f;
  const float Y = m[3] * a[0] + m[4] * a[1] + m[5] * 1.0f;
  const float W = m[6] * a[0] + m[7] * a[1] + m[8] * 1.0f;
  out[0] = X / W;
  out[1] = Y / W;
  out[2] = a[2];
}

__device__ __noinline__ void inverse3(const float* a, float* r) {
  const float c00 = a[4] * a[8] - a[5] * a[7];
  const float c10 = a[2] * a[7] - a[1] * a[8];
  const float c20 = a[1] * a[5] - a[2] * a[4];
  const float det = c00 * a[0] + c10 * a[3] + c20 * a[6];
  const float id = 1.0f / det;
  r[0] = c00 * id;
  r[1] = c10 * id;
  r[2] = c20 * id;
  r[3] = (a[5] * a[6] - a[3] * a[8]) * id;
  r[4] = (a[0] * a[8] - a[2] * a[6]) * id;
  r[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  r[6] = (a[3] * a[7] - a[4] * a[6]) * id;
  r[7] = (a[1] * a[6] - a[0] * a[7]) * id;
  r[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}

__device__ __forceinline__ float intersect_line(float p2x, float p2y, float p3x, float p3y, const float o[3], const float d[3]) {
  const float vx = p2x - o[0], vy = p2y - o[1];
  const float lx = p3x - p2x, ly = p3y - p2y;
  const float t1 = (ly * vx - lx * vy) / (ly * d[0] - lx * d[1]);
  const float t2 = (d[0] * vy - d[1] * vx) / (lx * d[1] - ly * d[0]);
  return (t2 >= 0.f && t2 <= 1.f) ? t1 : __int_as_float(0x7fc00000);
}

__device__ __forceinline__ bool intersect_frustum(const float* box, const float o[3], const float d[3], float& t_near, float& t_far) {
  const float rdz = 1.0f / zero_fix(d[2]);
  const float org_rdz = o[2] * rdz;
  const float t1z = box[0] * rdz - org_rdz;
  const float t2z = box[1] * rdz - org_rdz;
  const float t1x = intersect_line(box[2], box[3], box[6], box[7], o, d);
  const float t2x = intersect_line(box[4], box[5], box[8], box[9], o, d);
  const float t1y = intersect_line(box[2], box[3], box[4], box[5], o, d);
  const float t2y = intersect_line(box[6], box[7], box[8], box[9], o, d);
  const float near1 = fminf(fminf(t1x, t2x), fminf(t1y, t2y));
  const float far1 = fmaxf(fmaxf(t1x, t2x), fmaxf(t1y, t2y));
  t_near = fmaxf(fmaxf(fminf(t1z, t2z), near1), t_near);
  t_far = fminf(fminf(fmaxf(t1z, t2z), far1), t_far);
  return t_near <= t_far && near1 == near1 && far1 == far1;
}

__device__ __forceinline__ bool intersect_patch(uint32_t idx, float rcp_edges, float dz, float t1, float t2, float v0, float v1,
                                                float v2, float v3, const float box[6], const float org[3], const float dir[3],
                                                float& u, float& v, float& tt) {
  const float px = t1 * dir[0] + org[0], py = t1 * dir[1] + org[1], pz = t1 * dir[2] + org[2];
  const float p2x = t2 * dir[0] + org[0], p2y = t2 * dir[1] + org[1], p2z = t2 * dir[2] + org[2];
  const float lenX = 1.0f / (box[3] - box[0]);
  const float lenY = 1.0f / (box[4] - box[1]);
  const float fx1 = (px - box[0]) * lenX;
  const float fy1 = (py - box[1]) * lenY;
  const float mx = (float)compact1by1(idx), my = (float)compact1by1(idx >> 1);
  if (t2 - t1 < 1.0E-6f) {
    tt = t1;
    u = (fx1 + mx) * rcp_edges;
    v = (fy1 + my) * rcp_edges;
    return true;
  }
  const float fx2 = (p2x - box[0]) * lenX;
  const float fy2 = (p2y - box[1]) * lenY;
  const float dx1 = 1.f - fx1, dy1 = 1.f - fy1;
  float z1 = v0 * dx1 * dy1 + v1 * fx1 * dy1 + v2 * dx1 * fy1 + v3 * fx1 * fy1;
  const float dx2 = 1.f - fx2, dy2 = 1.f - fy2;
  float z2 = v0 * dx2 * dy2 + v1 * fx2 * dy2 + v2 * dx2 * fy2 + v3 * fx2 * fy2;
  if (pz >= z1 && pz <= z1 + dz) {
    tt = t1;
    u = (fx1 + mx) * rcp_edges;
    v = (fy1 + my) * rcp_edges;
    return true;
  }
  if (pz > z1 + dz) {
    z1 += dz;
    z2 += dz;
  }
  const float alpha = p2z - z2;
  const float beta = z1 - pz;
  const float t = (t1 * alpha + t2 * beta) / (alpha + beta);
  const float d = (t - t1) / (t2 - t1);
  const float fx = fx2 - fx1, fy = fy2 - fy1;
  if (t < tt && t >= t1 && t <= t2) {
    u = (fx * d + fx1 + mx) * rcp_edges;
    v = (fy * d + fy1 + my) * rcp_edges;
    tt = t;
    return true;
  }
  return false;
}

__device__ __forceinline__ bool intersect_triangle(const float* v0, const float* v1, const float* v2, const float rorg[3],
                                                   const float rdir[3], float tnear, float& tfar, float& ou, float& ov) {
  const float a = v0[0] - v1[0], b = v0[1] - v1[1], c = v0[2] - v1[2];
  const float d = v0[0] - v2[0], e = v0[1] - v2[1], f = v0[2] - v2[2];
  const float g = rdir[0], h = rdir[1], i = rdir[2];
  const float j = v0[0] - rorg[0], k = v0[1] - rorg[1], l = v0[2] - rorg[2];
  float common1 = e * i - h * f;
  float common2 = g * f - d * i;
  float common3 = d * h - e * g;
  const float M = 1.0f / (a * common1 + b * common2 + c * common3);
  float beta = j * common1 + k * common2 + l * common3;
  common1 = a * k - j * b;
  common2 = j * c - a * l;
  common3 = b * l - k * c;
  float gamma = i * common1 + h * common2 + g * common3;
  float tt = -(f * common1 + e * common2 + d * common3);
  beta *= M;
  gamma *= M;
  tt *= M;
  if (tt > 0 && tt < tfar && tt >= tnear)
    if (beta > 0 && gamma > 0 && beta + gamma <= 1) {
      tfar = tt;
      ou = beta;
      ov = gamma;
      return true;
    }
  return false;
}

struct Hit {
  float tfar, u, v;
  uint32_t primID, geomID, block;
};

// per-thread work counters of the instrumented (STATS) kernel variant: SURVEY 8d's N_top, N_blk, N_cnode, N_cell
struct Work {
  uint32_t top_nodes, blocks, cnodes, cells;
};

// ---------------------------------------------------------------------------------------------------------------
// block traversal (CompressedBVHIntersector1::intersect)
// ---------------------------------------------------------------------------------------------------------------
// INLINE_LEAVES: the leaves below a lowest-level quadtree node are tested right at that node, in pop order, instead of being
// pushed and popped one by one.  Same results bit for bit; it pays on deep blocks with many cell tests per entry (displaced
// cube S6/C4: 15.5 node decodes + 9.1 cells per ray, 0.65 -> 0.58 ms) and costs on shallow ones (bomberman S6/C3: 2.6 + 0.9,
// 0.389 -> 0.409 ms), so the launcher picks it for C >= 4 (profiles/r2r_leaves_inline_ab.txt).
template <int LEAF, bool STATS, bool INLINE_LEAVES = false>
__device__ __forceinline__ void block_intersect(const EcbDeviceScene& sc, const float* __restrict__ lut, uint32_t block_index,
                                                const float rorg[3], const float rdir[3], float ray_tnear, Hit& hit, Work& work) {
  if (STATS) work.blocks++;
  ECB_CHECK(block_index < sc.n_blocks, 2);
  const uint8_t* blk = sc.blocks + (size_t)block_index * sc.block_stride;
  const float4* h4 = reinterpret_cast<const float4*>(blk);
  // 144-byte header = 9 x 16-byte read-only loads
  float hdr[36];
#pragma unroll
  for (int i = 0; i < 9; i++) {
    const float4 q = __ldg(h4 + i);
    hdr[4 * i + 0] = q.x; hdr[4 * i + 1] = q.y; hdr[4 * i + 2] = q.z; hdr[4 * i + 3] = q.w;
  }
  const float* space = hdr;       // [3*c+k]
  const float* proj = hdr + 9;    // row-major
  const float* box = hdr + 18;
  const float uv0x = hdr[28], uv0y = hdr[29], uvEx = hdr[30], uvEy = hdr[31];
  const float extent = hdr[32];
  const uint32_t geomID = __float_as_uint(hdr[33]), primID = __float_as_uint(hdr[34]), root_word = __float_as_uint(hdr[35]);

  float lOrg[3], lDir[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    lOrg[k] = __fmaf_rn(rorg[0], space[k], __fmaf_rn(rorg[1], space[3 + k], rorg[2] * space[6 + k]));
    lDir[k] = __fmaf_rn(rdir[0], space[k], __fmaf_rn(rdir[1], space[3 + k], rdir[2] * space[6 + k]));
  }
  float near = ray_tnear;
  float far = hit.tfar;
  if (!intersect_frustum(box, lOrg, lDir, near, far)) return;

  float pin[3], pout[3], org[3], tar[3], dir[3];
#pragma unroll
  for (int k = 0; k < 3; k++) pin[k] = lOrg[k] + lDir[k] * near;
#pragma unroll
  for (int k = 0; k < 3; k++) pout[k] = lOrg[k] + lDir[k] * far;
  project3(proj, pin, org);
  project3(proj, pout, tar);
#pragma unroll
  for (int k = 0; k < 3; k++) dir[k] = tar[k] - org[k];
  float zFactor = 0.f;
  bool special = false;
  float tfar = 0.f;
  const float g_epsilon = 1.0E-4f;
  const float ax = fabsf(dir[0]), ay = fabsf(dir[1]), az = fabsf(dir[2]);
  if (ax < g_epsilon && ay < g_epsilon && az < g_epsilon) {
    dir[2] = copysignf(1.f, lDir[2]);
    org[2] -= dir[2];
    zFactor = FLT_MAX;
    tfar = FLT_MAX;
  } else if (az < g_epsilon) {
    special = true;
    const float d2 = (dir[0] * dir[0] + dir[1] * dir[1]) + (dir[2] * dir[2] + 0.f);
    tfar = sqrtf(d2);
    const float rl = 1.0f / sqrtf(d2);
#pragma unroll
    for (int k = 0; k < 3; k++) dir[k] = dir[k] * rl;
  } else {
    const float d2 = (dir[0] * dir[0] + dir[1] * dir[1]) + (dir[2] * dir[2] + 0.f);
    const float rl = 1.0f / sqrtf(d2);
#pragma unroll
    for (int k = 0; k < 3; k++) dir[k] = dir[k] * rl;
    zFactor = lDir[2] / dir[2];
    tfar = (hit.tfar - near) * zFactor;
  }

  TravRay tr;
  travray_init(tr, org, dir, 0.f, tfar);
  const uint32_t elems = sc.elems;
  const uint32_t* node_words = reinterpret_cast<const uint32_t*>(blk + ECB_BLOCK_NODES_OFFSET);

  if constexpr (INLINE_LEAVES) {
  // ---- quadtree.  The stack holds INNER nodes only (3(C-1)+1 entries of {node, box}): the leaves below a lowest-level
  // node are not pushed and popped one by one but tested right there, in the order the reference pops them (nearest box
  // first, lower child index on ties, compressed.h:668-749) -- nothing else can come between them on the reference's stack.
  // (A lane-interleaved shared-memory stack was measured too: slower on every scene, profiles/r2r_leaves_inline_ab.txt.)
  uint32_t lstack[7 * (3 * (ECB_MAX_C - 1) + 1)];
#define BST(e, w) lstack[7 * (e) + (w)]
  int sp = 0;
  BST(0, 0) = 0u;
  BST(0, 1) = __float_as_uint(-1.f); BST(0, 2) = __float_as_uint(-1.f); BST(0, 3) = __float_as_uint(box[0]);
  BST(0, 4) = __float_as_uint(1.f);  BST(0, 5) = __float_as_uint(1.f);  BST(0, 6) = __float_as_uint(box[1]);

  while (sp >= 0) {
    const unsigned curr = BST(sp, 0);
    float pBox[6];
#pragma unroll
    for (int k = 0; k < 6; k++) pBox[k] = __uint_as_float(BST(sp, 1 + k));
    sp--;

    // inner node: decode (Node::getNode), test 4 children
    const unsigned next = curr * 4 + 1;
    if (STATS) work.cnodes++;
    ECB_CHECK(curr < elems, 4);
    const uint32_t w = curr == 0 ? root_word : __ldg(node_words + (curr - 1));
    const unsigned xz = w & 0xff, xb = (w >> 8) & 0xff, yz = (w >> 16) & 0xff, yb = w >> 24;
    const float dimX = pBox[3] - pBox[0], dimY = pBox[4] - pBox[1], dimZ = pBox[5] - pBox[2];
    const float lx02 = lut[(xz & 0xe0) >> 5] * dimX + pBox[0];
    const float lx13 = lut[8 + ((xz & 0x1c) >> 2)] * dimX + pBox[0];
    const float ux02 = lut[16 + ((xb & 0xe0) >> 5)] * dimX + pBox[0];
    const float ux13 = lut[24 + ((xb & 0x1c) >> 2)] * dimX + pBox[0];
    const float ly01 = lut[(yz & 0xe0) >> 5] * dimY + pBox[1];
    const float ly23 = lut[8 + ((yz & 0x1c) >> 2)] * dimY + pBox[1];
    const float uy01 = lut[16 + ((yb & 0xe0) >> 5)] * dimY + pBox[1];
    const float uy23 = lut[24 + ((yb & 0x1c) >> 2)] * dimY + pBox[1];
    const float lz = (0.25f * (float)(xz & 3)) * dimZ + pBox[2];
    const float uz = (1.f - 0.25f * (float)(yz & 3)) * dimZ + pBox[2];

    // the four children share their planes: x ranges A (children 0,2) / B (1,3), y ranges A (0,1) / B (2,3), one z range
    // (compressed_node.h:488-510) -> 5 near + 5 far plane distances instead of 24; same operands, same floats
    const float tnxA = ((tr.nu[0] ? ux02 : lx02) - tr.org[0]) * tr.rdn[0], tfxA = ((tr.nu[0] ? lx02 : ux02) - tr.org[0]) * tr.rdf[0];
    const float tnxB = ((tr.nu[0] ? ux13 : lx13) - tr.org[0]) * tr.rdn[0], tfxB = ((tr.nu[0] ? lx13 : ux13) - tr.org[0]) * tr.rdf[0];
    const float tnyA = ((tr.nu[1] ? uy01 : ly01) - tr.org[1]) * tr.rdn[1], tfyA = ((tr.nu[1] ? ly01 : uy01) - tr.org[1]) * tr.rdf[1];
    const float tnyB = ((tr.nu[1] ? uy23 : ly23) - tr.org[1]) * tr.rdn[1], tfyB = ((tr.nu[1] ? ly23 : uy23) - tr.org[1]) * tr.rdf[1];
    const float tnz = ((tr.nu[2] ? uz : lz) - tr.org[2]) * tr.rdn[2], tfz = ((tr.nu[2] ? lz : uz) - tr.org[2]) * tr.rdf[2];
    const float tnzn = sse_max(tnz, tr.tnear), tfzf = sse_min(tfz, tr.tfar);
    float tNear[4], tFar[4];
    tNear[0] = sse_max(sse_max(tnxA, tnyA), tnzn); tFar[0] = sse_min(sse_min(tfxA, tfyA), tfzf);
    tNear[1] = sse_max(sse_max(tnxB, tnyA), tnzn); tFar[1] = sse_min(sse_min(tfxB, tfyA), tfzf);
    tNear[2] = sse_max(sse_max(tnxA, tnyB), tnzn); tFar[2] = sse_min(sse_min(tfxA, tfyB), tfzf);
    tNear[3] = sse_max(sse_max(tnxB, tnyB), tnzn); tFar[3] = sse_min(sse_min(tfxB, tfyB), tfzf);
    const unsigned mask = (tNear[0] <= tFar[0] ? 1u : 0u) | (tNear[1] <= tFar[1] ? 2u : 0u) | (tNear[2] <= tFar[2] ? 4u : 0u) |
                          (tNear[3] <= tFar[3] ? 8u : 0u);
    if (mask == 0) continue;

    // below[r] = number of other hit children popped AFTER r (farther, or equal and higher index)   (compressed.h:668-749)
    const int n = __popc(mask);
    int below[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
      below[r] = -1;
      if (!(mask & (1u << r))) continue;
      int cnt = 0;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        if (q == r || !(mask & (1u << q))) continue;
        const bool r_first = (q > r) ? (tNear[r] <= tNear[q]) : (tNear[r] < tNear[q]);
        cnt += r_first ? 1 : 0;
      }
      below[r] = cnt;
    }

    if (next < elems) {  // children are inner nodes: push them, nearest on top
      ECB_CHECK(sp + n < 3 * (int)sc.comp_level - 2, 3);
#pragma unroll
      for (int r = 0; r < 4; r++) {
        if (below[r] < 0) continue;
        const int pos = sp + 1 + below[r];
        BST(pos, 0) = next + r;
        BST(pos, 1) = __float_as_uint((r & 1) ? lx13 : lx02);
        BST(pos, 2) = __float_as_uint((r & 2) ? ly23 : ly01);
        BST(pos, 3) = __float_as_uint(lz);
        BST(pos, 4) = __float_as_uint((r & 1) ? ux13 : ux02);
        BST(pos, 5) = __float_as_uint((r & 2) ? uy23 : uy01);
        BST(pos, 6) = __float_as_uint(uz);
      }
      sp += n;
      continue;
    }

    // ---- children are leaves: test the hit ones in pop order (the one with the most children "below" it first) ----
    for (int k = n - 1; k >= 0; k--) {
      const int r = below[0] == k ? 0 : below[1] == k ? 1 : below[2] == k ? 2 : 3;
      const unsigned idx = next + (unsigned)r - elems;
      ECB_CHECK(idx < 3u * elems + 1u, 4);
      if (STATS) work.cells++;
      const float bl0 = (r & 1) ? lx13 : lx02, bl1 = (r & 2) ? ly23 : ly01, bu0 = (r & 1) ? ux13 : ux02, bu1 = (r & 2) ? uy23 : uy01;
      const float box_near = r == 0 ? tNear[0] : r == 1 ? tNear[1] : r == 2 ? tNear[2] : tNear[3];
      if (LEAF == 1) {
        if (box_near >= tfar) continue;
        const float box_far = r == 0 ? tFar[0] : r == 1 ? tFar[1] : r == 2 ? tFar[2] : tFar[3];
        const float lbox[6] = {bl0, bl1, lz, bu0, bu1, uz};
        const float cdimZ = uz - lz;
        const float range = (1.f + 2.f * extent) * cdimZ;
        const float dz = 0.0625f * range;
        const float offset = lz - cdimZ * extent;
        const float rcpF = 0.0625f * range;
        const uint32_t zz = __ldg(reinterpret_cast<const unsigned short*>(blk + sc.leaf_offset) + idx);
        const unsigned z12 = zz & 0xff, z34 = zz >> 8;
        const float z1 = offset + rcpF * (float)((z12 & 0xf0) >> 4);
        const float z2 = offset + rcpF * (float)(z12 & 0x0f);
        const float z3 = offset + rcpF * (float)((z34 & 0xf0) >> 4);
        const float z4 = offset + rcpF * (float)(z34 & 0x0f);
        float u, v, t = tfar;
        if (intersect_patch(idx, sc.rcp_edges, dz, box_near, box_far, z1, z2, z3, z4, lbox, org, dir, u, v, t)) {
          float tw;
          if (special) {
            float ip[9], q[3], p[3];
            inverse3(proj, ip);
#pragma unroll
            for (int c = 0; c < 3; c++) q[c] = org[c] + dir[c] * t;
            project3(ip, q, p);
            const float ex = p[0] - lOrg[0], ey = p[1] - lOrg[1], ez = p[2] - lOrg[2];
            tw = sqrtf((ex * ex + ey * ey) + (ez * ez + 0.f));
          } else {
            tw = t / zFactor + near;
          }
          hit.u = uv0x + u * uvEx;
          hit.v = uv0y + v * uvEy;
          hit.geomID = geomID;
          hit.primID = primID;
          hit.block = block_index;
          tfar = t;
          hit.tfar = tw;
        }
      } else if (LEAF == 0) {
        const unsigned x = compact1by1(idx), y = compact1by1(idx >> 1);
        const float* grid = reinterpret_cast<const float*>(blk + sc.leaf_offset);
        const unsigned g0 = y * sc.grid_width + x, g2 = g0 + sc.grid_width;
        float V[4][3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
          V[0][c] = __ldg(grid + 3 * g0 + c);
          V[1][c] = __ldg(grid + 3 * g0 + 3 + c);
          V[2][c] = __ldg(grid + 3 * g2 + c);
          V[3][c] = __ldg(grid + 3 * g2 + 3 + c);
        }
        float tf = hit.tfar, uu = 0.f, vv = 0.f;
        const bool hit1 = intersect_triangle(V[0], V[1], V[2], rorg, rdir, ray_tnear, tf, uu, vv);
        const bool hit2 = intersect_triangle(V[3], V[2], V[1], rorg, rdir, ray_tnear, tf, uu, vv);
        if (hit1 || hit2) {
          float lu, lv;
          if (hit2) {
            lu = ((float)x + (1.f - uu)) * sc.rcp_edges;
            lv = ((float)y + (1.f - vv)) * sc.rcp_edges;
          } else {
            lu = ((float)x + uu) * sc.rcp_edges;
            lv = ((float)y + vv) * sc.rcp_edges;
          }
          hit.tfar = tf;
          hit.u = uv0x + lu * uvEx;
          hit.v = uv0y + lv * uvEy;
          hit.geomID = geomID;
          hit.primID = primID;
          hit.block = block_index;
        }
      } else {
        const float is = box_near;
        if (is <= tfar) {
          const float mx = (float)compact1by1(idx), my = (float)compact1by1(idx >> 1);
          const float x = (((org[0] + dir[0] * is) - bl0) / (bu0 - bl0) + mx) * sc.rcp_edges;
          const float y = (((org[1] + dir[1] * is) - bl1) / (bu1 - bl1) + my) * sc.rcp_edges;
          float tw;
          if (special) {
            float ip[9], q[3], p[3];
            inverse3(proj, ip);
#pragma unroll
            for (int c = 0; c < 3; c++) q[c] = org[c] + dir[c] * is;
            project3(ip, q, p);
            const float ex = p[0] - lOrg[0], ey = p[1] - lOrg[1], ez = p[2] - lOrg[2];
            tw = sqrtf((ex * ex + ey * ey) + (ez * ez + 0.f));
          } else {
            tw = is / zFactor + near;
          }
          hit.u = uv0x + x * uvEx;
          hit.v = uv0y + y * uvEy;
          hit.geomID = geomID;
          hit.primID = primID;
          hit.block = block_index;
          tfar = is;
          hit.tfar = tw;
        }
      }
    }
  }
#undef BST
    return;
  }

  // ---- quadtree, leaves pushed and popped like inner nodes (shallow blocks: C <= 3) ----
  unsigned stack[16];
  float boxStack[16][6];
  int sp = 0;
  stack[0] = 0;
  boxStack[0][0] = -1.f; boxStack[0][1] = -1.f; boxStack[0][2] = box[0];
  boxStack[0][3] = 1.f;  boxStack[0][4] = 1.f;  boxStack[0][5] = box[1];
  float tNear[4] = {0.f, 0.f, 0.f, 0.f}, tFar[4] = {0.f, 0.f, 0.f, 0.f};

  while (sp >= 0) {
    const unsigned curr = stack[sp];
    float pBox[6];
#pragma unroll
    for (int k = 0; k < 6; k++) pBox[k] = boxStack[sp][k];
    sp--;

    if (curr >= elems) {
      const unsigned idx = curr - elems;
      ECB_CHECK(idx < 3u * elems + 1u, 4);
      if (STATS) work.cells++;
      if (LEAF == 1) {
        const float box_near = (idx & 2) ? ((idx & 1) ? tNear[3] : tNear[2]) : ((idx & 1) ? tNear[1] : tNear[0]);
        if (box_near >= tfar) continue;
        const float box_far = (idx & 2) ? ((idx & 1) ? tFar[3] : tFar[2]) : ((idx & 1) ? tFar[1] : tFar[0]);
        const float dimZ = pBox[5] - pBox[2];
        const float range = (1.f + 2.f * extent) * dimZ;
        const float dz = 0.0625f * range;
        const float offset = pBox[2] - dimZ * extent;
        const float rcpF = 0.0625f * range;
        const uint32_t zz = __ldg(reinterpret_cast<const unsigned short*>(blk + sc.leaf_offset) + idx);
        const unsigned z12 = zz & 0xff, z34 = zz >> 8;
        const float z1 = offset + rcpF * (float)((z12 & 0xf0) >> 4);
        const float z2 = offset + rcpF * (float)(z12 & 0x0f);
        const float z3 = offset + rcpF * (float)((z34 & 0xf0) >> 4);
        const float z4 = offset + rcpF * (float)(z34 & 0x0f);
        float u, v, t = tfar;
        if (intersect_patch(idx, sc.rcp_edges, dz, box_near, box_far, z1, z2, z3, z4, pBox, org, dir, u, v, t)) {
          float tw;
          if (special) {
            float ip[9], q[3], p[3];
            inverse3(proj, ip);
#pragma unroll
            for (int k = 0; k < 3; k++) q[k] = org[k] + dir[k] * t;
            project3(ip, q, p);
            const float ex = p[0] - lOrg[0], ey = p[1] - lOrg[1], ez = p[2] - lOrg[2];
            tw = sqrtf((ex * ex + ey * ey) + (ez * ez + 0.f));
          } else {
            tw = t / zFactor + near;
          }
          hit.u = uv0x + u * uvEx;
          hit.v = uv0y + v * uvEy;
          hit.geomID = geomID;
          hit.primID = primID;
          hit.block = block_index;
          tfar = t;
          hit.tfar = tw;
        }
      } else if (LEAF == 0) {
        const unsigned x = compact1by1(idx), y = compact1by1(idx >> 1);
        const float* grid = reinterpret_cast<const float*>(blk + sc.leaf_offset);
        const unsigned g0 = y * sc.grid_width + x, g2 = g0 + sc.grid_width;
        float V[4][3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
          V[0][k] = __ldg(grid + 3 * g0 + k);
          V[1][k] = __ldg(grid + 3 * g0 + 3 + k);
          V[2][k] = __ldg(grid + 3 * g2 + k);
          V[3][k] = __ldg(grid + 3 * g2 + 3 + k);
        }
        float tf = hit.tfar, uu = 0.f, vv = 0.f;
        const bool hit1 = intersect_triangle(V[0], V[1], V[2], rorg, rdir, ray_tnear, tf, uu, vv);
        const bool hit2 = intersect_triangle(V[3], V[2], V[1], rorg, rdir, ray_tnear, tf, uu, vv);
        if (hit1 || hit2) {
          float lu, lv;
          if (hit2) {
            lu = ((float)x + (1.f - uu)) * sc.rcp_edges;
            lv = ((float)y + (1.f - vv)) * sc.rcp_edges;
          } else {
            lu = ((float)x + uu) * sc.rcp_edges;
            lv = ((float)y + vv) * sc.rcp_edges;
          }
          hit.tfar = tf;
          hit.u = uv0x + lu * uvEx;
          hit.v = uv0y + lv * uvEy;
          hit.geomID = geomID;
          hit.primID = primID;
          hit.block = block_index;
        }
      } else {
        const float is = (idx & 2) ? ((idx & 1) ? tNear[3] : tNear[2]) : ((idx & 1) ? tNear[1] : tNear[0]);
        if (is <= tfar) {
          const float mx = (float)compact1by1(idx), my = (float)compact1by1(idx >> 1);
          const float x = (((org[0] + dir[0] * is) - pBox[0]) / (pBox[3] - pBox[0]) + mx) * sc.rcp_edges;
          const float y = (((org[1] + dir[1] * is) - pBox[1]) / (pBox[4] - pBox[1]) + my) * sc.rcp_edges;
          float tw;
          if (special) {
            float ip[9], q[3], p[3];
            inverse3(proj, ip);
#pragma unroll
            for (int k = 0; k < 3; k++) q[k] = org[k] + dir[k] * is;
            project3(ip, q, p);
            const float ex = p[0] - lOrg[0], ey = p[1] - lOrg[1], ez = p[2] - lOrg[2];
            tw = sqrtf((ex * ex + ey * ey) + (ez * ez + 0.f));
          } else {
            tw = is / zFactor + near;
          }
          hit.u = uv0x + x * uvEx;
          hit.v = uv0y + y * uvEy;
          hit.geomID = geomID;
          hit.primID = primID;
          hit.block = block_index;
          tfar = is;
          hit.tfar = tw;
        }
      }
      continue;
    }

    // inner node: decode (Node::getNode), test 4 children, push the hit ones nearest-on-top
    const unsigned next = curr * 4 + 1;
    if (STATS) work.cnodes++;
    ECB_CHECK(curr < elems, 4);
    const uint32_t w = curr == 0 ? root_word : __ldg(node_words + (curr - 1));
    const unsigned xz = w & 0xff, xb = (w >> 8) & 0xff, yz = (w >> 16) & 0xff, yb = w >> 24;
    const float dimX = pBox[3] - pBox[0], dimY = pBox[4] - pBox[1], dimZ = pBox[5] - pBox[2];
    const float lx02 = lut[(xz & 0xe0) >> 5] * dimX + pBox[0];
    const float lx13 = lut[8 + ((xz & 0x1c) >> 2)] * dimX + pBox[0];
    const float ux02 = lut[16 + ((xb & 0xe0) >> 5)] * dimX + pBox[0];
    const float ux13 = lut[24 + ((xb & 0x1c) >> 2)] * dimX + pBox[0];
    const float ly01 = lut[(yz & 0xe0) >> 5] * dimY + pBox[1];
    const float ly23 = lut[8 + ((yz & 0x1c) >> 2)] * dimY + pBox[1];
    const float uy01 = lut[16 + ((yb & 0xe0) >> 5)] * dimY + pBox[1];
    const float uy23 = lut[24 + ((yb & 0x1c) >> 2)] * dimY + pBox[1];
    const float lz = (0.25f * (float)(xz & 3)) * dimZ + pBox[2];
    const float uz = (1.f - 0.25f * (float)(yz & 3)) * dimZ + pBox[2];

    // the four children share their planes: x ranges A (children 0,2) / B (1,3), y ranges A (0,1) / B (2,3), one z range
    // (compressed_node.h:488-510) -> 5 near + 5 far plane distances instead of 24; same operands, same floats
    const float tnxA = ((tr.nu[0] ? ux02 : lx02) - tr.org[0]) * tr.rdn[0], tfxA = ((tr.nu[0] ? lx02 : ux02) - tr.org[0]) * tr.rdf[0];
    const float tnxB = ((tr.nu[0] ? ux13 : lx13) - tr.org[0]) * tr.rdn[0], tfxB = ((tr.nu[0] ? lx13 : ux13) - tr.org[0]) * tr.rdf[0];
    const float tnyA = ((tr.nu[1] ? uy01 : ly01) - tr.org[1]) * tr.rdn[1], tfyA = ((tr.nu[1] ? ly01 : uy01) - tr.org[1]) * tr.rdf[1];
    const float tnyB = ((tr.nu[1] ? uy23 : ly23) - tr.org[1]) * tr.rdn[1], tfyB = ((tr.nu[1] ? ly23 : uy23) - tr.org[1]) * tr.rdf[1];
    const float tnz = ((tr.nu[2] ? uz : lz) - tr.org[2]) * tr.rdn[2], tfz = ((tr.nu[2] ? lz : uz) - tr.org[2]) * tr.rdf[2];
    const float tnzn = sse_max(tnz, tr.tnear), tfzf = sse_min(tfz, tr.tfar);
    tNear[0] = sse_max(sse_max(tnxA, tnyA), tnzn); tFar[0] = sse_min(sse_min(tfxA, tfyA), tfzf);
    tNear[1] = sse_max(sse_max(tnxB, tnyA), tnzn); tFar[1] = sse_min(sse_min(tfxB, tfyA), tfzf);
    tNear[2] = sse_max(sse_max(tnxA, tnyB), tnzn); tFar[2] = sse_min(sse_min(tfxA, tfyB), tfzf);
    tNear[3] = sse_max(sse_max(tnxB, tnyB), tnzn); tFar[3] = sse_min(sse_min(tfxB, tfyB), tfzf);
    const unsigned mask = (tNear[0] <= tFar[0] ? 1u : 0u) | (tNear[1] <= tFar[1] ? 2u : 0u) | (tNear[2] <= tFar[2] ? 4u : 0u) |
                          (tNear[3] <= tFar[3] ? 8u : 0u);
    if (mask == 0) continue;

    // stack position of child r = sp+1 + #(hit children that must be popped after it)   (compressed.h:668-749)
    const int n = __popc(mask);
    ECB_CHECK(sp + n < 16, 3);
#pragma unroll
    for (int r = 0; r < 4; r++) {
      if (!(mask & (1u << r))) continue;
      int below = 0;  // number of other hit children popped AFTER r  (farther, or equal and higher index)
#pragma unroll
      for (int q = 0; q < 4; q++) {
        if (q == r || !(mask & (1u << q))) continue;
        const bool r_first = (q > r) ? (tNear[r] <= tNear[q]) : (tNear[r] < tNear[q]);
        below += r_first ? 1 : 0;
      }
      const int pos = sp + 1 + below;
      stack[pos] = next + r;
      boxStack[pos][0] = (r & 1) ? lx13 : lx02;
      boxStack[pos][1] = (r & 2) ? ly23 : ly01;
      boxStack[pos][2] = lz;
      boxStack[pos][3] = (r & 1) ? ux13 : ux02;
      boxStack[pos][4] = (r & 2) ? uy23 : uy01;
      boxStack[pos][5] = uz;
    }
    sp += n;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// top-level traversal
// ---------------------------------------------------------------------------------------------------------------
// The near / far plane of each axis is picked by ADDRESS (like the reference's nearX/farX byte offsets,
// node_intersector1.h:126-131): six 16-byte read-only loads deliver near and far planes of all four children.
__device__ __forceinline__ unsigned top_node(const EcbTopNode* __restrict__ nodes, uint32_t index, const TravRay& tr, float tNear[4],
                                             uint32_t child[4]) {
  const float4* p = reinterpret_cast<const float4*>(nodes + index);
  const float4 nx = __ldg(p + tr.nu[0]), fx = __ldg(p + (1 - tr.nu[0]));
  const float4 ny = __ldg(p + 2 + tr.nu[1]), fy = __ldg(p + (3 - tr.nu[1]));
  const float4 nz = __ldg(p + 4 + tr.nu[2]), fz = __ldg(p + (5 - tr.nu[2]));
  const uint4 c = __ldg(reinterpret_cast<const uint4*>(p + 6));
  child[0] = c.x; child[1] = c.y; child[2] = c.z; child[3] = c.w;
  unsigned mask = 0;
#define ECB_TOP_CHILD(i, m)                                                                                  \
  {                                                                                                          \
    const float tnx = (nx.m - tr.org[0]) * tr.rdn[0], tny = (ny.m - tr.org[1]) * tr.rdn[1], tnz = (nz.m - tr.org[2]) * tr.rdn[2]; \
    const float tfx = (fx.m - tr.org[0]) * tr.rdf[0], tfy = (fy.m - tr.org[1]) * tr.rdf[1], tfz = (fz.m - tr.org[2]) * tr.rdf[2]; \
    const float tn = sse_max(sse_max(tnx, tny), sse_max(tnz, tr.tnear));                                     \
    const float tf = sse_min(sse_min(tfx, tfy), sse_min(tfz, tr.tfar));                                      \
    tNear[i] = tn;                                                                                           \
    if (tn <= tf) mask |= 1u << i;                                                                           \
  }
  ECB_TOP_CHILD(0, x)
  ECB_TOP_CHILD(1, y)
  ECB_TOP_CHILD(2, z)
  ECB_TOP_CHILD(3, w)
#undef ECB_TOP_CHILD
  return mask;
}

#ifdef ECB_PRETEST
// experiment: can the ray hit ANY of the four children's outward-rounded bfloat16 boxes?  false => the exact
// test cannot succeed either, the 112-byte node is not fetched
__device__ __forceinline__ bool pretest_node(const uint4* __restrict__ pre, uint32_t index, const TravRay& tr) {
  const uint4 qx = __ldg(pre + 3 * (size_t)index), qy = __ldg(pre + 3 * (size_t)index + 1), qz = __ldg(pre + 3 * (size_t)index + 2);
  // q.x, q.y = lower planes of children 0..3 (two halfs per word), q.z, q.w = upper planes
  const uint32_t nxw0 = tr.nu[0] ? qx.z : qx.x, nxw1 = tr.nu[0] ? qx.w : qx.y, fxw0 = tr.nu[0] ? qx.x : qx.z, fxw1 = tr.nu[0] ? qx.y : qx.w;
  const uint32_t nyw0 = tr.nu[1] ? qy.z : qy.x, nyw1 = tr.nu[1] ? qy.w : qy.y, fyw0 = tr.nu[1] ? qy.x : qy.z, fyw1 = tr.nu[1] ? qy.y : qy.w;
  const uint32_t nzw0 = tr.nu[2] ? qz.z : qz.x, nzw1 = tr.nu[2] ? qz.w : qz.y, fzw0 = tr.nu[2] ? qz.x : qz.z, fzw1 = tr.nu[2] ? qz.y : qz.w;
  bool any = false;
#define ECB_PRE_CHILD(w0, w1, hi)                                                                                                   \
  {                                                                                                                                 \
    const float nx = __uint_as_float(hi ? (nx##w0 & 0xFFFF0000u) : (nx##w0 << 16)), fx = __uint_as_float(hi ? (fx##w0 & 0xFFFF0000u) : (fx##w0 << 16)); \
    const float ny = __uint_as_float(hi ? (ny##w0 & 0xFFFF0000u) : (ny##w0 << 16)), fy = __uint_as_float(hi ? (fy##w0 & 0xFFFF0000u) : (fy##w0 << 16)); \
    const float nz = __uint_as_float(hi ? (nz##w0 & 0xFFFF0000u) : (nz##w0 << 16)), fz = __uint_as_float(hi ? (fz##w0 & 0xFFFF0000u) : (fz##w0 << 16)); \
    const float tn = sse_max(sse_max((nx - tr.org[0]) * tr.rdn[0], (ny - tr.org[1]) * tr.rdn[1]), sse_max((nz - tr.org[2]) * tr.rdn[2], tr.tnear)); \
    const float tf = sse_min(sse_min((fx - tr.org[0]) * tr.rdf[0], (fy - tr.org[1]) * tr.rdf[1]), sse_min((fz - tr.org[2]) * tr.rdf[2], tr.tfar));   \
    any = any || tn <= tf;                                                                                                          \
  }
  ECB_PRE_CHILD(w0, w0, 0)
  ECB_PRE_CHILD(w0, w0, 1)
  ECB_PRE_CHILD(w1, w1, 0)
  ECB_PRE_CHILD(w1, w1, 1)
#undef ECB_PRE_CHILD
  return any;
}
#endif

__device__ __forceinline__ void cswap(uint32_t& ra, uint32_t& da, uint32_t& rb, uint32_t& db) {
  // if (b.dist < a.dist) xchg(b,a)   (stack_item.h:40-60)
  if (db < da) {
    const uint32_t tr = ra, td = da;
    ra = rb; da = db;
    rb = tr; db = td;
  }
}

// Pops / descends the top-level BVH4 until the ray reaches a block (returns its leaf ref) or the stack runs empty
// (returns ECB_REF_EMPTY).  Resumable: the caller keeps stack and sp between calls (one call per block entered).
// Restates the pop loop + down loop of bvh_intersector1.cpp:73-105 with traverseClosestHit (bvh_traverser1.h:436-506).
template <bool STATS>
__device__ __forceinline__ uint32_t next_block(const EcbTopNode* __restrict__ nodes, const TravRay& tr, uint32_t* stack_ref, uint32_t* stack_dist,
                                               int& sp, float ray_tfar, Work& work, uint32_t n_nodes = 0xFFFFFFFFu, const uint4* pre = nullptr) {
  while (sp > 0) {
    sp--;
    uint32_t cur = stack_ref[sp];
    if (__uint_as_float(stack_dist[sp]) > ray_tfar) continue;
    bool popped = false;
    while (!(cur & ECB_REF_LEAF)) {
      if (STATS) work.top_nodes++;
      ECB_CHECK(cur < n_nodes, 1);
      ECB_CHECK(sp + 3 <= (int)ECB_MAX_TOP_STACK, 0);
#ifdef ECB_PRETEST
      if (pre && !pretest_node(pre, cur, tr)) { popped = true; break; }
#endif
      float tNear[4];
      uint32_t child[4];
      const unsigned mask = top_node(nodes, cur, tr, tNear, child);
      if (mask == 0) { popped = true; break; }
      cur = ecb_order_hit_children(mask, child, tNear, stack_ref, stack_dist, sp);  // traverseClosestHit + sort networks
    }
    if (popped || cur == ECB_REF_EMPTY) continue;
    return cur;
  }
  return ECB_REF_EMPTY;
}

// any-hit: the first block whose world box the ray segment overlaps occludes (compressed.h:754-756 returns true)
__device__ __forceinline__ bool trace_any(const EcbDeviceScene& sc, const float org[3], const float dir[3], float ray_tnear,
                                          float ray_tfar) {
  uint32_t stack_ref[ECB_MAX_TOP_STACK];
  int sp = 1;
  stack_ref[0] = sc.root;
  TravRay tr;
  travray_init(tr, org, dir, sse_max_exact(ray_tnear, 0.0f), sse_max_exact(ray_tfar, 0.0f));
  while (sp > 0) {
    sp--;
    uint32_t cur = stack_ref[sp];
    bool popped = false;
    while (!(cur & ECB_REF_LEAF)) {
      ECB_CHECK(cur < sc.n_top, 1);
      ECB_CHECK(sp + 3 <= (int)ECB_MAX_TOP_STACK, 0);
      float tNear[4];
      uint32_t child[4];
      const unsigned mask = top_node(sc.top, cur, tr, tNear, child);
      if (mask == 0) { popped = true; break; }
      int last = -1;
#pragma unroll
      for (int r = 0; r < 4; r++)
        if (mask & (1u << r)) {
          if (last >= 0) stack_ref[sp++] = child[last];
          last = r;
        }
      cur = child[last];
    }
    if (popped || cur == ECB_REF_EMPTY) continue;
    return true;
  }
  return false;
}

// ---------------------------------------------------------------------------------------------------------------
// closest-hit kernel: persistent warps, dynamic ray fetch with lane refill, while-while traversal
//
//   * The grid is sized to the machine (SMs x ECB_BLOCKS_PER_SM CTAs); warps pull chunks of ECB_CHUNK consecutive ray
//     indices from a global queue (one atomicAdd per chunk) and hand them to their idle lanes.
//   * Every lane keeps its own ray, top-level stack and hit.  One outer iteration lets each busy lane run to its next
//     block (or to termination) and then traverse that block; when fewer than ECB_REFILL_MIN lanes are still busy the
//     warp goes back to refill the idle ones, so short rays (sky pixels, bounce rays that leave the scene at once)
//     do not leave lanes idle for the duration of the longest ray.
//   * Rays are read from `in` (stride in_stride: RTCRayHit = 80, or a packed RTCRay copy = 48) and hits are written
//     to `out` (the caller's RTCRayHit array; may be mapped pinned host memory -> no device-to-host copy at all).
// ---------------------------------------------------------------------------------------------------------------
template <int LEAF, bool STATS, bool DEEP>
__global__ void __launch_bounds__(ECB_THREADS, ECB_MIN_BLOCKS) ecb_intersect_kernel(EcbDeviceScene sc, EcbLaunch L) {
  __shared__ float lut[32];
  init_lut(lut);
  __syncthreads();
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  const unsigned block_min = L.block_min ? L.block_min : ECB_BLOCK_MIN, refill_min = L.refill_min ? L.refill_min : ECB_REFILL_IDLE_MIN;
  Work work = {0u, 0u, 0u, 0u};
  uint32_t n_rays = 0, n_hits = 0;

  // warp-private chunk of ray indices (uniform across the warp)
  uint64_t chunk_next = 0, chunk_end = 0;
  bool exhausted = false;

  // per-lane ray state
  bool busy = false;
  uint64_t index = 0;
  float org[3] = {0.f, 0.f, 0.f}, dir[3] = {0.f, 0.f, 1.f};
  float ray_tnear = 0.f;
  TravRay tr;
  Hit hit;
  hit.tfar = 0.f; hit.u = hit.v = 0.f; hit.primID = hit.geomID = hit.block = 0xFFFFFFFFu;
  uint32_t stack_ref[ECB_MAX_TOP_STACK];
  uint32_t stack_dist[ECB_MAX_TOP_STACK];
  int sp = 0;
  travray_init(tr, org, dir, 0.f, 0.f);

  // lane states: idle (no ray), busy && pending == EMPTY (traversing the top level), busy && pending != EMPTY (waiting
  // at a block).  Block traversal is the expensive phase, so lanes WAIT at their block until at least ECB_BLOCK_MIN
  // lanes have one (or no other work can be found) -- the other lanes keep fetching and traversing meanwhile.
  uint32_t pending = ECB_REF_EMPTY;
  for (;;) {
    unsigned idle_mask = __ballot_sync(0xffffffffu, !busy);
    const unsigned n_trav0 = __popc(__ballot_sync(0xffffffffu, busy && pending == ECB_REF_EMPTY));
    const unsigned n_pend0 = __popc(__ballot_sync(0xffffffffu, busy && pending != ECB_REF_EMPTY));

    // ---- refill idle lanes (when enough are idle to make the ray set-up worth a pass, or nothing else can run) ----
    if (!exhausted && idle_mask && (__popc(idle_mask) >= refill_min || (n_trav0 == 0 && n_pend0 < block_min))) {
      unsigned need = idle_mask;
      while (need && !exhausted) {
        if (chunk_next == chunk_end) {
          unsigned long long base = 0;
          if (lane == 0) base = atomicAdd(L.queue, (unsigned long long)ECB_CHUNK);
          base = __shfl_sync(0xffffffffu, base, 0);
          if (base >= L.M) { exhausted = true; break; }
          chunk_next = base;
          chunk_end = base + ECB_CHUNK < L.M ? base + ECB_CHUNK : L.M;
        }
        const unsigned avail = (unsigned)(chunk_end - chunk_next);
        const unsigned n_need = __popc(need);
        const unsigned take = avail < n_need ? avail : n_need;
        const unsigned rank = __popc(need & lt_mask);
        const bool mine = ((need >> lane) & 1u) && rank < take;
        if (mine) {
          index = chunk_next + rank;
          ECB_CHECK(index < L.M, 7);
          float4 r0, r1;
          float tfar_in;
          load_ray(L, index, r0, r1, tfar_in);
          if (L.block_out) L.block_out[index] = 0xFFFFFFFFu;
          if (L.hit_geom && !L.keep_results) L.hit_geom[index] = 0xFFFFFFFFu;
          if (!(r0.w > tfar_in) && sc.root != ECB_REF_EMPTY && !(L.any_hit_exact && tfar_in < 0.0f)) {  // tnear > tfar: inactive stream ray
            org[0] = r0.x; org[1] = r0.y; org[2] = r0.z;
            dir[0] = r1.x; dir[1] = r1.y; dir[2] = r1.z;
            ray_tnear = r0.w;
            hit.tfar = tfar_in;
            hit.block = 0xFFFFFFFFu;
            travray_init(tr, org, dir, sse_max_exact(ray_tnear, 0.0f), sse_max_exact(tfar_in, 0.0f));
            stack_ref[0] = sc.root;
            stack_dist[0] = 0xff800000u;  // -inf
            sp = 1;
            busy = true;
            pending = ECB_REF_EMPTY;
            if (STATS) n_rays++;
          }
        }
        chunk_next += take;
        // lanes served in this round (busy or consumed an inactive ray) leave the need mask
        need &= ~__ballot_sync(0xffffffffu, mine);
      }
    }

    // ---- top level: every traversing lane runs to its next block or terminates ----
    if (busy && pending == ECB_REF_EMPTY) {
      pending = next_block<STATS>(sc.top, tr, stack_ref, stack_dist, sp, hit.tfar, work, sc.n_top, static_cast<const uint4*>(sc.pre));
      if (pending == ECB_REF_EMPTY) {
        // terminated: write the hit (a miss writes nothing: rtcIntersect contract)
        busy = false;
        if (hit.block != 0xFFFFFFFFu && !L.any_hit_exact) {
          if (L.hit_geom) {  // compact result: the host layer scatters it into the caller's record
            L.hit_rec[index] = make_float4(hit.tfar, hit.u, hit.v, __uint_as_float(hit.primID));
            L.hit_geom[index] = hit.geomID;
            if (L.hit_ng) { L.hit_ng[3 * index] = 1.f; L.hit_ng[3 * index + 1] = 0.f; L.hit_ng[3 * index + 2] = 0.f; }  // replaces a triangle hit's normal
          } else {
            uint8_t* op = L.out + index * L.out_stride;
            *reinterpret_cast<float*>(op + 32) = hit.tfar;
            *reinterpret_cast<float4*>(op + 48) = make_float4(1.f, 0.f, 0.f, hit.u);  // Ng = (1,0,0) dummy (compressed.h:574), u
            *reinterpret_cast<float4*>(op + 64) =
                make_float4(hit.v, __uint_as_float(hit.primID), __uint_as_float(hit.geomID), __uint_as_float(L.instID));
          }
          if (L.block_out) L.block_out[index] = hit.block;
          if (STATS) n_hits++;
        }
      }
    }

    // ---- blocks: when enough lanes wait at one, or the idle lanes cannot be given other work ----
    const unsigned n_pend = __popc(__ballot_sync(0xffffffffu, busy && pending != ECB_REF_EMPTY));
    const unsigned n_idle = __popc(__ballot_sync(0xffffffffu, !busy));
    if (n_pend >= block_min || (n_pend > 0 && (exhausted || n_idle < refill_min))) {
      if (busy && pending != ECB_REF_EMPTY) {
        block_intersect<LEAF, STATS, DEEP>(sc, lut, pending & ~ECB_REF_LEAF, org, dir, ray_tnear, hit, work);
        tr.tfar = hit.tfar;  // bvh_intersector1.cpp:117
        pending = ECB_REF_EMPTY;
        if (L.any_hit_exact && hit.block != 0xFFFFFFFFu) {
          // opt-in exact shadow query (NOT the reference's behaviour, which stops at the block's bounding box): any
          // accepted leaf hit inside [tnear, tfar] occludes
          if (L.hit_geom) L.hit_geom[index] = 0u;
          else *reinterpret_cast<float*>(L.out + index * L.out_stride + 32) = __int_as_float(0xff800000);
          busy = false;
          if (STATS) n_hits++;
        }
      }
    } else if (n_pend == 0 && exhausted && n_idle == 32u) {
      break;
    }
  }
  if (STATS) {
    atomicAdd(L.counters + 0, (unsigned long long)n_rays);
    atomicAdd(L.counters + 1, (unsigned long long)work.top_nodes);
    atomicAdd(L.counters + 2, (unsigned long long)work.blocks);
    atomicAdd(L.counters + 3, (unsigned long long)work.cnodes);
    atomicAdd(L.counters + 4, (unsigned long long)work.cells);
    atomicAdd(L.counters + 5, (unsigned long long)n_hits);
  }
}

#include "ecb_quad_kernel.cuh"

// ---------------------------------------------------------------------------------------------------------------
// instanced scenes: BVHNIntersector1<4,BVH_AN1,robust=false,ArrayIntersector1<ObjectIntersector1>> over the instances
// (kernels/bvh/bvh_intersector1_bvh4.cpp:81), FastInstanceIntersectorN::intersect1/occluded1 per instance
// (kernels/geometry/instance_intersector.cpp:45-83).  One thread per ray, two stacks (instance level, object level).
// ---------------------------------------------------------------------------------------------------------------
struct FastRay {
  float rdir[3], org_rdir[3];
  int nu[3];
  float tnear, tfar;
};

__device__ __forceinline__ void fastray_init(FastRay& r, const float org[3], const float dir[3], float tnear, float tfar) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    r.rdir[k] = 1.0f / zero_fix(dir[k]);  // rcp_safe (node_intersector1.h:41)
    r.org_rdir[k] = org[k] * r.rdir[k];
    r.nu[k] = r.rdir[k] >= 0.0f ? 0 : 1;
  }
  r.tnear = tnear;
  r.tfar = tfar;
}

// intersectNode<4,4> AVX2 / SSE4.1 path (node_intersector1.h:208-232): fused plane*rdir - org*rdir, integer max/min
// on the float bits (maxi/mini), integer compare for the mask
__device__ __forceinline__ float maxi_f(float a, float b) { return __float_as_int(a) > __float_as_int(b) ? a : b; }
__device__ __forceinline__ float mini_f(float a, float b) { return __float_as_int(a) < __float_as_int(b) ? a : b; }

__device__ __forceinline__ unsigned fast_node(const EcbTopNode* nodes, uint32_t index, const FastRay& tr, float tNear[4], uint32_t child[4]) {
  const float4* p = reinterpret_cast<const float4*>(nodes + index);
  const float4 nx = __ldg(p + tr.nu[0]), fx = __ldg(p + (1 - tr.nu[0]));
  const float4 ny = __ldg(p + 2 + tr.nu[1]), fy = __ldg(p + (3 - tr.nu[1]));
  const float4 nz = __ldg(p + 4 + tr.nu[2]), fz = __ldg(p + (5 - tr.nu[2]));
  const uint4 c = __ldg(reinterpret_cast<const uint4*>(p + 6));
  child[0] = c.x; child[1] = c.y; child[2] = c.z; child[3] = c.w;
  unsigned mask = 0;
#define ECB_FAST_CHILD(i, m)                                                                           \
  {                                                                                                    \
    const float tnx = __fmaf_rn(nx.m, tr.rdir[0], -tr.org_rdir[0]), tny = __fmaf_rn(ny.m, tr.rdir[1], -tr.org_rdir[1]); \
    const float tnz = __fmaf_rn(nz.m, tr.rdir[2], -tr.org_rdir[2]);                                    \
    const float tfx = __fmaf_rn(fx.m, tr.rdir[0], -tr.org_rdir[0]), tfy = __fmaf_rn(fy.m, tr.rdir[1], -tr.org_rdir[1]); \
    const float tfz = __fmaf_rn(fz.m, tr.rdir[2], -tr.org_rdir[2]);                                    \
    const float tn = maxi_f(maxi_f(tnx, tny), maxi_f(tnz, tr.tnear));                                  \
    const float tf = mini_f(mini_f(tfx, tfy), mini_f(tfz, tr.tfar));                                   \
    tNear[i] = tn;                                                                                     \
    if (!(__float_as_int(tn) > __float_as_int(tf))) mask |= 1u << i;                                   \
  }
  ECB_FAST_CHILD(0, x)
  ECB_FAST_CHILD(1, y)
  ECB_FAST_CHILD(2, z)
  ECB_FAST_CHILD(3, w)
#undef ECB_FAST_CHILD
  return mask;
}

// closest hit inside one object scene (what rtcIntersect1 on the instanced scene does)
template <int LEAF, bool STATS>
__device__ __forceinline__ void trace_object_body(const EcbDeviceScene& sc, const float* lut, const float org[3], const float dir[3],
                                                  float ray_tnear, Hit& hit, Work& work) {
  uint32_t stack_ref[ECB_MAX_TOP_STACK];
  uint32_t stack_dist[ECB_MAX_TOP_STACK];
  if (sc.root == ECB_REF_EMPTY) return;
  TravRay tr;
  travray_init(tr, org, dir, sse_max_exact(ray_tnear, 0.0f), sse_max_exact(hit.tfar, 0.0f));
  stack_ref[0] = sc.root;
  stack_dist[0] = 0xff800000u;
  int sp = 1;
  for (;;) {
    const uint32_t leaf = next_block<STATS>(sc.top, tr, stack_ref, stack_dist, sp, hit.tfar, work, sc.n_top);
    if (leaf == ECB_REF_EMPTY) break;
    block_intersect<LEAF, STATS>(sc, lut, leaf & ~ECB_REF_LEAF, org, dir, ray_tnear, hit, work);
    tr.tfar = hit.tfar;
  }
}
// out-of-line form for instancing scenes whose object scenes mix leaf kinds (three instantiations would triple the kernel)
template <int LEAF, bool STATS>
__device__ __noinline__ void trace_object(const EcbDeviceScene& sc, const float* lut, const float org[3], const float dir[3],
                                          float ray_tnear, Hit& hit, Work& work) {
  trace_object_body<LEAF, STATS>(sc, lut, org, dir, ray_tnear, hit, work);
}

__device__ __forceinline__ void instance_ray(const EcbInstance* in, const float o[3], const float d[3], float lo[3], float ld[3],
                                             uint32_t& geomID, uint32_t& object) {
  const float4* q = reinterpret_cast<const float4*>(in);
  const float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
  const uint4 e = __ldg(reinterpret_cast<const uint4*>(q + 3));
  const float m[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
  for (int k = 0; k < 3; k++) {  // xfmPoint / xfmVector: nested madd = FMA on AVX2 (affinespace.h:110-111)
    lo[k] = __fmaf_rn(o[0], m[k], __fmaf_rn(o[1], m[3 + k], __fmaf_rn(o[2], m[6 + k], m[9 + k])));
    ld[k] = __fmaf_rn(d[0], m[k], __fmaf_rn(d[1], m[3 + k], d[2] * m[6 + k]));
  }
  geomID = e.x;
  object = e.y;
}

// KIND: leaf kind shared by every object scene (the usual case: one subdiv_accel per device) -> the object traversal is
// inlined; -1: mixed kinds, dispatched per visit through the out-of-line form
#ifndef ECB_INST_MIN_BLOCKS
#define ECB_INST_MIN_BLOCKS 1
#endif
template <bool ANYHIT, bool STATS, int KIND>
__global__ void __launch_bounds__(ECB_THREADS, ECB_INST_MIN_BLOCKS) ecb_instanced_kernel(EcbDeviceInstScene is, EcbLaunch L) {
  __shared__ float lut[32];
  init_lut(lut);
  __syncthreads();
  Work work = {0u, 0u, 0u, 0u};
  uint32_t n_visits = 0, n_hits = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < L.M; i += (uint64_t)gridDim.x * blockDim.x) {
    float4 r0, r1;
    float tfar_in;
    load_ray(L, i, r0, r1, tfar_in);
    if (L.hit_geom) L.hit_geom[i] = 0xFFFFFFFFu;  // compact result mode: miss / inactive / not occluded
    if (r0.w > tfar_in) continue;
    if (ANYHIT && tfar_in < 0.0f) continue;
    if (is.root == ECB_REF_EMPTY) continue;
    const float org[3] = {r0.x, r0.y, r0.z};
    const float dir[3] = {r1.x, r1.y, r1.z};
    Hit hit;
    hit.tfar = tfar_in;
    hit.u = hit.v = 0.f;
    hit.primID = hit.geomID = hit.block = 0xFFFFFFFFu;
    uint32_t hit_inst = 0xFFFFFFFFu;
    bool occluded = false;
    FastRay tr;
    fastray_init(tr, org, dir, sse_max_exact(r0.w, 0.0f), sse_max_exact(tfar_in, 0.0f));
    uint32_t stack_ref[ECB_MAX_TOP_STACK];
    uint32_t stack_dist[ECB_MAX_TOP_STACK];
    int sp = 1;
    stack_ref[0] = is.root;
    stack_dist[0] = 0xff800000u;
    while (sp > 0 && !occluded) {
      sp--;
      uint32_t cur = stack_ref[sp];
      if (!ANYHIT && __uint_as_float(stack_dist[sp]) > hit.tfar) continue;
      bool popped = false;
      while (!(cur & ECB_REF_LEAF)) {
        float tNear[4];
        uint32_t child[4];
        ECB_CHECK(cur < is.n_nodes, 1);
        ECB_CHECK(sp + 3 <= (int)ECB_MAX_TOP_STACK, 0);
        const unsigned mask = fast_node(is.nodes, cur, tr, tNear, child);
        if (mask == 0) { popped = true; break; }
        if ((mask & (mask - 1u)) == 0u) {  // one child hit: no compaction, no distances (closest-hit and any-hit alike)
          cur = (mask & 1u) ? child[0] : (mask & 2u) ? child[1] : (mask & 4u) ? child[2] : child[3];
          continue;
        }
        uint32_t c[4], d[4];
        int nh = 0;
#pragma unroll
        for (int r = 0; r < 4; r++)
          if (mask & (1u << r)) { c[nh] = child[r]; d[nh] = __float_as_uint(tNear[r]); nh++; }
        if (ANYHIT) {  // traverseAnyHit: push all but the last hit child in index order
          for (int k = 0; k + 1 < nh; k++) stack_ref[sp++] = c[k];
          cur = c[nh - 1];
          continue;
        }
        if (nh == 1) { cur = c[0]; continue; }
        if (nh == 2) {
          if (d[0] < d[1]) { stack_ref[sp] = c[1]; stack_dist[sp] = d[1]; sp++; cur = c[0]; }
          else             { stack_ref[sp] = c[0]; stack_dist[sp] = d[0]; sp++; cur = c[1]; }
          continue;
        }
        if (nh == 3) {
          uint32_t r1_ = c[2], d1 = d[2], r2 = c[1], d2 = d[1], r3 = c[0], d3 = d[0];
          cswap(r1_, d1, r2, d2);
          cswap(r2, d2, r3, d3);
          cswap(r1_, d1, r2, d2);
          stack_ref[sp] = r3; stack_dist[sp] = d3; sp++;
          stack_ref[sp] = r2; stack_dist[sp] = d2; sp++;
          cur = r1_;
          continue;
        }
        {
          uint32_t r1_ = c[3], d1 = d[3], r2 = c[2], d2 = d[2], r3 = c[1], d3 = d[1], r4 = c[0], d4 = d[0];
          cswap(r1_, d1, r2, d2);
          cswap(r3, d3, r4, d4);
          cswap(r1_, d1, r3, d3);
          cswap(r2, d2, r4, d4);
          cswap(r2, d2, r3, d3);
          stack_ref[sp] = r4; stack_dist[sp] = d4; sp++;
          stack_ref[sp] = r3; stack_dist[sp] = d3; sp++;
          stack_ref[sp] = r2; stack_dist[sp] = d2; sp++;
          cur = r1_;
        }
      }
      if (popped || cur == ECB_REF_EMPTY) continue;
      const uint32_t first = (cur & ~ECB_REF_LEAF) >> 3, num = cur & 7u;
      ECB_CHECK(first + num <= is.n_items, 5);
      for (uint32_t k = 0; k < num && !occluded; k++) {
        ECB_CHECK(__ldg(is.items + first + k) < is.n_instances, 5);
        const EcbInstance* in = is.instances + __ldg(is.items + first + k);
        float lo[3], ld[3];
        uint32_t geomID, object;
        instance_ray(in, org, dir, lo, ld, geomID, object);
        ECB_CHECK(object < is.n_objects, 5);
        const EcbDeviceScene sc = is.objects[object];
        if (ANYHIT) {
          if (sc.root != ECB_REF_EMPTY && trace_any(sc, lo, ld, r0.w, hit.tfar)) occluded = true;
        } else {
          const uint32_t before = hit.block;
          hit.block = 0xFFFFFFFFu;
          if (STATS) n_visits++;
          if (KIND >= 0) {
            trace_object_body<(KIND >= 0 ? KIND : 0), STATS>(sc, lut, lo, ld, r0.w, hit, work);
          } else {
            switch (sc.leaf_kind) {
              case 0: trace_object<0, STATS>(sc, lut, lo, ld, r0.w, hit, work); break;
              case 1: trace_object<1, STATS>(sc, lut, lo, ld, r0.w, hit, work); break;
              default: trace_object<2, STATS>(sc, lut, lo, ld, r0.w, hit, work); break;
            }
          }
          if (hit.block != 0xFFFFFFFFu) hit_inst = geomID;
          else hit.block = before;
        }
      }
      tr.tfar = hit.tfar;
    }
    if (STATS && hit_inst != 0xFFFFFFFFu) n_hits++;
    if (L.hit_geom) {  // compact result: the host layer scatters it into the caller's record
      if (ANYHIT) {
        if (occluded) L.hit_geom[i] = 0u;
      } else if (hit_inst != 0xFFFFFFFFu) {
        L.hit_rec[i] = make_float4(hit.tfar, hit.u, hit.v, __uint_as_float(hit.primID));
        L.hit_geom[i] = hit.geomID;
        L.hit_inst[i] = hit_inst;
      }
      continue;
    }
    uint8_t* op = L.out + i * L.out_stride;
    if (ANYHIT) {
      if (occluded) *reinterpret_cast<float*>(op + 32) = __int_as_float(0xff800000);
    } else if (hit_inst != 0xFFFFFFFFu) {
      *reinterpret_cast<float*>(op + 32) = hit.tfar;
      *reinterpret_cast<float4*>(op + 48) = make_float4(1.f, 0.f, 0.f, hit.u);
      *reinterpret_cast<float4*>(op + 64) =
          make_float4(hit.v, __uint_as_float(hit.primID), __uint_as_float(hit.geomID), __uint_as_float(hit_inst));
    }
  }
  if (STATS) {  // counters[0] = object-scene traversals (instance visits), then the plain kernel's counters
    atomicAdd(L.counters + 0, (unsigned long long)n_visits);
    atomicAdd(L.counters + 1, (unsigned long long)work.top_nodes);
    atomicAdd(L.counters + 2, (unsigned long long)work.blocks);
    atomicAdd(L.counters + 3, (unsigned long long)work.cnodes);
    atomicAdd(L.counters + 4, (unsigned long long)work.cells);
    atomicAdd(L.counters + 5, (unsigned long long)n_hits);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// mixed scenes: the triangle accel AccelN visits before the compressed accel.
// BVHNIntersector1<4,BVH_AN1,true,ArrayIntersector1<TriangleMvIntersector1Pluecker<4,4,true>>> (bvh_intersector1_bvh4.cpp:44):
// the same robust traversal, leaves = arrays of Triangle4v blocks, PlueckerIntersector1<4>::intersect
// (triangle_intersector_pluecker.h:77-127) + Intersect1EpilogM / Occluded1EpilogM (intersector_epilog.h:227-306).
// The reference runs the AVX (non-FMA) build of this intersector (bvh4_factory.cpp:408): madd/msub are separate
// multiply and add -- which is what -fmad=false compiles the expressions below to.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float min_ps(float a, float b) { return a < b ? a : b; }  // _mm_min_ps operand order
__device__ __forceinline__ float max_ps(float a, float b) { return a > b ? a : b; }
__device__ __forceinline__ float dot3(const float a[3], const float b[3]) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); }
__device__ __forceinline__ void cross3(const float a[3], const float b[3], float r[3]) {
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
}

struct TriHit {
  float t, u, v, ng[3];
  uint32_t primID, geomID;
};

// One Triangle4v block against one ray; returns whether a lane survived (and, for closest hit, the selected lane in `out`).
template <bool ANYHIT>
__device__ __forceinline__ bool tri4v_intersect(const EcbTriBlock* __restrict__ blk, const float O[3], const float D[3], float ray_tnear,
                                                float ray_tfar, TriHit& out) {
  const float4* q = reinterpret_cast<const float4*>(blk);
  float P[9][4];  // v0.x[4], v0.y[4], v0.z[4], v1..., v2...
#pragma unroll
  for (int r = 0; r < 9; r++) {
    const float4 w = __ldg(q + r);
    P[r][0] = w.x; P[r][1] = w.y; P[r][2] = w.z; P[r][3] = w.w;
  }
  float U[4], V[4], W[4], T[4], den[4], Ng[4][3];
  unsigned valid = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    float v0[3], v1[3], v2[3], e0[3], e1[3], e2[3], s[3], c[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      v0[k] = P[k][i] - O[k];
      v1[k] = P[3 + k][i] - O[k];
      v2[k] = P[6 + k][i] - O[k];
    }
#pragma unroll
    for (int k = 0; k < 3; k++) { e0[k] = v2[k] - v0[k]; e1[k] = v0[k] - v1[k]; e2[k] = v1[k] - v2[k]; }
#pragma unroll
    for (int k = 0; k < 3; k++) s[k] = v2[k] + v0[k];
    cross3(e0, s, c); U[i] = dot3(c, D);
#pragma unroll
    for (int k = 0; k < 3; k++) s[k] = v0[k] + v1[k];
    cross3(e1, s, c); V[i] = dot3(c, D);
#pragma unroll
    for (int k = 0; k < 3; k++) s[k] = v1[k] + v2[k];
    cross3(e2, s, c); W[i] = dot3(c, D);
    const float minUVW = min_ps(min_ps(U[i], V[i]), W[i]);
    const float maxUVW = max_ps(max_ps(U[i], V[i]), W[i]);
    if (minUVW >= 0.0f || maxUVW <= 0.0f) valid |= 1u << i;
    // stable_triangle_normal(e0,e1,e2) (common/math/vec3.h:200-212)
    const float ab_x = e0[2] * e1[1], ab_y = e0[0] * e1[2], ab_z = e0[1] * e1[0];
    const float bc_x = e1[2] * e2[1], bc_y = e1[0] * e2[2], bc_z = e1[1] * e2[0];
    const float cab0 = e0[1] * e1[2] - ab_x, cab1 = e0[2] * e1[0] - ab_y, cab2 = e0[0] * e1[1] - ab_z;
    const float cbc0 = e1[1] * e2[2] - bc_x, cbc1 = e1[2] * e2[0] - bc_y, cbc2 = e1[0] * e2[1] - bc_z;
    Ng[i][0] = fabsf(ab_x) < fabsf(bc_x) ? cab0 : cbc0;
    Ng[i][1] = fabsf(ab_y) < fabsf(bc_y) ? cab1 : cbc1;
    Ng[i][2] = fabsf(ab_z) < fabsf(bc_z) ? cab2 : cbc2;
    const float dn = dot3(Ng[i], D);
    den[i] = dn + dn;
    const float tt = dot3(v0, Ng[i]);
    T[i] = tt + tt;
  }
  if (!valid) return false;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float absDen = fabsf(den[i]);
    const float Ts = __uint_as_float(__float_as_uint(T[i]) ^ (__float_as_uint(den[i]) & 0x80000000u));
    const bool keep = (absDen * ray_tnear < Ts) && (Ts <= absDen * ray_tfar) && (den[i] != 0.0f);
    if (!keep) valid &= ~(1u << i);
  }
  if (!valid) return false;
  if constexpr (ANYHIT) {
    return true;
  } else {
    float vt[4], a[4], h[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      vt[i] = T[i] * (1.0f / den[i]);
      a[i] = (valid & (1u << i)) ? vt[i] : __int_as_float(0x7f800000);
    }
    // select_min (common/simd/vfloat4_sse2.h:654-659)
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = min_ps(a[i ^ 1], a[i]);
    unsigned vmin = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float m = min_ps(h[i ^ 2], h[i]);
      if ((valid & (1u << i)) && a[i] == m) vmin |= 1u << i;
    }
    const int i = __ffs(vmin ? vmin : valid) - 1;
    float Us = U[0], Vs = V[0], Ws = W[0];
    out.t = vt[0]; out.ng[0] = Ng[0][0]; out.ng[1] = Ng[0][1]; out.ng[2] = Ng[0][2];
#pragma unroll
    for (int k = 1; k < 4; k++)
      if (i == k) {
        Us = U[k]; Vs = V[k]; Ws = W[k];
        out.t = vt[k]; out.ng[0] = Ng[k][0]; out.ng[1] = Ng[k][1]; out.ng[2] = Ng[k][2];
      }
    const float UVW = (Us + Vs) + Ws;
    const float rcpUVW = fabsf(UVW) < 1E-18f ? 0.0f : 1.0f / UVW;
    out.u = Us * rcpUVW;
    out.v = Vs * rcpUVW;
    const uint32_t* ids = reinterpret_cast<const uint32_t*>(q + 9);
    out.geomID = __ldg(ids + i);
    out.primID = __ldg(ids + 4 + i);
    return true;
  }
}

template <bool ANYHIT>
__global__ void __launch_bounds__(ECB_THREADS) ecb_triangle_kernel(EcbDeviceTriScene ts, EcbLaunch L) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < L.M; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint8_t* rp = L.in + i * L.in_stride;
    float4 r0, r1;
    float tfar_in;
    load_ray(L, i, r0, r1, tfar_in);
    if (L.hit_geom) L.hit_geom[i] = 0xFFFFFFFFu;  // compact result mode: this pass runs first and owns the initial state
    if (r0.w > tfar_in) continue;               // inactive stream ray
    if (ANYHIT && tfar_in < 0.0f) continue;     // bvh_intersector1.cpp:133
    const float org[3] = {r0.x, r0.y, r0.z};
    const float dir[3] = {r1.x, r1.y, r1.z};
    TravRay tr;
    travray_init(tr, org, dir, sse_max_exact(r0.w, 0.0f), sse_max_exact(tfar_in, 0.0f));
    uint32_t stack_ref[ECB_MAX_TOP_STACK];
    uint32_t stack_dist[ECB_MAX_TOP_STACK];
    int sp = 1;
    stack_ref[0] = ts.root;
    stack_dist[0] = 0xff800000u;
    // compact result mode: tfar goes back into the input (staging) record, where the compressed-accel pass reads it
    uint8_t* op = L.hit_geom ? nullptr : L.out + i * L.out_stride;
    float* tfar_out = !L.hit_geom ? reinterpret_cast<float*>(op + 32)
                      : L.soa[0]  ? const_cast<float*>(L.soa[7]) + i
                                  : reinterpret_cast<float*>(const_cast<uint8_t*>(rp) + L.tfar_offset);
    if (ANYHIT) {
      bool occluded = false;
      while (sp > 0 && !occluded) {
        sp--;
        uint32_t cur = stack_ref[sp];
        bool popped = false;
        while (!(cur & ECB_REF_LEAF)) {
          float tNear[4];
          uint32_t child[4];
          ECB_CHECK(cur < ts.n_nodes, 1);
          ECB_CHECK(sp + 3 <= (int)ECB_MAX_TOP_STACK, 0);
          const unsigned mask = top_node(ts.nodes, cur, tr, tNear, child);
          if (mask == 0) { popped = true; break; }
          int last = -1;
#pragma unroll
          for (int r = 0; r < 4; r++)
            if (mask & (1u << r)) {
              if (last >= 0) stack_ref[sp++] = child[last];
              last = r;
            }
          cur = child[last];
        }
        if (popped || cur == ECB_REF_EMPTY) continue;
        const uint32_t first = (cur & ~ECB_REF_LEAF) >> 3, num = cur & 7u;
        ECB_CHECK(first + num <= ts.n_blocks, 6);
        TriHit th;
        for (uint32_t k = 0; k < num && !occluded; k++)
          occluded = tri4v_intersect<true>(ts.blocks + first + k, org, dir, r0.w, tfar_in, th);
      }
      if (occluded) {
        *tfar_out = __int_as_float(0xff800000);
        if (L.hit_geom) L.hit_geom[i] = 0u;
      }
    } else {
      Work work = {0u, 0u, 0u, 0u};
      float tfar = tfar_in;
      TriHit best;
      bool found = false;
      for (;;) {
        const uint32_t leaf = next_block<false>(ts.nodes, tr, stack_ref, stack_dist, sp, tfar, work, ts.n_nodes);
        if (leaf == ECB_REF_EMPTY) break;
        const uint32_t first = (leaf & ~ECB_REF_LEAF) >> 3, num = leaf & 7u;
        ECB_CHECK(first + num <= ts.n_blocks, 6);
        for (uint32_t k = 0; k < num; k++) {
          TriHit th;
          if (tri4v_intersect<false>(ts.blocks + first + k, org, dir, r0.w, tfar, th)) {
            best = th;
            tfar = th.t;
            found = true;
          }
        }
        tr.tfar = tfar;  // bvh_intersector1.cpp:117
      }
      if (found && L.hit_geom) {
        *tfar_out = best.t;
        L.hit_rec[i] = make_float4(best.t, best.u, best.v, __uint_as_float(best.primID));
        L.hit_geom[i] = best.geomID;
        L.hit_ng[3 * i] = best.ng[0]; L.hit_ng[3 * i + 1] = best.ng[1]; L.hit_ng[3 * i + 2] = best.ng[2];
      } else if (found) {
        *reinterpret_cast<float*>(op + 32) = best.t;
        *reinterpret_cast<float4*>(op + 48) = make_float4(best.ng[0], best.ng[1], best.ng[2], best.u);
        *reinterpret_cast<float4*>(op + 64) =
            make_float4(best.v, __uint_as_float(best.primID), __uint_as_float(best.geomID), __uint_as_float(L.instID));
      }
    }
  }
}

__global__ void __launch_bounds__(ECB_THREADS) ecb_occluded_kernel(EcbDeviceScene sc, EcbLaunch L) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < L.M; i += (uint64_t)gridDim.x * blockDim.x) {
    float4 r0, r1;
    float tfar_in;
    load_ray(L, i, r0, r1, tfar_in);
    bool occluded = false;
    if (!(r0.w > tfar_in) && !(tfar_in < 0.0f)) {  // inactive stream ray; already occluded (bvh_intersector1.cpp:133)
      const float org[3] = {r0.x, r0.y, r0.z};
      const float dir[3] = {r1.x, r1.y, r1.z};
      occluded = sc.root != ECB_REF_EMPTY && trace_any(sc, org, dir, r0.w, tfar_in);
    }
    if (L.hit_geom) {
      if (!L.keep_results) L.hit_geom[i] = occluded ? 0u : 0xFFFFFFFFu;
      else if (occluded) L.hit_geom[i] = 0u;  // the triangle pass wrote the entry first
    } else if (occluded) *reinterpret_cast<float*>(L.out + i * L.out_stride + 32) = __int_as_float(0xff800000);  // -inf (bvh_intersector1.cpp:199)
  }
}

// ---------------------------------------------------------------------------------------------------------------
// read-bandwidth probe (measurement support, not on the traversal path): every thread streams 128-bit ld.global.cg loads
// (L2-coherent, L1 bypassed) over a buffer `passes` times.  A working set well inside the 126 MB L2 measures the L2 read
// bandwidth the traversal kernels' scene traffic is served at; one of several GB measures HBM.  The roofline's L2
// denominator comes from this kernel (tools/bw_probe.py -> profiles/measured_l2.json).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ecb_bw_probe_kernel(const uint4* __restrict__ buf, size_t n_vec, int passes, unsigned* sink) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  unsigned acc = 0;
  for (int p = 0; p < passes; p++) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n_vec; i += 4 * stride) {  // four independent 16-byte loads in flight per thread
      const uint4 a = __ldcg(buf + i), b = __ldcg(buf + i + stride), c = __ldcg(buf + i + 2 * stride), d = __ldcg(buf + i + 3 * stride);
      acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
    }
    for (; i < n_vec; i += stride) {
      const uint4 a = __ldcg(buf + i);
      acc ^= a.x ^ a.y ^ a.z ^ a.w;
    }
  }
  if (acc == 0x9E3779B9u) *sink = acc;  // keeps the loads alive; the buffer is zero-filled, so this never fires
}

}  // namespace

cudaError_t ecb_bounds_check_read(unsigned long long out[8], int* enabled) {
#ifdef ECB_BOUNDS_CHECK
  *enabled = 1;
  return cudaMemcpyFromSymbol(out, g_ecb_check, 8 * sizeof(unsigned long long));
#else
  *enabled = 0;
  for (int i = 0; i < 8; i++) out[i] = 0;
  return cudaSuccess;
#endif
}

cudaError_t ecb_probe_read_bandwidth(size_t bytes, int passes, int device, double* gbs_out) {
  uint4* buf = nullptr;
  unsigned* sink = nullptr;
  cudaEvent_t a = nullptr, b = nullptr;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const size_t n_vec = bytes / 16;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&buf), n_vec * 16);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&sink), 4);
  if (e == cudaSuccess) e = cudaMemset(buf, 0, n_vec * 16);
  if (e == cudaSuccess) e = cudaEventCreate(&a);
  if (e == cudaSuccess) e = cudaEventCreate(&b);
  double best = 0.0;
  for (int rep = 0; rep < 6 && e == cudaSuccess; rep++) {  // first repetition warms the L2, best of the rest
    cudaEventRecord(a, nullptr);
    ecb_bw_probe_kernel<<<sms * 8, 256>>>(buf, n_vec, passes, sink);
    cudaEventRecord(b, nullptr);
    e = cudaEventSynchronize(b);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, a, b);
    if (e == cudaSuccess && rep > 0 && ms > 0.f) {
      const double gbs = (double)n_vec * 16.0 * passes / (ms * 1e-3) / 1e9;
      if (gbs > best) best = gbs;
    }
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (a) cudaEventDestroy(a);
  if (b) cudaEventDestroy(b);
  if (buf) cudaFree(buf);
  if (sink) cudaFree(sink);
  *gbs_out = best;
  return e;
}

// ---------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------
static int g_grid_cache[16] = {0};

static int grid_for(uint64_t M, int device) {
  if (device < 0 || device >= 16) device = 0;
  if (!g_grid_cache[device]) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    g_grid_cache[device] = sms;
  }
  // small launches: at least `rays_per_thread` rays per thread, so that a warp's refill keeps its lanes busy instead of 8 CTAs
  // per SM sharing a few chunks each (ECB200_RAYS_PER_THREAD, experiments)
  static const uint64_t rays_per_thread = [] { const char* e = getenv("ECB200_RAYS_PER_THREAD"); const int v = e ? atoi(e) : 1; return (uint64_t)(v < 1 ? 1 : v); }();
  const uint64_t need = (M + ECB_THREADS * rays_per_thread - 1) / (ECB_THREADS * rays_per_thread);
  const uint64_t cap = (uint64_t)g_grid_cache[device] * ECB_BLOCKS_PER_SM;
  return (int)(need < cap ? (need ? need : 1) : cap);
}

// Which closest-hit kernel a launch uses: 0 = one lane per ray (ecb_intersect_kernel), 1 = four lanes per ray
// (ecb_intersect_quad_kernel).  Both produce the same bits.  ECB200_KERNEL=lane|quad overrides the built-in choice.
static std::atomic<int> g_kernel_mode{-1};
int ecb_kernel_mode() {
  int m = g_kernel_mode.load(std::memory_order_relaxed);
  if (m < 0) {
    const char* e = getenv("ECB200_KERNEL");
    m = (e && strcmp(e, "quad") == 0) ? 1 : (e && strcmp(e, "lane") == 0) ? 0 : ECB_DEFAULT_KERNEL_MODE;
    g_kernel_mode.store(m, std::memory_order_relaxed);
  }
  return m;
}
void ecb_set_kernel_mode(int mode) { g_kernel_mode.store(mode < 0 ? -1 : (mode ? 1 : 0), std::memory_order_relaxed); }

static cudaError_t launch_quad(const EcbDeviceScene& sc, int device, const EcbLaunch& L, cudaStream_t stream) {
  if (device < 0 || device >= 16) device = 0;
  if (!g_grid_cache[device]) grid_for(1, device);
  // shared memory: quantization tables | per-quad top-level stacks ({ref, dist}, odd entry count: no bank aliasing between
  // quads) | per-quad block stacks (inner nodes only: 3*(C-1)+1 entries of {node, box} = 7 words)
  uint32_t top_entries = 3u * sc.top_depth + 2u;
  top_entries |= 1u;
  const uint32_t blk_entries = 3u * (sc.comp_level > 0 ? sc.comp_level - 1u : 0u) + 1u;
  const size_t smem = 4u * (32u + 2u * ECB_Q_QUADS * top_entries + 7u * ECB_Q_QUADS * blk_entries);
  const uint64_t need = (L.M * 4 + ECB_Q_THREADS - 1) / ECB_Q_THREADS;
  const uint64_t cap = (uint64_t)g_grid_cache[device] * ECB_Q_BLOCKS_PER_SM;
  const int grid = (int)(need < cap ? (need ? need : 1) : cap);
#define ECB_LAUNCH_Q(LEAF)                                                                                          \
  {                                                                                                                 \
    if (smem > 48u * 1024u) {                                                                                       \
      cudaError_t ae = L.counters ? cudaFuncSetAttribute(ecb_intersect_quad_kernel<LEAF, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) \
                                  : cudaFuncSetAttribute(ecb_intersect_quad_kernel<LEAF, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (ae != cudaSuccess) return ae;                                                                             \
    }                                                                                                               \
    if (L.counters)                                                                                                 \
      ecb_intersect_quad_kernel<LEAF, true><<<grid, ECB_Q_THREADS, smem, stream>>>(sc, L, top_entries, blk_entries); \
    else                                                                                                            \
      ecb_intersect_quad_kernel<LEAF, false><<<grid, ECB_Q_THREADS, smem, stream>>>(sc, L, top_entries, blk_entries); \
  }
  switch (sc.leaf_kind) {
    case 0: ECB_LAUNCH_Q(0); break;
    case 1: ECB_LAUNCH_Q(1); break;
    default: ECB_LAUNCH_Q(2); break;
  }
#undef ECB_LAUNCH_Q
  return cudaGetLastError();
}

cudaError_t ecb_launch_intersect(const EcbDeviceScene& sc, int device, const EcbLaunch& L_in, cudaStream_t stream) {
  if (L_in.M == 0) return cudaSuccess;
  EcbLaunch L = L_in;
  {  // experiments: ECB200_BLOCK_MIN / ECB200_REFILL_MIN override the batching thresholds of every launch
    static const int bm = [] { const char* e = getenv("ECB200_BLOCK_MIN"); return e ? atoi(e) : 0; }();
    static const int rm = [] { const char* e = getenv("ECB200_REFILL_MIN"); return e ? atoi(e) : 0; }();
    if (bm > 0) L.block_min = (uint32_t)bm;
    if (rm > 0) L.refill_min = (uint32_t)rm;
  }
  cudaError_t e = cudaMemsetAsync(L.queue, 0, sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return e;
  if (ecb_kernel_mode() == 1) return launch_quad(sc, device, L, stream);
  const int grid = grid_for(L.M, device);
  // deep blocks (C >= 4) take the instantiation that tests leaves inline (block_intersect); ECB200_INLINE_LEAVES=0/1 forces one
  static const int force_inline = [] { const char* e = getenv("ECB200_INLINE_LEAVES"); return e ? atoi(e) : -1; }();
  const bool deep = force_inline >= 0 ? force_inline != 0 : sc.comp_level >= 4;
#define ECB_LAUNCH(LEAF)                                                                                   \
  if (L.counters) {                                                                                        \
    if (deep) ecb_intersect_kernel<LEAF, true, true><<<grid, ECB_THREADS, 0, stream>>>(sc, L);              \
    else ecb_intersect_kernel<LEAF, true, false><<<grid, ECB_THREADS, 0, stream>>>(sc, L);                  \
  } else {                                                                                                 \
    if (deep) ecb_intersect_kernel<LEAF, false, true><<<grid, ECB_THREADS, 0, stream>>>(sc, L);             \
    else ecb_intersect_kernel<LEAF, false, false><<<grid, ECB_THREADS, 0, stream>>>(sc, L);                 \
  }
  switch (sc.leaf_kind) {
    case 0: ECB_LAUNCH(0); break;
    case 1: ECB_LAUNCH(1); break;
    default: ECB_LAUNCH(2); break;
  }
#undef ECB_LAUNCH
  return cudaGetLastError();
}

cudaError_t ecb_launch_occluded(const EcbDeviceScene& sc, int device, const EcbLaunch& L, cudaStream_t stream) {
  if (L.M == 0) return cudaSuccess;
  const int grid = grid_for(L.M, device);
  ecb_occluded_kernel<<<grid, ECB_THREADS, 0, stream>>>(sc, L);
  return cudaGetLastError();
}

// ECB200_INSTANCED_INLINE=0 forces the out-of-line object traversal (A/B measurements)
static bool instanced_inline() {
  static const bool on = [] { const char* e = getenv("ECB200_INSTANCED_INLINE"); return !(e && atoi(e) == 0); }();
  return on;
}

cudaError_t ecb_launch_instanced(const EcbDeviceInstScene& is, int device, const EcbLaunch& L, bool any_hit, cudaStream_t stream) {
  if (L.M == 0) return cudaSuccess;
  const int grid = grid_for(L.M, device);
  if (any_hit) ecb_instanced_kernel<true, false, -1><<<grid, ECB_THREADS, 0, stream>>>(is, L);
  else if (L.counters) ecb_instanced_kernel<false, true, -1><<<grid, ECB_THREADS, 0, stream>>>(is, L);
  else switch (instanced_inline() ? (int)is.uniform_leaf_kind : -1) {
    case 0: ecb_instanced_kernel<false, false, 0><<<grid, ECB_THREADS, 0, stream>>>(is, L); break;
    case 1: ecb_instanced_kernel<false, false, 1><<<grid, ECB_THREADS, 0, stream>>>(is, L); break;
    case 2: ecb_instanced_kernel<false, false, 2><<<grid, ECB_THREADS, 0, stream>>>(is, L); break;
    default: ecb_instanced_kernel<false, false, -1><<<grid, ECB_THREADS, 0, stream>>>(is, L); break;
  }
  return cudaGetLastError();
}

cudaError_t ecb_launch_triangles(const EcbDeviceTriScene& ts, int device, const EcbLaunch& L, bool any_hit, cudaStream_t stream) {
  if (L.M == 0 || ts.root == ECB_REF_EMPTY) return cudaSuccess;
  const int grid = grid_for(L.M, device);
  if (any_hit) ecb_triangle_kernel<true><<<grid, ECB_THREADS, 0, stream>>>(ts, L);
  else ecb_triangle_kernel<false><<<grid, ECB_THREADS, 0, stream>>>(ts, L);
  return cudaGetLastError();
}
