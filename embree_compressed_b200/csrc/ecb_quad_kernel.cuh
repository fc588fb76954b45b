// ecb_quad_kernel.cuh -- closest-hit traversal with FOUR LANES PER RAY (included by ecb_kernels.cu inside its anonymous
// namespace; uses the helpers defined there).
//
// Why: the one-lane-per-ray kernel (ecb_intersect_kernel) runs incoherent rays at 5-8 of 32 lanes per instruction -- every
// lane walks its own path through the top-level BVH4 and fetches its own 112-byte node.  Both BVHs on this path are 4-wide
// (top level: BVH4 AlignedNode, kernels/bvh/bvh.h:403-430; block: quadtree node with four children,
// kernels/geometry/compressed_node.h:488-510), so here a quad of lanes owns one ray and lane c owns child c:
//   * a warp holds 8 rays: divergence is over 8 paths instead of 32, and a node test is one instruction stream for the quad
//     (6 plane loads + 12 mul/sub + 6 min/max per lane instead of 24 loads' worth + 80 arithmetic instructions per ray);
//   * each ray has four loads in flight per step (memory-level parallelism is what incoherent rays lack);
//   * the block set-up parallelises where the reference's arithmetic does: the four frustum edge crossings
//     (compressed_help.h:93-133) run one per lane, the two homographies (compressed_help.h:86-90) two lanes each, the three
//     reciprocal directions one per lane;
//   * the leaves of a lowest-level quadtree node are evaluated by their four lanes at once and then accepted in the
//     reference's order (nearest box first, lower child index on ties, each accepted hit shortening the ray for the next:
//     compressed.h:545-612,668-749) -- see the equivalence notes at resolve_*.
// Stacks live in shared memory (top level: 3*depth+1 entries of {ref, dist}; block: inner nodes only, 3*(C-1)+1 entries of
// {node, box}); nothing per-ray is in local memory.
//
// Per value the operations and their order are exactly those of the one-lane kernel (and therefore of the scalar oracle):
// the results are bit-identical, which tests/test_quad_kernel_gpu.py and the golden tests hold it to.
#pragma once

#ifndef ECB_Q_THREADS
#define ECB_Q_THREADS 128
#endif
#ifndef ECB_Q_MIN_BLOCKS
#define ECB_Q_MIN_BLOCKS 6
#endif
#ifndef ECB_Q_BLOCKS_PER_SM
#define ECB_Q_BLOCKS_PER_SM 8
#endif
#ifndef ECB_Q_REFILL_MIN
#define ECB_Q_REFILL_MIN 2   // idle quads (of 8) that trigger a refill
#endif
#ifndef ECB_Q_BLOCK_MIN
#define ECB_Q_BLOCK_MIN 4    // quads waiting at a block that trigger the block phase
#endif
#define ECB_Q_QUADS (ECB_Q_THREADS / 4)

#define QSHFL(v, srclane) __shfl_sync(qmask, (v), (int)(srclane))
#define QXOR(v, m) __shfl_xor_sync(qmask, (v), (m))

// world-space t of a hit in the flat-frame special case (compressed.h:585,647): un-project the hit point and measure its
// distance to the local ray origin.  Rare; out of line so that the 9-float inverse does not sit in the hot loop's registers.
__device__ __noinline__ float special_world_t(const uint8_t* blk, const float rorg[3], const float org[3], const float dir[3], float t) {
  const float* hf = reinterpret_cast<const float*>(blk);
  float space[9], proj[9], ip[9], q[3], p[3], lOrg[3];
#pragma unroll
  for (int i = 0; i < 9; i++) { space[i] = __ldg(hf + i); proj[i] = __ldg(hf + 9 + i); }
#pragma unroll
  for (int k = 0; k < 3; k++) lOrg[k] = __fmaf_rn(rorg[0], space[k], __fmaf_rn(rorg[1], space[3 + k], rorg[2] * space[6 + k]));
  inverse3(proj, ip);
#pragma unroll
  for (int k = 0; k < 3; k++) q[k] = org[k] + dir[k] * t;
  project3(ip, q, p);
  const float ex = p[0] - lOrg[0], ey = p[1] - lOrg[1], ez = p[2] - lOrg[2];
  return sqrtf((ex * ex + ey * ey) + (ez * ez + 0.f));
}

struct QuadHit {
  float tfar, u, v;
  uint32_t primID, geomID, block;
};

// One block for one quad (all four lanes call it together; every value that is not indexed by the lane is replicated).
template <int LEAF, bool STATS>
__device__ __forceinline__ void quad_block(const EcbDeviceScene& sc, const float* __restrict__ lut, uint32_t* __restrict__ bstack,
                                           const unsigned qmask, const unsigned qbase, const unsigned c, const uint32_t block_index,
                                           const float rorg[3], const float rdir[3], const float ray_tnear, QuadHit& hit, Work& work,
                                           const uint32_t blk_entries) {
  if (STATS && c == 0) work.blocks++;
  ECB_CHECK(block_index < sc.n_blocks, 2);
  const uint8_t* blk = sc.blocks + (size_t)block_index * sc.block_stride;
  const float4* h4 = reinterpret_cast<const float4*>(blk);
  const float* hf = reinterpret_cast<const float*>(blk);
  // header words (ecb_format.h): space 0..8 | proj 9..17 | box 18..27 | uv 28..31 | extent 32 | geomID 33 | primID 34 | root word 35
  const float4 s0 = __ldg(h4), s1 = __ldg(h4 + 1), s2 = __ldg(h4 + 2);
  const float space[9] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w, s2.x};
  float lOrg[3], lDir[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    lOrg[k] = __fmaf_rn(rorg[0], space[k], __fmaf_rn(rorg[1], space[3 + k], rorg[2] * space[6 + k]));
    lDir[k] = __fmaf_rn(rdir[0], space[k], __fmaf_rn(rdir[1], space[3 + k], rdir[2] * space[6 + k]));
  }
  const float4 s4 = __ldg(h4 + 4);  // proj[7], proj[8], box[0], box[1]
  const float box0 = s4.z, box1 = s4.w;

  // ---- frustum (intersect_frustum): lane 0: t1x  1: t2x  2: t1y  3: t2y; min / max trees by butterfly (fminf / fmaxf are
  // commutative, NaN-dropping) give every lane the reference's fminf(fminf(t1x,t2x), fminf(t1y,t2y)) ----
  const unsigned p2o = (c == 1) ? 22u : (c == 3) ? 24u : 20u;
  const unsigned p3o = (c == 0) ? 24u : (c == 2) ? 22u : 26u;
  const float2 p2 = __ldg(reinterpret_cast<const float2*>(hf + p2o)), p3 = __ldg(reinterpret_cast<const float2*>(hf + p3o));
  const float tl = intersect_line(p2.x, p2.y, p3.x, p3.y, lOrg, lDir);
  float near1 = fminf(tl, QXOR(tl, 1)), far1 = fmaxf(tl, QXOR(tl, 1));
  near1 = fminf(near1, QXOR(near1, 2));
  far1 = fmaxf(far1, QXOR(far1, 2));
  const float rdz = 1.0f / zero_fix(lDir[2]);
  const float org_rdz = lOrg[2] * rdz;
  const float t1z = box0 * rdz - org_rdz;
  const float t2z = box1 * rdz - org_rdz;
  const float near = fmaxf(fmaxf(fminf(t1z, t2z), near1), ray_tnear);
  const float far = fminf(fminf(fmaxf(t1z, t2z), far1), hit.tfar);
  if (!(near <= far && near1 == near1 && far1 == far1)) return;

  // ---- homographies of entry and exit point (project): lane 0: X(pin)  1: X(pout)  2: Y(pin)  3: Y(pout) ----
  float pin[3], pout[3], org[3], tar[3], dir[3];
#pragma unroll
  for (int k = 0; k < 3; k++) pin[k] = lOrg[k] + lDir[k] * near;
#pragma unroll
  for (int k = 0; k < 3; k++) pout[k] = lOrg[k] + lDir[k] * far;
  {
    const float4 s3 = __ldg(h4 + 3);  // proj[3..6]
    const float a0 = (c & 1) ? pout[0] : pin[0], a1 = (c & 1) ? pout[1] : pin[1];
    const float r0 = (c & 2) ? s3.x : s2.y, r1 = (c & 2) ? s3.y : s2.z, r2 = (c & 2) ? s3.z : s2.w;
    const float num = r0 * a0 + r1 * a1 + r2 * 1.0f;
    const float W = s3.w * a0 + s4.x * a1 + s4.y * 1.0f;
    const float q = num / W;
    org[0] = QSHFL(q, qbase + 0);
    tar[0] = QSHFL(q, qbase + 1);
    org[1] = QSHFL(q, qbase + 2);
    tar[1] = QSHFL(q, qbase + 3);
    org[2] = pin[2];
    tar[2] = pout[2];
  }
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

  // ---- TravRay of the local ray (travray_init(tr, org, dir, 0, tfar)): one reciprocal per lane ----
  float rdn[3], rdf[3];
  {
    const float dk = (c == 0) ? dir[0] : (c == 1) ? dir[1] : dir[2];
    const float rk = 1.0f / zero_fix(dk);
    const float ulp3 = 1.0f + 3.0f * FLT_EPSILON;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      rdn[k] = QSHFL(rk, qbase + k);
      rdf[k] = rdn[k] * ulp3;
    }
  }
  const bool nuX = !(rdn[0] >= 0.0f), nuY = !(rdn[1] >= 0.0f), nuZ = !(rdn[2] >= 0.0f);
  const float tr_tfar = tfar;  // the node tests keep the entry value (block_intersect never updates tr.tfar)

  const float4 s7 = __ldg(h4 + 7);  // uv0.x, uv0.y, uvE.x, uvE.y
  const float4 s8 = __ldg(h4 + 8);  // extent, geomID, primID, root word
  const float extent = s8.x;
  const uint32_t elems = sc.elems;
  const uint32_t* node_words = reinterpret_cast<const uint32_t*>(blk + ECB_BLOCK_NODES_OFFSET);

  // ---- quadtree: the stack holds inner nodes only ----
  int sp = 0;
  if (c == 0) {
    bstack[0] = 0u;
    bstack[1] = __float_as_uint(-1.f); bstack[2] = __float_as_uint(-1.f); bstack[3] = __float_as_uint(box0);
    bstack[4] = __float_as_uint(1.f);  bstack[5] = __float_as_uint(1.f);  bstack[6] = __float_as_uint(box1);
  }
  __syncwarp(qmask);
  int win = -1;          // lane of the last accepted leaf hit of this block (-1: none)
  float win_t = 0.f;     // its local t (cLeaf / cBox) or world t (cGrid)
  float win_u = 0.f, win_v = 0.f;

  while (sp >= 0) {
    const uint32_t* e = bstack + 7 * sp;
    const unsigned curr = e[0];
    const float pb0 = __uint_as_float(e[1]), pb1 = __uint_as_float(e[2]), pb2 = __uint_as_float(e[3]);
    const float pb3 = __uint_as_float(e[4]), pb4 = __uint_as_float(e[5]), pb5 = __uint_as_float(e[6]);
    sp--;
    __syncwarp(qmask);  // every lane has read the entry before a push below may overwrite it
    if (STATS && c == 0) work.cnodes++;
    ECB_CHECK(curr < elems, 4);
    const uint32_t w = curr == 0 ? __float_as_uint(s8.w) : __ldg(node_words + (curr - 1));
    const unsigned xz = w & 0xff, xb = (w >> 8) & 0xff, yz = (w >> 16) & 0xff, yb = w >> 24;
    const float dimX = pb3 - pb0, dimY = pb4 - pb1, dimZ = pb5 - pb2;
    // Node::getNode for child c: x range B for children 1,3, y range B for children 2,3 (compressed_node.h:488-510)
    const unsigned lxi = (c & 1) ? 8u + ((xz & 0x1c) >> 2) : ((xz & 0xe0) >> 5);
    const unsigned uxi = (c & 1) ? 24u + ((xb & 0x1c) >> 2) : 16u + ((xb & 0xe0) >> 5);
    const unsigned lyi = (c & 2) ? 8u + ((yz & 0x1c) >> 2) : ((yz & 0xe0) >> 5);
    const unsigned uyi = (c & 2) ? 24u + ((yb & 0x1c) >> 2) : 16u + ((yb & 0xe0) >> 5);
    const float lx = lut[lxi] * dimX + pb0, ux = lut[uxi] * dimX + pb0;
    const float ly = lut[lyi] * dimY + pb1, uy = lut[uyi] * dimY + pb1;
    const float lz = (0.25f * (float)(xz & 3)) * dimZ + pb2;
    const float uz = (1.f - 0.25f * (float)(yz & 3)) * dimZ + pb2;
    const float tnx = ((nuX ? ux : lx) - org[0]) * rdn[0], tfx = ((nuX ? lx : ux) - org[0]) * rdf[0];
    const float tny = ((nuY ? uy : ly) - org[1]) * rdn[1], tfy = ((nuY ? ly : uy) - org[1]) * rdf[1];
    const float tnz = ((nuZ ? uz : lz) - org[2]) * rdn[2], tfz = ((nuZ ? lz : uz) - org[2]) * rdf[2];
    const float tnzn = sse_max(tnz, 0.f), tfzf = sse_min(tfz, tr_tfar);
    const float tN = sse_max(sse_max(tnx, tny), tnzn), tF = sse_min(sse_min(tfx, tfy), tfzf);
    const bool chit = tN <= tF;
    const unsigned m4 = (__ballot_sync(qmask, chit) >> qbase) & 15u;
    if (m4 == 0) continue;
    const int nh = __popc(m4);
    // pop order of the hit children (compressed.h:668-749): nearer box first, lower child index first on ties
    int rank = 0;
#pragma unroll
    for (int j = 1; j < 4; j++) {
      const float tk = QXOR(tN, j);
      const unsigned k = c ^ (unsigned)j;
      const bool k_first = (k < c) ? (tk <= tN) : (tk < tN);
      rank += (((m4 >> k) & 1u) && k_first) ? 1 : 0;
    }
    const unsigned next = curr * 4 + 1;

    if (next < elems) {  // children are inner nodes: push them, nearest on top
      ECB_CHECK(sp + nh < (int)blk_entries, 3);
      if (chit) {
        uint32_t* o = bstack + 7 * (sp + 1 + (nh - 1 - rank));
        o[0] = next + c;
        o[1] = __float_as_uint(lx); o[2] = __float_as_uint(ly); o[3] = __float_as_uint(lz);
        o[4] = __float_as_uint(ux); o[5] = __float_as_uint(uy); o[6] = __float_as_uint(uz);
      }
      sp += nh;
      __syncwarp(qmask);
      continue;
    }

    // ---- children are leaves: evaluate the hit ones at once, then accept in pop order ----
    const unsigned idx = next + c - elems;
    ECB_CHECK(idx < 3u * elems + 1u, 4);
    if (STATS && c == 0) work.cells += (uint32_t)nh;
    if (LEAF == 1) {
      // intersect_patch (compressed_help.h:135-229) without its dependence on the running tfar: kind 1 = accepted whenever
      // the cell is not culled (tiny cell / entry point inside the slab, t = t1), kind 2 = accepted if t < tfar
      int kind = 0;
      float lt = 0.f, lu = 0.f, lv = 0.f;
      if (chit) {
        const float cdimZ = uz - lz;
        const float range = (1.f + 2.f * extent) * cdimZ;
        const float dz = 0.0625f * range;
        const float offset = lz - cdimZ * extent;
        const float rcpF = 0.0625f * range;
        const uint32_t zz = __ldg(reinterpret_cast<const unsigned short*>(blk + sc.leaf_offset) + idx);
        const unsigned z12 = zz & 0xff, z34 = zz >> 8;
        const float v0 = offset + rcpF * (float)((z12 & 0xf0) >> 4);
        const float v1 = offset + rcpF * (float)(z12 & 0x0f);
        const float v2 = offset + rcpF * (float)((z34 & 0xf0) >> 4);
        const float v3 = offset + rcpF * (float)(z34 & 0x0f);
        const float t1 = tN, t2 = tF;
        const float px = t1 * dir[0] + org[0], py = t1 * dir[1] + org[1], pz = t1 * dir[2] + org[2];
        const float p2x = t2 * dir[0] + org[0], p2y = t2 * dir[1] + org[1], p2z = t2 * dir[2] + org[2];
        const float lenX = 1.0f / (ux - lx);
        const float lenY = 1.0f / (uy - ly);
        const float fx1 = (px - lx) * lenX;
        const float fy1 = (py - ly) * lenY;
        const float mx = (float)compact1by1(idx), my = (float)compact1by1(idx >> 1);
        if (t2 - t1 < 1.0E-6f) {
          kind = 1; lt = t1;
          lu = (fx1 + mx) * sc.rcp_edges;
          lv = (fy1 + my) * sc.rcp_edges;
        } else {
          const float fx2 = (p2x - lx) * lenX;
          const float fy2 = (p2y - ly) * lenY;
          const float dx1 = 1.f - fx1, dy1 = 1.f - fy1;
          float z1 = v0 * dx1 * dy1 + v1 * fx1 * dy1 + v2 * dx1 * fy1 + v3 * fx1 * fy1;
          const float dx2 = 1.f - fx2, dy2 = 1.f - fy2;
          float z2 = v0 * dx2 * dy2 + v1 * fx2 * dy2 + v2 * dx2 * fy2 + v3 * fx2 * fy2;
          if (pz >= z1 && pz <= z1 + dz) {
            kind = 1; lt = t1;
            lu = (fx1 + mx) * sc.rcp_edges;
            lv = (fy1 + my) * sc.rcp_edges;
          } else {
            if (pz > z1 + dz) {
              z1 += dz;
              z2 += dz;
            }
            const float alpha = p2z - z2;
            const float beta = z1 - pz;
            const float t = (t1 * alpha + t2 * beta) / (alpha + beta);
            const float d = (t - t1) / (t2 - t1);
            const float fx = fx2 - fx1, fy = fy2 - fy1;
            if (t >= t1 && t <= t2) {
              kind = 2; lt = t;
              lu = (fx * d + fx1 + mx) * sc.rcp_edges;
              lv = (fy * d + fy1 + my) * sc.rcp_edges;
            }
          }
        }
      }
      // resolve in pop order.  Equivalent to the reference's one-leaf-at-a-time loop: a leaf is skipped when its box entry
      // is not in front of the CURRENT tfar (compressed.h:549), a kind-1 hit is then unconditional, a kind-2 hit needs
      // t < the CURRENT tfar (compressed_help.h:219); t itself does not depend on tfar.
      for (int s = 0; s < nh; s++) {
        const unsigned src = qbase + (unsigned)__ffs((__ballot_sync(qmask, chit && rank == s) >> qbase) & 15u) - 1u;
        const float bn = QSHFL(tN, src), ts = QSHFL(lt, src);
        const int kd = QSHFL(kind, src);
        if (!(bn >= tfar) && (kd == 1 || (kd == 2 && ts < tfar))) {
          tfar = ts;
          win = (int)src;
          win_t = ts;
          win_u = QSHFL(lu, src);  // uniform branch within the quad: every lane evaluates the same condition
          win_v = QSHFL(lv, src);
        }
      }
    } else if (LEAF == 0) {
      // two triangles of the cell against the WORLD ray (compressed_help.h:232-308), first against the ray's tfar at node
      // time; accepting in pop order with `tt < current tfar` reproduces the sequential result (a later cell can only
      // replace an earlier hit with a strictly nearer one, and the second triangle of a cell wins only if strictly nearer
      // than the first -- both unchanged by a shorter tfar; see DESIGN.md "quad kernel").
      bool found = false, second = false;
      float tt = hit.tfar, uu = 0.f, vv = 0.f;
      const unsigned x = compact1by1(idx), y = compact1by1(idx >> 1);
      if (chit) {
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
        const bool hit1 = intersect_triangle(V[0], V[1], V[2], rorg, rdir, ray_tnear, tt, uu, vv);
        second = intersect_triangle(V[3], V[2], V[1], rorg, rdir, ray_tnear, tt, uu, vv);
        found = hit1 || second;
      }
      for (int s = 0; s < nh; s++) {
        const unsigned src = qbase + (unsigned)__ffs((__ballot_sync(qmask, chit && rank == s) >> qbase) & 15u) - 1u;
        const float ts = QSHFL(tt, src);
        const int f = QSHFL((int)found, src);
        if (f && ts < hit.tfar) {
          hit.tfar = ts;
          win = (int)src;
          const float su = QSHFL(uu, src), sv = QSHFL(vv, src);
          const int sec = QSHFL((int)second, src);
          const unsigned sx = QSHFL(x, src), sy = QSHFL(y, src);
          float lu, lv;
          if (sec) {
            lu = ((float)sx + (1.f - su)) * sc.rcp_edges;
            lv = ((float)sy + (1.f - sv)) * sc.rcp_edges;
          } else {
            lu = ((float)sx + su) * sc.rcp_edges;
            lv = ((float)sy + sv) * sc.rcp_edges;
          }
          win_u = lu;
          win_v = lv;
        }
      }
    } else {
      // voxel leaves (compressed.h:614-654): accepted iff the box entry is not behind the running tfar
      float bx = 0.f, by = 0.f;
      if (chit) {
        const float mx = (float)compact1by1(idx), my = (float)compact1by1(idx >> 1);
        bx = (((org[0] + dir[0] * tN) - lx) / (ux - lx) + mx) * sc.rcp_edges;
        by = (((org[1] + dir[1] * tN) - ly) / (uy - ly) + my) * sc.rcp_edges;
      }
      for (int s = 0; s < nh; s++) {
        const unsigned src = qbase + (unsigned)__ffs((__ballot_sync(qmask, chit && rank == s) >> qbase) & 15u) - 1u;
        const float is = QSHFL(tN, src);
        if (is <= tfar) {
          tfar = is;
          win = (int)src;
          win_t = is;
          win_u = QSHFL(bx, src);
          win_v = QSHFL(by, src);
        }
      }
    }
  }

  if (win >= 0) {
    if (LEAF != 0) hit.tfar = special ? special_world_t(blk, rorg, org, dir, win_t) : win_t / zFactor + near;
    hit.u = s7.x + win_u * s7.z;
    hit.v = s7.y + win_v * s7.w;
    hit.geomID = __float_as_uint(s8.y);
    hit.primID = __float_as_uint(s8.z);
    hit.block = block_index;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// kernel: persistent warps of 8 quads, dynamic ray fetch with quad refill, block batching
// ---------------------------------------------------------------------------------------------------------------
template <int LEAF, bool STATS>
__global__ void __launch_bounds__(ECB_Q_THREADS, ECB_Q_MIN_BLOCKS)
ecb_intersect_quad_kernel(EcbDeviceScene sc, EcbLaunch L, uint32_t top_entries, uint32_t blk_entries) {
  extern __shared__ __align__(16) uint32_t q_smem[];
  float* lut = reinterpret_cast<float*>(q_smem);
  init_lut(lut);
  __syncthreads();
  const unsigned lane = threadIdx.x & 31u, c = lane & 3u, qbase = lane & ~3u;
  const unsigned qmask = 0xFu << qbase;
  const unsigned quad = threadIdx.x >> 2;
  uint2* tstack = reinterpret_cast<uint2*>(q_smem + 32) + (size_t)quad * top_entries;
  uint32_t* bstack = q_smem + 32 + 2u * ECB_Q_QUADS * top_entries + (size_t)quad * (7u * blk_entries);
  Work work = {0u, 0u, 0u, 0u};
  uint32_t n_rays = 0, n_hits = 0;

  uint64_t chunk_next = 0, chunk_end = 0;  // warp-private chunk of ray indices (uniform across the warp)
  bool exhausted = false;

  // quad state (replicated in its four lanes): 0 idle, 1 traversing the top level, 2 waiting at a block
  int state = 0;
  uint64_t index = 0;
  float org[3] = {0.f, 0.f, 0.f}, dir[3] = {0.f, 0.f, 1.f};
  float ray_tnear = 0.f, tr_tnear = 0.f, tr_tfar = 0.f;  // TravRay::tnear / tfar of the top level
  float rdn[3] = {1.f, 1.f, 1.f}, rdf[3] = {1.f, 1.f, 1.f};
  unsigned nu = 0;  // bit k: the near plane along axis k is the upper bound
  QuadHit hit;
  hit.tfar = 0.f; hit.u = hit.v = 0.f; hit.primID = hit.geomID = hit.block = 0xFFFFFFFFu;
  uint32_t cur = ECB_REF_EMPTY;  // node to test next; ECB_REF_EMPTY = pop the stack
  uint32_t pending = ECB_REF_EMPTY;
  int sp = 0;

  // an 8-bit mask of the warp's quads for which `pred` holds (every lane of a quad holds the same state)
#define QUADS_WHERE(pred) __popc(__ballot_sync(0xffffffffu, (pred)) & 0x11111111u)

  for (;;) {
    const unsigned idle_bits = __ballot_sync(0xffffffffu, state == 0) & 0x11111111u;
    const int n_idle0 = __popc(idle_bits);
    const int n_trav0 = QUADS_WHERE(state == 1);
    const int n_pend0 = QUADS_WHERE(state == 2);

    // ---- refill idle quads ----
    if (!exhausted && n_idle0 > 0 && (n_idle0 >= ECB_Q_REFILL_MIN || (n_trav0 == 0 && n_pend0 < ECB_Q_BLOCK_MIN))) {
      unsigned need = idle_bits;  // bit 4q set: quad q needs a ray
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
        const unsigned qrank = __popc(need & ((1u << qbase) - 1u));
        const bool mine = ((need >> qbase) & 1u) && qrank < take;
        if (mine) {
          index = chunk_next + qrank;
          ECB_CHECK(index < L.M, 7);
          float4 r0, r1;
          float tfar_in;
          load_ray(L, index, r0, r1, tfar_in);
          if (c == 0) {
            if (L.block_out) L.block_out[index] = 0xFFFFFFFFu;
            if (L.hit_geom && !L.keep_results) L.hit_geom[index] = 0xFFFFFFFFu;
          }
          if (!(r0.w > tfar_in) && sc.root != ECB_REF_EMPTY && !(L.any_hit_exact && tfar_in < 0.0f)) {  // tnear > tfar: inactive stream ray
            org[0] = r0.x; org[1] = r0.y; org[2] = r0.z;
            dir[0] = r1.x; dir[1] = r1.y; dir[2] = r1.z;
            ray_tnear = r0.w;
            hit.tfar = tfar_in;
            hit.block = 0xFFFFFFFFu;
            // travray_init: one reciprocal per lane
            const float dk = (c == 0) ? dir[0] : (c == 1) ? dir[1] : dir[2];
            const float rk = 1.0f / zero_fix(dk);
            const float ulp3 = 1.0f + 3.0f * FLT_EPSILON;
            nu = 0;
#pragma unroll
            for (int k = 0; k < 3; k++) {
              rdn[k] = QSHFL(rk, qbase + k);
              rdf[k] = rdn[k] * ulp3;
              nu |= (rdn[k] >= 0.0f ? 0u : 1u) << k;
            }
            tr_tnear = sse_max_exact(ray_tnear, 0.0f);
            tr_tfar = sse_max_exact(tfar_in, 0.0f);  // replaced by ray.tfar after every block (bvh_intersector1.cpp:117)
            cur = sc.root;
            pending = ECB_REF_EMPTY;
            sp = 0;
            state = 1;
            if (STATS && c == 0) n_rays++;
          }
        }
        chunk_next += take;
        need &= ~(__ballot_sync(0xffffffffu, mine) & 0x11111111u);
      }
    }

    // ---- top level: one node test (or one pop) per traversing quad and iteration ----
    for (;;) {
      const int n_trav = QUADS_WHERE(state == 1);
      if (n_trav == 0) break;
      if (state == 1) {
        if (cur == ECB_REF_EMPTY) {  // pop (bvh_intersector1.cpp:73-86)
          if (sp == 0) {
            // terminated: write the hit (a miss writes nothing: rtcIntersect contract)
            state = 0;
            if (c == 0 && hit.block != 0xFFFFFFFFu && !L.any_hit_exact) {
              if (L.hit_geom) {  // compact result
                L.hit_rec[index] = make_float4(hit.tfar, hit.u, hit.v, __uint_as_float(hit.primID));
                L.hit_geom[index] = hit.geomID;
                if (L.hit_ng) { L.hit_ng[3 * index] = 1.f; L.hit_ng[3 * index + 1] = 0.f; L.hit_ng[3 * index + 2] = 0.f; }
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
          } else {
            sp--;
            const uint2 e = tstack[sp];
            if (!(__uint_as_float(e.y) > hit.tfar)) cur = e.x;
          }
        }
        if (state == 1 && cur != ECB_REF_EMPTY) {
          if (cur & ECB_REF_LEAF) {
            pending = cur;
            cur = ECB_REF_EMPTY;
            state = 2;
          } else {
            // node test: lane c owns child c (intersectNodeRobust, node_intersector1.h:334-349)
            if (STATS && c == 0) work.top_nodes++;
            ECB_CHECK(cur < sc.n_top, 1);
            ECB_CHECK(sp + 3 <= (int)top_entries, 0);
            const float* nb = reinterpret_cast<const float*>(sc.top + cur);
            const unsigned ux_ = nu & 1u, uy_ = (nu >> 1) & 1u, uz_ = (nu >> 2) & 1u;
            const float pnx = __ldg(nb + 4 * ux_ + c), pfx = __ldg(nb + 4 * (1u - ux_) + c);
            const float pny = __ldg(nb + 4 * (2u + uy_) + c), pfy = __ldg(nb + 4 * (3u - uy_) + c);
            const float pnz = __ldg(nb + 4 * (4u + uz_) + c), pfz = __ldg(nb + 4 * (5u - uz_) + c);
            const uint32_t child = __ldg(reinterpret_cast<const uint32_t*>(nb) + 24 + c);
            const float tnx = (pnx - org[0]) * rdn[0], tny = (pny - org[1]) * rdn[1], tnz = (pnz - org[2]) * rdn[2];
            const float tfx = (pfx - org[0]) * rdf[0], tfy = (pfy - org[1]) * rdf[1], tfz = (pfz - org[2]) * rdf[2];
            const float tn = sse_max(sse_max(tnx, tny), sse_max(tnz, tr_tnear));
            const float tf = sse_min(sse_min(tfx, tfy), sse_min(tfz, tr_tfar));
            const bool chit = tn <= tf;
            const unsigned m4 = (__ballot_sync(qmask, chit) >> qbase) & 15u;
            if (m4 == 0) {
              cur = ECB_REF_EMPTY;
            } else {
              const int nh = __popc(m4);
              const uint32_t d = __float_as_uint(tn);
              if (nh == 1) {
                cur = QSHFL(child, qbase + (unsigned)__ffs(m4) - 1u);
              } else {
                // order of traverseClosestHit (bvh_traverser1.h:436-506): 2 and 3 hits = ascending distance (compared on the
                // float bits), HIGHER child index first on ties; 4 hits = the reference's five-exchange network, whose
                // result on ties is not a stable order, so the network itself is evaluated (stack_item.h:52-60)
                const uint32_t d0 = QSHFL(d, qbase + 0), d1 = QSHFL(d, qbase + 1), d2 = QSHFL(d, qbase + 2), d3 = QSHFL(d, qbase + 3);
                int rank;
                if (nh < 4) {
                  rank = 0;
                  const uint32_t dk[4] = {d0, d1, d2, d3};
#pragma unroll
                  for (unsigned k = 0; k < 4; k++) {
                    const bool k_first = dk[k] < d || (dk[k] == d && k > c);
                    rank += (k != c && ((m4 >> k) & 1u) && k_first) ? 1 : 0;
                  }
                } else {
                  uint32_t sd0 = d3, sd1 = d2, sd2 = d1, sd3 = d0;
                  unsigned si0 = 3, si1 = 2, si2 = 1, si3 = 0;
#define ECB_QSWAP(da, ia, db, ib)                       \
  if (db < da) {                                         \
    const uint32_t td = da; da = db; db = td;            \
    const unsigned ti = ia; ia = ib; ib = ti;            \
  }
                  ECB_QSWAP(sd0, si0, sd1, si1)
                  ECB_QSWAP(sd2, si2, sd3, si3)
                  ECB_QSWAP(sd0, si0, sd2, si2)
                  ECB_QSWAP(sd1, si1, sd3, si3)
                  ECB_QSWAP(sd1, si1, sd2, si2)
#undef ECB_QSWAP
                  rank = si0 == c ? 0 : si1 == c ? 1 : si2 == c ? 2 : 3;
                }
                const unsigned first = (unsigned)__ffs((__ballot_sync(qmask, chit && rank == 0) >> qbase) & 15u) - 1u;
                if (chit && rank > 0) tstack[sp + (nh - 1 - rank)] = make_uint2(child, d);  // farthest at the bottom
                sp += nh - 1;
                cur = QSHFL(child, qbase + first);
                __syncwarp(qmask);
              }
            }
          }
        }
      }
      // leave the loop when enough quads wait at a block, or enough are idle to make a refill worth it
      const int n_pend = QUADS_WHERE(state == 2);
      if (n_pend >= ECB_Q_BLOCK_MIN) break;
      if (!exhausted && QUADS_WHERE(state == 0) >= ECB_Q_REFILL_MIN) break;
    }

    // ---- blocks: when enough quads wait at one, or the others cannot be given work ----
    const int n_pend = QUADS_WHERE(state == 2);
    const int n_idle = QUADS_WHERE(state == 0);
    const int n_trav = QUADS_WHERE(state == 1);
    if (n_pend >= ECB_Q_BLOCK_MIN || (n_pend > 0 && (n_trav == 0 && (exhausted || n_idle < ECB_Q_REFILL_MIN)))) {
      if (state == 2) {
        quad_block<LEAF, STATS>(sc, lut, bstack, qmask, qbase, c, pending & ~ECB_REF_LEAF, org, dir, ray_tnear, hit, work, blk_entries);
        tr_tfar = hit.tfar;  // bvh_intersector1.cpp:117
        pending = ECB_REF_EMPTY;
        state = 1;  // back to the top level: pop
        if (L.any_hit_exact && hit.block != 0xFFFFFFFFu) {
          // opt-in exact shadow query (NOT the reference's behaviour): any accepted leaf hit inside [tnear, tfar] occludes
          if (c == 0) {
            if (L.hit_geom) L.hit_geom[index] = 0u;
            else *reinterpret_cast<float*>(L.out + index * L.out_stride + 32) = __int_as_float(0xff800000);
            if (STATS) n_hits++;
          }
          state = 0;
        }
      }
    } else if (n_pend == 0 && n_trav == 0 && exhausted && n_idle == 8) {
      break;
    }
  }
#undef QUADS_WHERE
  if (STATS) {
    atomicAdd(L.counters + 0, (unsigned long long)n_rays);
    atomicAdd(L.counters + 1, (unsigned long long)work.top_nodes);
    atomicAdd(L.counters + 2, (unsigned long long)work.blocks);
    atomicAdd(L.counters + 3, (unsigned long long)work.cnodes);
    atomicAdd(L.counters + 4, (unsigned long long)work.cells);
    atomicAdd(L.counters + 5, (unsigned long long)n_hits);
  }
}

#undef QSHFL
#undef QXOR
