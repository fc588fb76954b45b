/* oracle/cbvh_port.c -- scalar C restatement of the reference's compressed-BVH traversal.
 * TEST INFRASTRUCTURE ONLY (see cbvh_port.h).  Compile with: gcc -O2 -ffp-contract=off -fPIC -shared -pthread
 *
 * Reference files restated (paths relative to /root/reference):
 *   kernels/bvh/bvh_intersector1.cpp:40-209       top-level BVH4 closest-hit / any-hit loops
 *   kernels/bvh/node_intersector1.h:113-139,334-368  TravRay<4,4,robust>, intersectNodeRobust
 *   kernels/bvh/bvh_traverser1.h:436-537          traverseClosestHit / traverseAnyHit (non-AVX512 path)
 *   kernels/common/stack_item.h:39-60             sort networks used by traverseClosestHit
 *   kernels/geometry/compressed.h:454-756         CompressedBVHIntersector1::intersect / occluded
 *   kernels/geometry/compressed_node.h:22-39,261-295,488-510  tables, bit layout, getNode
 *   kernels/geometry/compressed_leaf.h:93-111     leaf heights
 *   kernels/geometry/compressed_help.h:19-50,86-308  Morton, project, frustum, patch, triangles
 */
#include "cbvh_port.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ---- reference memory layout (probe-verified through oracle/refhost.cpp:ref_layout) ---- */
enum {
  BLK_GEOMID = 0, BLK_PRIMID = 4, BLK_UV = 8, BLK_RCP_EDGES = 24, BLK_NODES = 32, BLK_LEAVES = 40, BLK_GRID = 48,
  BLK_EXTENT = 56, BLK_ELEMS = 60, BLK_GRID_WIDTH = 64, BLK_SPACE = 68, BLK_PROJ = 104, BLK_BOX = 140,
  AN_CHILDREN = 0, AN_LOWER_X = 32, AN_UPPER_X = 48, AN_LOWER_Y = 64, AN_UPPER_Y = 80, AN_LOWER_Z = 96, AN_UPPER_Z = 112
};
#define REF_ALIGN_MASK 15ull      /* kernels/bvh/bvh.h:113-114 */
#define REF_TY_LEAF 8ull          /* kernels/bvh/bvh.h:124 ; emptyNode == tyLeaf (:127) */

/* kernels/geometry/compressed_node.h:22-39 ; z table is LookupTable<uni,2> = i/4 (:68-78) */
static const float TABLE_BORDER[8] = {0.000f, 0.005f, 0.010f, 0.050f, 0.100f, 0.200f, 0.400f, 0.600f};
static const float TABLE_MID[8] = {0.00f, 0.40f, 0.48f, 0.49f, 0.50f, 0.51f, 0.52f, 0.60f};
static const float TABLE_Z[4] = {0.0f, 0.25f, 0.5f, 0.75f};

static inline float ldf(const uint8_t* p) { float f; memcpy(&f, p, 4); return f; }
static inline uint32_t ldu(const uint8_t* p) { uint32_t u; memcpy(&u, p, 4); return u; }
static inline uint64_t ldq(const uint8_t* p) { uint64_t u; memcpy(&u, p, 8); return u; }
static inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float bitsf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* _mm_max_ps / _mm_min_ps semantics (second operand on NaN or equality): common/simd/vfloat4_sse2.h:310-316 */
static inline float sse_max(float a, float b) { return a > b ? a : b; }
static inline float sse_min(float a, float b) { return a < b ? a : b; }

/* common/math/vec3fa.h:163-165 / vec3.h:74-79 ; min_rcp_input = 1e-18 (constants.h:30) */
static inline float zero_fix(float a) { return fabsf(a) < 1E-18f ? 1E-18f : a; }

/* compressed_help.h:32-50 */
static inline uint32_t compact1by1(uint32_t x) {
  x &= 0x55555555u;
  x = (x ^ (x >> 1)) & 0x33333333u;
  x = (x ^ (x >> 2)) & 0x0f0f0f0fu;
  x = (x ^ (x >> 4)) & 0x00ff00ffu;
  x = (x ^ (x >> 8)) & 0x0000ffffu;
  return x;
}
static inline uint32_t morton_x(uint32_t c) { return compact1by1(c); }
static inline uint32_t morton_y(uint32_t c) { return compact1by1(c >> 1); }

/* ---- TravRay<4,4,robust=true> (node_intersector1.h:113-139,181-199) ---- */
typedef struct travray {
  float org[3], dir[3], rdir_near[3], rdir_far[3];
  int near_is_upper[3]; /* nearX = rdir>=0 ? lower : upper */
  float tnear, tfar;
} travray;

static void travray_init(travray* r, const float org[3], const float dir[3], float tnear, float tfar) {
  const float ulp3 = 1.0f + 3.0f * FLT_EPSILON;
  for (int k = 0; k < 3; k++) {
    r->org[k] = org[k];
    r->dir[k] = dir[k];
    r->rdir_near[k] = 1.0f / zero_fix(dir[k]);
    r->rdir_far[k] = r->rdir_near[k] * ulp3;
    r->near_is_upper[k] = r->rdir_near[k] >= 0.0f ? 0 : 1;
  }
  r->tnear = tnear;
  r->tfar = tfar;
}

/* 4 child boxes, SoA like BVH4::AlignedNode (bvh.h:588-594): b[0]=lower_x b[1]=upper_x b[2]=lower_y ... b[5]=upper_z */
typedef struct node4 { float b[6][4]; } node4;

/* intersectNodeRobust (node_intersector1.h:334-368): returns hit mask, writes tNear/tFar per child */
static unsigned node4_intersect(const node4* n, const travray* r, float tNear[4], float tFar[4]) {
  unsigned mask = 0;
  for (int i = 0; i < 4; i++) {
    const float tnx = (n->b[0 + r->near_is_upper[0]][i] - r->org[0]) * r->rdir_near[0];
    const float tny = (n->b[2 + r->near_is_upper[1]][i] - r->org[1]) * r->rdir_near[1];
    const float tnz = (n->b[4 + r->near_is_upper[2]][i] - r->org[2]) * r->rdir_near[2];
    const float tfx = (n->b[1 - r->near_is_upper[0]][i] - r->org[0]) * r->rdir_far[0];
    const float tfy = (n->b[3 - r->near_is_upper[1]][i] - r->org[1]) * r->rdir_far[1];
    const float tfz = (n->b[5 - r->near_is_upper[2]][i] - r->org[2]) * r->rdir_far[2];
    /* max(a,b,c,d) = max(max(a,b),max(c,d)) (common/math/math.h:201) */
    const float tn = sse_max(sse_max(tnx, tny), sse_max(tnz, r->tnear));
    const float tf = sse_min(sse_min(tfx, tfy), sse_min(tfz, r->tfar));
    tNear[i] = tn;
    tFar[i] = tf;
    if (tn <= tf) mask |= 1u << i;
  }
  return mask;
}

/* ---- block (CompressedBVH) access ---- */
typedef struct block {
  const uint8_t* hdr;
  const uint8_t* nodes;   /* 4 bytes per node: xz, x, yz, y  (compressed_node.h:264-268) */
  const uint8_t* leaves;  /* 2 bytes per cell: z12, z34      (compressed_leaf.h:24-30) */
  const uint8_t* grid;    /* 12 bytes per vertex             */
  uint32_t geomID, primID, elems, grid_width;
  float uv0x, uv0y, uvEx, uvEy, rcp_edges, extent;
  float space[3][3];      /* columns vx,vy,vz of LinearSpace3f (row index = component) : space[c][k] */
  float proj[3][3];       /* proj[r][c] */
  float box[10];
} block;

static void block_load(const cbvh_scene* s, const uint8_t* p, block* b) {
  b->hdr = p;
  b->geomID = ldu(p + BLK_GEOMID);
  b->primID = ldu(p + BLK_PRIMID);
  b->uv0x = ldf(p + BLK_UV + 0);
  b->uv0y = ldf(p + BLK_UV + 4);
  b->uvEx = ldf(p + BLK_UV + 8);
  b->uvEy = ldf(p + BLK_UV + 12);
  b->rcp_edges = ldf(p + BLK_RCP_EDGES);
  b->nodes = s->base + ldq(p + BLK_NODES);
  b->leaves = s->base + ldq(p + BLK_LEAVES);
  b->grid = s->base + ldq(p + BLK_GRID);
  b->extent = ldf(p + BLK_EXTENT);
  b->elems = ldu(p + BLK_ELEMS);
  b->grid_width = ldu(p + BLK_GRID_WIDTH);
  for (int c = 0; c < 3; c++)
    for (int k = 0; k < 3; k++) b->space[c][k] = ldf(p + BLK_SPACE + 4 * (3 * c + k));
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++)
      b->proj[r][c] = ldf(p + BLK_PROJ + 4 * (s->proj_col_major ? (3 * c + r) : (3 * r + c)));
  for (int i = 0; i < 10; i++) b->box[i] = ldf(p + BLK_BOX + 4 * i);
}

/* xfmPoint == xfmVector for LinearSpace3 (linearspace3.h:168-169): madd(x,vx,madd(y,vy,z*vz)); FMA on AVX2 */
static inline void xfm(const block* b, const float p[3], float out[3]) {
  for (int k = 0; k < 3; k++) out[k] = fmaf(p[0], b->space[0][k], fmaf(p[1], b->space[1][k], p[2] * b->space[2][k]));
}

/* project (compressed_help.h:86-90): homography on (x,y,1), z passes through */
static inline void project3(const float m[3][3], const float a[3], float out[3]) {
  const float X = m[0][0] * a[0] + m[0][1] * a[1] + m[0][2] * 1.0f;
  const float Y = m[1][0] * a[0] + m[1][1] * a[1] + m[1][2] * 1.0f;
  const float W = m[2][0] * a[0] + m[2][1] * a[1] + m[2][2] * 1.0f;
  out[0] = X / W;
  out[1] = Y / W;
  out[2] = a[2];
}

/* 3x3 inverse by adjugate/determinant (what a fixed-size Matrix3f::inverse() evaluates; used only in the
 * flat-frame special case, compressed.h:585,647) */
static void inverse3(const float a[3][3], float r[3][3]) {
  const float c00 = a[1][1] * a[2][2] - a[1][2] * a[2][1];
  const float c10 = a[0][2] * a[2][1] - a[0][1] * a[2][2];
  const float c20 = a[0][1] * a[1][2] - a[0][2] * a[1][1];
  const float det = c00 * a[0][0] + c10 * a[1][0] + c20 * a[2][0];
  const float id = 1.0f / det;
  r[0][0] = c00 * id;
  r[0][1] = c10 * id;
  r[0][2] = c20 * id;
  r[1][0] = (a[1][2] * a[2][0] - a[1][0] * a[2][2]) * id;
  r[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * id;
  r[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * id;
  r[2][0] = (a[1][0] * a[2][1] - a[1][1] * a[2][0]) * id;
  r[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * id;
  r[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * id;
}

/* intersect_line (compressed_help.h:93-106) */
static inline float intersect_line(float p2x, float p2y, float p3x, float p3y, const float o[3], const float d[3]) {
  const float vx = p2x - o[0], vy = p2y - o[1];
  const float lx = p3x - p2x, ly = p3y - p2y;
  const float t1 = (ly * vx - lx * vy) / (ly * d[0] - lx * d[1]);
  const float t2 = (d[0] * vy - d[1] * vx) / (lx * d[1] - ly * d[0]);
  return (t2 >= 0.f && t2 <= 1.f) ? t1 : NAN;
}

/* intersect_frustum (compressed_help.h:109-133) */
static int intersect_frustum(const float* box, const float o[3], const float d[3], float* t_near, float* t_far) {
  const float rdz = 1.0f / zero_fix(d[2]); /* rcp_safe; only .z is consumed */
  const float org_rdz = o[2] * rdz;
  const float t1z = box[0] * rdz - org_rdz;
  const float t2z = box[1] * rdz - org_rdz;
  const float t1x = intersect_line(box[2], box[3], box[6], box[7], o, d);
  const float t2x = intersect_line(box[4], box[5], box[8], box[9], o, d);
  const float t1y = intersect_line(box[2], box[3], box[4], box[5], o, d);
  const float t2y = intersect_line(box[6], box[7], box[8], box[9], o, d);
  const float near1 = fminf(fminf(t1x, t2x), fminf(t1y, t2y));
  const float far1 = fmaxf(fmaxf(t1x, t2x), fmaxf(t1y, t2y));
  *t_near = fmaxf(fmaxf(fminf(t1z, t2z), near1), *t_near);
  *t_far = fminf(fminf(fmaxf(t1z, t2z), far1), *t_far);
  return *t_near <= *t_far && near1 == near1 && far1 == far1;
}

/* Node::getNode (compressed_node.h:488-510) with NodeStorage<com,3,3,2> bit layout (:261-295) */
static void get_node(const uint8_t* w, const float pb[6] /* lx,ly,lz,ux,uy,uz */, node4* n) {
  const unsigned xz = w[0], x = w[1], yz = w[2], y = w[3];
  const float dimX = pb[3] - pb[0], dimY = pb[4] - pb[1], dimZ = pb[5] - pb[2];
  const float lx02 = TABLE_BORDER[(xz & 0xe0) >> 5] * dimX + pb[0];
  const float lx13 = TABLE_MID[(xz & 0x1c) >> 2] * dimX + pb[0];
  const float ux02 = (1.f - TABLE_MID[(x & 0xe0) >> 5]) * dimX + pb[0];
  const float ux13 = (1.f - TABLE_BORDER[(x & 0x1c) >> 2]) * dimX + pb[0];
  const float ly01 = TABLE_BORDER[(yz & 0xe0) >> 5] * dimY + pb[1];
  const float ly23 = TABLE_MID[(yz & 0x1c) >> 2] * dimY + pb[1];
  const float uy01 = (1.f - TABLE_MID[(y & 0xe0) >> 5]) * dimY + pb[1];
  const float uy23 = (1.f - TABLE_BORDER[(y & 0x1c) >> 2]) * dimY + pb[1];
  const float lz = TABLE_Z[xz & 3] * dimZ + pb[2];
  const float uz = (1.f - TABLE_Z[yz & 3]) * dimZ + pb[2];
  n->b[0][0] = n->b[0][2] = lx02; n->b[0][1] = n->b[0][3] = lx13;
  n->b[1][0] = n->b[1][2] = ux02; n->b[1][1] = n->b[1][3] = ux13;
  n->b[2][0] = n->b[2][1] = ly01; n->b[2][2] = n->b[2][3] = ly23;
  n->b[3][0] = n->b[3][1] = uy01; n->b[3][2] = n->b[3][3] = uy23;
  for (int i = 0; i < 4; i++) { n->b[4][i] = lz; n->b[5][i] = uz; }
}

/* intersect_patch (compressed_help.h:135-229); box = lx,ly,lz,ux,uy,uz */
static int intersect_patch(unsigned idx, float rcp_edges, float dz, float t1, float t2, float v0, float v1, float v2,
                           float v3, const float box[6], const float org[3], const float dir[3], float* u, float* v,
                           float* tt) {
  const float px = t1 * dir[0] + org[0], py = t1 * dir[1] + org[1], pz = t1 * dir[2] + org[2];
  const float p2x = t2 * dir[0] + org[0], p2y = t2 * dir[1] + org[1], p2z = t2 * dir[2] + org[2];
  const float lenX = 1.0f / (box[3] - box[0]);
  const float lenY = 1.0f / (box[4] - box[1]);
  const float fx1 = (px - box[0]) * lenX;
  const float fy1 = (py - box[1]) * lenY;
  if (t2 - t1 < 1.0E-6f) {
    *tt = t1;
    *u = (fx1 + (float)morton_x(idx)) * rcp_edges;
    *v = (fy1 + (float)morton_y(idx)) * rcp_edges;
    return 1;
  }
  const float fx2 = (p2x - box[0]) * lenX;
  const float fy2 = (p2y - box[1]) * lenY;
  const float dx1 = 1.f - fx1, dy1 = 1.f - fy1;
  float z1 = v0 * dx1 * dy1 + v1 * fx1 * dy1 + v2 * dx1 * fy1 + v3 * fx1 * fy1;
  const float dx2 = 1.f - fx2, dy2 = 1.f - fy2;
  float z2 = v0 * dx2 * dy2 + v1 * fx2 * dy2 + v2 * dx2 * fy2 + v3 * fx2 * fy2;
  if (pz >= z1 && pz <= z1 + dz) {
    *tt = t1;
    *u = (fx1 + (float)morton_x(idx)) * rcp_edges;
    *v = (fy1 + (float)morton_y(idx)) * rcp_edges;
    return 1;
  }
  if (pz > z1 + dz) { z1 += dz; z2 += dz; }
  const float alpha = p2z - z2;
  const float beta = z1 - pz;
  const float t = (t1 * alpha + t2 * beta) / (alpha + beta);
  const float d = (t - t1) / (t2 - t1);
  const float fx = fx2 - fx1, fy = fy2 - fy1;
  if (t < *tt && t >= t1 && t <= t2) {
    *u = (fx * d + fx1 + (float)morton_x(idx)) * rcp_edges;
    *v = (fy * d + fy1 + (float)morton_y(idx)) * rcp_edges;
    *tt = t;
    return 1;
  }
  return 0;
}

/* intersect_triangle (compressed_help.h:232-275) on the world-space ray */
static int intersect_triangle(const float* v0, const float* v1, const float* v2, const float rorg[3],
                              const float rdir[3], float tnear, float* tfar, float* ou, float* ov) {
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
  if (tt > 0 && tt < *tfar && tt >= tnear)
    if (beta > 0 && gamma > 0 && beta + gamma <= 1) {
      *tfar = tt;
      *ou = beta;
      *ov = gamma;
      return 1;
    }
  return 0;
}

typedef struct collector {
  cbvh_candidates* out;
} collector;

static void collect(collector* c, float t, uint32_t prim) {
  cbvh_candidates* o = c->out;
  if (t < o->t_best) {
    o->t_second = o->t_best; o->prim_second = o->prim_best;
    o->t_best = t; o->prim_best = prim;
  } else if (t < o->t_second) {
    o->t_second = t; o->prim_second = prim;
  }
}

/* CompressedBVHIntersector1::intersect (compressed.h:454-752).  `col` != NULL: candidate-collection mode
 * (never shrinks the interval; used only to compute the near-tie set). */
static void block_intersect(const cbvh_scene* s, const uint8_t* blk, cbvh_rayhit* rh, uint32_t instID,
                            cbvh_stats* st, collector* col) {
  block b;
  block_load(s, blk, &b);
  if (st) st->blocks++;
  const float rorg[3] = {rh->ray.org_x, rh->ray.org_y, rh->ray.org_z};
  const float rdir[3] = {rh->ray.dir_x, rh->ray.dir_y, rh->ray.dir_z};
  float lOrg[3], lDir[3];
  xfm(&b, rorg, lOrg);
  xfm(&b, rdir, lDir);

  float near = rh->ray.tnear;
  float far = rh->ray.tfar;
  if (!intersect_frustum(b.box, lOrg, lDir, &near, &far)) return;

  float pin[3], pout[3], org[3], tar[3], dir[3];
  for (int k = 0; k < 3; k++) pin[k] = lOrg[k] + lDir[k] * near;
  for (int k = 0; k < 3; k++) pout[k] = lOrg[k] + lDir[k] * far;
  project3(b.proj, pin, org);
  project3(b.proj, pout, tar);
  for (int k = 0; k < 3; k++) dir[k] = tar[k] - org[k];
  float zFactor = 0.f; /* the reference leaves it uninitialised in the flat-frame case; never read there for cLeaf/cBox */
  int special = 0;
  float tfar = 0.f;
  const float g_epsilon = 1.0E-4f;
  const float ax = fabsf(dir[0]), ay = fabsf(dir[1]), az = fabsf(dir[2]);
  if (ax < g_epsilon && ay < g_epsilon && az < g_epsilon) {
    dir[2] = copysignf(1.f, lDir[2]);
    org[2] -= dir[2];
    zFactor = FLT_MAX;
    tfar = FLT_MAX;
  } else if (az < g_epsilon) {
    special = 1;
    const float d2 = (dir[0] * dir[0] + dir[1] * dir[1]) + (dir[2] * dir[2] + 0.f); /* _mm_dp_ps 0x7F summation order */
    tfar = sqrtf(d2);
    const float rl = 1.0f / sqrtf(d2); /* rsqrt (math.h:86-97) */
    for (int k = 0; k < 3; k++) dir[k] = dir[k] * rl;
  } else {
    const float d2 = (dir[0] * dir[0] + dir[1] * dir[1]) + (dir[2] * dir[2] + 0.f);
    const float rl = 1.0f / sqrtf(d2);
    for (int k = 0; k < 3; k++) dir[k] = dir[k] * rl;
    zFactor = lDir[2] / dir[2];
    tfar = (rh->ray.tfar - near) * zFactor;
  }

  unsigned stack[16];
  float boxStack[16][6];
  int sp = 0;
  stack[0] = 0;
  boxStack[0][0] = -1.f; boxStack[0][1] = -1.f; boxStack[0][2] = b.box[0];
  boxStack[0][3] = 1.f;  boxStack[0][4] = 1.f;  boxStack[0][5] = b.box[1];

  travray tr;
  travray_init(&tr, org, dir, 0.f, tfar);
  float tNear[4] = {0, 0, 0, 0}, tFar[4] = {0, 0, 0, 0};

  while (sp >= 0) {
    const unsigned curr = stack[sp];
    float pBox[6];
    memcpy(pBox, boxStack[sp], sizeof(pBox));
    sp--;

    if (curr >= b.elems) {
      const unsigned idx = curr - b.elems;
      if (st) st->cells++;
      if (s->leaf_kind == CBVH_LEAF_LEAF) {
        const float box_near = tNear[idx & 3];
        if (box_near >= tfar) continue;
        const float box_far = tFar[idx & 3];
        const float dimZ = pBox[5] - pBox[2];
        const float range = (1.f + 2.f * b.extent) * dimZ;
        const float dz = 0.0625f * range;               /* getDelta() = rcp(16) (compressed_leaf.h:109-111) */
        const float offset = pBox[2] - dimZ * b.extent; /* compressed.h:559 */
        const float rcpF = 0.0625f * range;
        const unsigned z12 = b.leaves[2 * idx], z34 = b.leaves[2 * idx + 1];
        const float z1 = offset + rcpF * (float)((z12 & 0xf0) >> 4);
        const float z2 = offset + rcpF * (float)(z12 & 0x0f);
        const float z3 = offset + rcpF * (float)((z34 & 0xf0) >> 4);
        const float z4 = offset + rcpF * (float)(z34 & 0x0f);
        float u, v, t = tfar;
        if (intersect_patch(idx, b.rcp_edges, dz, box_near, box_far, z1, z2, z3, z4, pBox, org, dir, &u, &v, &t)) {
          float tw;
          if (special) {
            float ip[3][3], q[3], p[3];
            inverse3(b.proj, ip);
            for (int k = 0; k < 3; k++) q[k] = org[k] + dir[k] * t;
            project3(ip, q, p);
            const float ex = p[0] - lOrg[0], ey = p[1] - lOrg[1], ez = p[2] - lOrg[2];
            tw = sqrtf((ex * ex + ey * ey) + (ez * ez + 0.f));
          } else {
            tw = t / zFactor + near;
          }
          if (col) { collect(col, tw, b.primID); continue; }
          rh->hit.u = b.uv0x + u * b.uvEx;
          rh->hit.v = b.uv0y + v * b.uvEy;
          rh->hit.Ng_x = 1.f; rh->hit.Ng_y = 0.f; rh->hit.Ng_z = 0.f;
          rh->hit.geomID = b.geomID;
          rh->hit.primID = b.primID;
          rh->hit.instID = instID;
          tfar = t;
          rh->ray.tfar = tw;
        }
      } else if (s->leaf_kind == CBVH_LEAF_GRID) {
        const unsigned x = morton_x(idx), y = morton_y(idx);
        const unsigned g0 = y * b.grid_width + x, g1 = g0 + 1, g2 = g0 + b.grid_width, g3 = g2 + 1;
        float V[4][3];
        memcpy(V[0], b.grid + 12 * (size_t)g0, 12);
        memcpy(V[1], b.grid + 12 * (size_t)g1, 12);
        memcpy(V[2], b.grid + 12 * (size_t)g2, 12);
        memcpy(V[3], b.grid + 12 * (size_t)g3, 12);
        float tf = rh->ray.tfar, uu = 0.f, vv = 0.f;
        const float tf0 = tf;
        const int hit1 = intersect_triangle(V[0], V[1], V[2], rorg, rdir, rh->ray.tnear, &tf, &uu, &vv);
        const int hit2 = intersect_triangle(V[3], V[2], V[1], rorg, rdir, rh->ray.tnear, &tf, &uu, &vv);
        if (hit1 || hit2) {
          float lu, lv;
          if (hit2) { /* compressed_help.h:295-299 */
            lu = ((float)x + (1.f - uu)) * b.rcp_edges;
            lv = ((float)y + (1.f - vv)) * b.rcp_edges;
          } else {
            lu = ((float)x + uu) * b.rcp_edges;
            lv = ((float)y + vv) * b.rcp_edges;
          }
          if (col) { collect(col, tf, b.primID); (void)tf0; continue; }
          rh->ray.tfar = tf;
          rh->hit.Ng_x = 1.f; rh->hit.Ng_y = 0.f; rh->hit.Ng_z = 0.f;
          rh->hit.u = b.uv0x + lu * b.uvEx;
          rh->hit.v = b.uv0y + lv * b.uvEy;
          rh->hit.geomID = b.geomID;
          rh->hit.primID = b.primID;
          rh->hit.instID = instID;
          tfar = (rh->ray.tfar - near) * zFactor; /* compressed.h:608 (dead value: travRay keeps its own copy) */
        }
      } else { /* CBVH_LEAF_BOX, compressed.h:614-654 */
        const float is = tNear[idx & 3];
        if (is <= tfar) {
          const float x = (((org[0] + dir[0] * is) - pBox[0]) / (pBox[3] - pBox[0]) + (float)morton_x(idx)) * b.rcp_edges;
          const float y = (((org[1] + dir[1] * is) - pBox[1]) / (pBox[4] - pBox[1]) + (float)morton_y(idx)) * b.rcp_edges;
          float tw;
          if (special) {
            float ip[3][3], q[3], p[3];
            inverse3(b.proj, ip);
            for (int k = 0; k < 3; k++) q[k] = org[k] + dir[k] * is;
            project3(ip, q, p);
            const float ex = p[0] - lOrg[0], ey = p[1] - lOrg[1], ez = p[2] - lOrg[2];
            tw = sqrtf((ex * ex + ey * ey) + (ez * ez + 0.f));
          } else {
            tw = is / zFactor + near;
          }
          if (col) { collect(col, tw, b.primID); continue; }
          rh->hit.u = b.uv0x + x * b.uvEx;
          rh->hit.v = b.uv0y + y * b.uvEy;
          rh->hit.Ng_x = 1.f; rh->hit.Ng_y = 0.f; rh->hit.Ng_z = 0.f;
          rh->hit.geomID = b.geomID;
          rh->hit.primID = b.primID;
          rh->hit.instID = instID;
          tfar = is;
          rh->ray.tfar = tw;
        }
      }
      continue;
    }

    /* inner node: decode against the parent's decoded box, test, push hit children nearest-on-top */
    const unsigned next = curr * 4 + 1;
    node4 en;
    get_node(b.nodes + 4 * (size_t)curr, pBox, &en);
    if (st) st->cnodes++;
    unsigned mask = node4_intersect(&en, &tr, tNear, tFar);
    if (mask == 0) continue;
    unsigned c[4];
    uint32_t d[4];
    int n = 0;
    for (unsigned r = 0; r < 4; r++)
      if (mask & (1u << r)) { c[n] = r; d[n] = fbits(tNear[r]); n++; }
    int pos[4];
    if (n == 1) {
      pos[0] = sp + 1;
    } else if (n == 2) { /* compressed.h:690-697 */
      if (d[0] > d[1]) { pos[0] = sp + 1; pos[1] = sp + 2; } else { pos[1] = sp + 1; pos[0] = sp + 2; }
    } else if (n == 3) { /* compressed.h:708-710 */
      pos[0] = sp + 1 + (d[0] <= d[1]) + (d[0] <= d[2]);
      pos[1] = sp + 1 + (d[1] < d[0]) + (d[1] <= d[2]);
      pos[2] = sp + 1 + (d[2] < d[0]) + (d[2] < d[1]);
    } else { /* compressed.h:732-735 */
      pos[0] = sp + 1 + (d[0] <= d[1]) + (d[0] <= d[2]) + (d[0] <= d[3]);
      pos[1] = sp + 1 + (d[1] < d[0]) + (d[1] <= d[2]) + (d[1] <= d[3]);
      pos[2] = sp + 1 + (d[2] < d[0]) + (d[2] < d[1]) + (d[2] <= d[3]);
      pos[3] = sp + 1 + (d[3] < d[0]) + (d[3] < d[1]) + (d[3] < d[2]);
    }
    for (int i = 0; i < n; i++) {
      const unsigned r = c[i];
      float* bx = boxStack[pos[i]];
      bx[0] = en.b[0][r]; bx[1] = en.b[2][r]; bx[2] = en.b[4][r];
      bx[3] = en.b[1][r]; bx[4] = en.b[3][r]; bx[5] = en.b[5][r];
      stack[pos[i]] = next + r;
    }
    sp += n;
  }
}


/* ---- triangle meshes in the same scene -------------------------------------------------------------
 * A ROBUST scene puts its triangle meshes into a BVH4 over Triangle4v blocks (kernels/common/scene.cpp:158-164 with
 * tri_accel=bvh4.triangle4v; kernels/geometry/trianglev.h: v0,v1,v2 as Vec3vf4 SoA + geomIDs + primIDs = 176 bytes),
 * leaves = arrays of `num` blocks (NodeRef::leaf, kernels/bvh/bvh.h:380-384).  Traversal is the same robust
 * BVHNIntersector1<4,BVH_AN1,true,...> as above; the leaf test is PlueckerIntersector1<4>
 * (kernels/geometry/triangle_intersector_pluecker.h:77-127) with Intersect1EpilogM<4,4,true> / Occluded1EpilogM
 * (kernels/geometry/intersector_epilog.h:227-306).  That intersector exists for SSE4.2 / AVX / AVX512SKX only
 * (bvh4_factory.cpp:408), so an AVX2 host runs the AVX build: madd/msub are a separate multiply and add/subtract
 * (common/math/math.h), i.e. no fused operations anywhere below. */
#define TRI_BLOCK_BYTES 176
#define TRI_V0 0
#define TRI_V1 48
#define TRI_V2 96
#define TRI_GEOMIDS 144
#define TRI_PRIMIDS 160

static inline void cross3(const float a[3], const float b[3], float r[3]) { /* common/math/vec3.h:198, msub(a,b,c) = a*b-c */
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
}
static inline float dot3(const float a[3], const float b[3]) { /* vec3.h:193, madd(a.x,b.x,madd(a.y,b.y,a.z*b.z)) */
  return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]);
}
static inline void stable_triangle_normal(const float a[3], const float b[3], const float c[3], float n[3]) { /* vec3.h:200-212 */
  const float ab_x = a[2] * b[1], ab_y = a[0] * b[2], ab_z = a[1] * b[0];
  const float bc_x = b[2] * c[1], bc_y = b[0] * c[2], bc_z = b[1] * c[0];
  const float cab[3] = {a[1] * b[2] - ab_x, a[2] * b[0] - ab_y, a[0] * b[1] - ab_z};
  const float cbc[3] = {b[1] * c[2] - bc_x, b[2] * c[0] - bc_y, b[0] * c[1] - bc_z};
  n[0] = fabsf(ab_x) < fabsf(bc_x) ? cab[0] : cbc[0];
  n[1] = fabsf(ab_y) < fabsf(bc_y) ? cab[1] : cbc[1];
  n[2] = fabsf(ab_z) < fabsf(bc_z) ? cab[2] : cbc[2];
}

/* One Triangle4v block against one ray.  any_hit: Occluded1EpilogM (returns 1 as soon as a lane survives).
 * col != NULL: candidate collection (ray not updated). */
static int tri4v_intersect(const uint8_t* blk, cbvh_rayhit* rh, uint32_t instID, int any_hit, collector* col) {
  const float O[3] = {rh->ray.org_x, rh->ray.org_y, rh->ray.org_z};
  const float D[3] = {rh->ray.dir_x, rh->ray.dir_y, rh->ray.dir_z};
  float U[4], V[4], W[4], T[4], den[4], Ng[4][3];
  unsigned valid = 0;
  for (int i = 0; i < 4; i++) {
    float v0[3], v1[3], v2[3], e0[3], e1[3], e2[3], s[3], c[3];
    for (int k = 0; k < 3; k++) {
      v0[k] = ldf(blk + TRI_V0 + 16 * k + 4 * i) - O[k];
      v1[k] = ldf(blk + TRI_V1 + 16 * k + 4 * i) - O[k];
      v2[k] = ldf(blk + TRI_V2 + 16 * k + 4 * i) - O[k];
    }
    for (int k = 0; k < 3; k++) { e0[k] = v2[k] - v0[k]; e1[k] = v0[k] - v1[k]; e2[k] = v1[k] - v2[k]; }
    for (int k = 0; k < 3; k++) s[k] = v2[k] + v0[k];
    cross3(e0, s, c); U[i] = dot3(c, D);
    for (int k = 0; k < 3; k++) s[k] = v0[k] + v1[k];
    cross3(e1, s, c); V[i] = dot3(c, D);
    for (int k = 0; k < 3; k++) s[k] = v1[k] + v2[k];
    cross3(e2, s, c); W[i] = dot3(c, D);
    const float minUVW = sse_min(sse_min(U[i], V[i]), W[i]);
    const float maxUVW = sse_max(sse_max(U[i], V[i]), W[i]);
    if (minUVW >= 0.0f || maxUVW <= 0.0f) valid |= 1u << i; /* no backface culling (kernels/config.h) */
    stable_triangle_normal(e0, e1, e2, Ng[i]);
    den[i] = dot3(Ng[i], D);
    den[i] = den[i] + den[i]; /* twice() */
    T[i] = dot3(v0, Ng[i]);
    T[i] = T[i] + T[i];
  }
  if (!valid) return 0;
  for (int i = 0; i < 4; i++) {
    const float absDen = fabsf(den[i]);
    const float Ts = bitsf(fbits(T[i]) ^ (fbits(den[i]) & 0x80000000u));
    if (!(absDen * rh->ray.tnear < Ts)) valid &= ~(1u << i);
    if (!(Ts <= absDen * rh->ray.tfar)) valid &= ~(1u << i);
    if (!(den[i] != 0.0f)) valid &= ~(1u << i);
  }
  if (!valid) return 0;
  if (any_hit) return 1; /* Occluded1EpilogM without filter functions / ray masks */
  /* PlueckerHitM::finalize (rcp = rcpps + one Newton step in the reference; IEEE division here) */
  float vt[4], vu[4], vv[4];
  for (int i = 0; i < 4; i++) {
    const float rcpDen = 1.0f / den[i];
    vt[i] = T[i] * rcpDen;
    const float UVW = (U[i] + V[i]) + W[i];
    const float rcpUVW = fabsf(UVW) < 1E-18f ? 0.0f : 1.0f / UVW;
    vu[i] = U[i] * rcpUVW;
    vv[i] = V[i] * rcpUVW;
  }
  if (col) {
    for (int i = 0; i < 4; i++)
      if (valid & (1u << i)) collect(col, vt[i], ldu(blk + TRI_PRIMIDS + 4 * i));
    return 1;
  }
  /* select_min(valid, vt) (common/simd/vfloat4_sse2.h:654-659): lowest lane holding the minimum */
  float a[4], h[4], m[4];
  for (int i = 0; i < 4; i++) a[i] = (valid & (1u << i)) ? vt[i] : INFINITY;
  for (int i = 0; i < 4; i++) h[i] = sse_min(a[i ^ 1], a[i]);
  for (int i = 0; i < 4; i++) m[i] = sse_min(h[i ^ 2], h[i]);
  unsigned vmin = 0;
  for (int i = 0; i < 4; i++)
    if ((valid & (1u << i)) && a[i] == m[i]) vmin |= 1u << i;
  const unsigned pick = vmin ? vmin : valid;
  int i = 0;
  while (!(pick & (1u << i))) i++;
  rh->ray.tfar = vt[i];
  rh->hit.Ng_x = Ng[i][0]; rh->hit.Ng_y = Ng[i][1]; rh->hit.Ng_z = Ng[i][2];
  rh->hit.u = vu[i]; rh->hit.v = vv[i];
  rh->hit.primID = ldu(blk + TRI_PRIMIDS + 4 * i);
  rh->hit.geomID = ldu(blk + TRI_GEOMIDS + 4 * i);
  rh->hit.instID = instID;
  return 1;
}

/* ---- top level: BVHNIntersector1<4,BVH_AN1,robust,...>::intersect (bvh_intersector1.cpp:40-126) ---- */
typedef struct stack_item { uint64_t ref; uint32_t dist; } stack_item;
static inline void xchg(stack_item* a, stack_item* b) { stack_item t = *a; *a = *b; *b = t; }

static void load_aligned_node(const uint8_t* p, node4* n) {
  memcpy(n->b, p + AN_LOWER_X, 96); /* lower_x, upper_x, lower_y, upper_y, lower_z, upper_z */
}

/* tri != 0: the same traversal over the scene's Triangle4v BVH4 (leaves = arrays of blocks) */
static void top_intersect_ex(const cbvh_scene* s, cbvh_rayhit* rh, uint32_t instID, cbvh_stats* st, collector* col, int tri) {
  stack_item stack[256];
  int sp = 1;
  stack[0].ref = tri ? s->tri_root : s->root;
  if (stack[0].ref == 0) return; /* no such accel in this scene */
  stack[0].dist = 0xff800000u; /* neg_inf */
  const float org[3] = {rh->ray.org_x, rh->ray.org_y, rh->ray.org_z};
  const float dir[3] = {rh->ray.dir_x, rh->ray.dir_y, rh->ray.dir_z};
  travray tr;
  travray_init(&tr, org, dir, sse_max(rh->ray.tnear, 0.0f), sse_max(rh->ray.tfar, 0.0f));
  if (st) st->rays++;

  while (sp > 0) {
    sp--;
    uint64_t cur = stack[sp].ref;
    if (bitsf(stack[sp].dist) > rh->ray.tfar) continue;
    for (;;) {
      if (cur & REF_TY_LEAF) break;
      const uint8_t* np = s->base + (cur & ~REF_ALIGN_MASK);
      node4 n;
      load_aligned_node(np, &n);
      float tNear[4], tFar[4];
      if (st) st->top_nodes++;
      unsigned mask = node4_intersect(&n, &tr, tNear, tFar);
      if (mask == 0) goto pop;
      /* traverseClosestHit (bvh_traverser1.h:436-506) */
      unsigned c[4];
      int nh = 0;
      for (unsigned r = 0; r < 4; r++)
        if (mask & (1u << r)) c[nh++] = r;
      const uint64_t c0 = ldq(np + AN_CHILDREN + 8 * c[0]);
      if (nh == 1) { cur = c0; continue; }
      const uint32_t d0 = fbits(tNear[c[0]]);
      const uint64_t c1 = ldq(np + AN_CHILDREN + 8 * c[1]);
      const uint32_t d1 = fbits(tNear[c[1]]);
      if (nh == 2) {
        if (d0 < d1) { stack[sp].ref = c1; stack[sp].dist = d1; sp++; cur = c0; }
        else         { stack[sp].ref = c0; stack[sp].dist = d0; sp++; cur = c1; }
        continue;
      }
      stack[sp].ref = c0; stack[sp].dist = d0; sp++;
      stack[sp].ref = c1; stack[sp].dist = d1; sp++;
      stack[sp].ref = ldq(np + AN_CHILDREN + 8 * c[2]); stack[sp].dist = fbits(tNear[c[2]]); sp++;
      if (nh == 3) { /* sort(stackPtr[-1],stackPtr[-2],stackPtr[-3]) (stack_item.h:45-50) */
        stack_item *s1 = &stack[sp - 1], *s2 = &stack[sp - 2], *s3 = &stack[sp - 3];
        if (s2->dist < s1->dist) xchg(s2, s1);
        if (s3->dist < s2->dist) xchg(s3, s2);
        if (s2->dist < s1->dist) xchg(s2, s1);
        cur = stack[sp - 1].ref; sp--;
        continue;
      }
      stack[sp].ref = ldq(np + AN_CHILDREN + 8 * c[3]); stack[sp].dist = fbits(tNear[c[3]]); sp++;
      { /* stack_item.h:53-60 */
        stack_item *s1 = &stack[sp - 1], *s2 = &stack[sp - 2], *s3 = &stack[sp - 3], *s4 = &stack[sp - 4];
        if (s2->dist < s1->dist) xchg(s2, s1);
        if (s4->dist < s3->dist) xchg(s4, s3);
        if (s3->dist < s1->dist) xchg(s3, s1);
        if (s4->dist < s2->dist) xchg(s4, s2);
        if (s3->dist < s2->dist) xchg(s3, s2);
        cur = stack[sp - 1].ref; sp--;
      }
    }
    /* leaf: one CompressedBVH block per leaf (encodeTypedLeaf(cbvh,1), bvh_builder_subdiv.cpp:727-728) */
    if (cur == REF_TY_LEAF) continue; /* emptyNode */
    if (tri) { /* ArrayIntersector1 (kernels/geometry/intersector_iterators.h:32-36) */
      const unsigned num = (unsigned)(cur & REF_ALIGN_MASK) - (unsigned)REF_TY_LEAF;
      for (unsigned i = 0; i < num; i++)
        tri4v_intersect(s->base + (cur & ~REF_ALIGN_MASK) + (size_t)TRI_BLOCK_BYTES * i, rh, instID, 0, col);
    } else {
      block_intersect(s, s->base + (cur & ~REF_ALIGN_MASK), rh, instID, st, col);
    }
    tr.tfar = rh->ray.tfar; /* bvh_intersector1.cpp:117 */
  pop:;
  }
}

/* AccelN::intersect (kernels/common/acceln.cpp:51-56): every accel in creation order on the same ray -- triangle meshes
 * (Scene::createTriangleAccel) before subdivision surfaces (createSubdivAccel), kernels/common/scene.cpp:650-654 */
static void top_intersect(const cbvh_scene* s, cbvh_rayhit* rh, uint32_t instID, cbvh_stats* st, collector* col) {
  if (s->tri_root) top_intersect_ex(s, rh, instID, NULL, col, 1);
  top_intersect_ex(s, rh, instID, st, col, 0);
}

void cbvh_intersect1(const cbvh_scene* scene, cbvh_rayhit* rh, uint32_t instID, cbvh_stats* stats) {
  const uint32_t before = rh->hit.geomID;
  const float t_before = rh->ray.tfar;
  top_intersect(scene, rh, instID, stats, NULL);
  if (stats && (rh->hit.geomID != before || rh->ray.tfar != t_before)) stats->hits++;
}

void cbvh_candidates1(const cbvh_scene* scene, const cbvh_rayhit* rh, cbvh_candidates* out) {
  cbvh_rayhit tmp = *rh;
  out->t_best = out->t_second = INFINITY;
  out->prim_best = out->prim_second = 0xffffffffu;
  collector c = {out};
  top_intersect(scene, &tmp, 0xffffffffu, NULL, &c);
}

/* BVHNIntersector1::occluded (bvh_intersector1.cpp:128-209) with Prim::occluded == `return true`
 * (compressed.h:754-756): top-level any-hit; the first non-empty leaf reached occludes. */
static void occluded_ex(const cbvh_scene* s, cbvh_ray* ray, cbvh_stats* st, int tri) {
  if (ray->tfar < 0.0f) return;
  uint64_t stack[256];
  int sp = 1;
  stack[0] = tri ? s->tri_root : s->root;
  if (stack[0] == 0) return;
  const float org[3] = {ray->org_x, ray->org_y, ray->org_z};
  const float dir[3] = {ray->dir_x, ray->dir_y, ray->dir_z};
  travray tr;
  travray_init(&tr, org, dir, sse_max(ray->tnear, 0.0f), sse_max(ray->tfar, 0.0f));
  if (st) st->rays++;
  while (sp > 0) {
    sp--;
    uint64_t cur = stack[sp];
    for (;;) {
      if (cur & REF_TY_LEAF) break;
      const uint8_t* np = s->base + (cur & ~REF_ALIGN_MASK);
      node4 n;
      load_aligned_node(np, &n);
      float tNear[4], tFar[4];
      if (st) st->top_nodes++;
      unsigned mask = node4_intersect(&n, &tr, tNear, tFar);
      if (mask == 0) goto pop;
      /* traverseAnyHit (bvh_traverser1.h:509-537): continue with the LAST hit child, push the others in index order */
      int last = -1;
      for (unsigned r = 0; r < 4; r++)
        if (mask & (1u << r)) {
          if (last >= 0) stack[sp++] = ldq(np + AN_CHILDREN + 8 * (unsigned)last);
          last = (int)r;
        }
      cur = ldq(np + AN_CHILDREN + 8 * (unsigned)last);
    }
    if (cur == REF_TY_LEAF) continue;
    if (tri) {
      const unsigned num = (unsigned)(cur & REF_ALIGN_MASK) - (unsigned)REF_TY_LEAF;
      cbvh_rayhit tmp;
      memset(&tmp, 0, sizeof(tmp));
      tmp.ray = *ray;
      int occ = 0;
      for (unsigned i = 0; i < num && !occ; i++)
        occ = tri4v_intersect(s->base + (cur & ~REF_ALIGN_MASK) + (size_t)TRI_BLOCK_BYTES * i, &tmp, 0xffffffffu, 1, NULL);
      if (!occ) continue;
      ray->tfar = -INFINITY;
      return;
    }
    if (st) { st->blocks++; st->hits++; }
    ray->tfar = -INFINITY; /* bvh_intersector1.cpp:199 */
    return;
  pop:;
  }
}

/* AccelN::occluded (acceln.cpp:86-93): accels in order, stop once the ray is occluded */
void cbvh_occluded1(const cbvh_scene* s, cbvh_ray* ray, cbvh_stats* st) {
  if (s->tri_root) {
    occluded_ex(s, ray, NULL, 1);
    if (ray->tfar < 0.0f) { if (st) { st->rays++; st->hits++; } return; }
  }
  occluded_ex(s, ray, st, 0);
}

/* ---- instancing: BVHNIntersector1<4,BVH_AN1,robust=false,ArrayIntersector1<ObjectIntersector1>> over instances ----
 * kernels/bvh/bvh_intersector1_bvh4.cpp:81 ; non-robust TravRay (node_intersector1.h:38-55) and
 * intersectNode<4,4> AVX2/SSE4.1 path (:208-232): t = plane*rdir - org*rdir as ONE fused multiply-subtract, integer
 * max/min on the float bits (maxi/mini) and an integer compare for the hit mask. */
typedef struct travray_fast {
  float rdir[3], org_rdir[3];
  int near_is_upper[3];
  float tnear, tfar;
} travray_fast;

static inline int32_t ibits(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
static inline float maxi_f(float a, float b) { return ibits(a) > ibits(b) ? a : b; } /* _mm_max_epi32 on float bits */
static inline float mini_f(float a, float b) { return ibits(a) < ibits(b) ? a : b; }

static void travray_fast_init(travray_fast* r, const float org[3], const float dir[3], float tnear, float tfar) {
  for (int k = 0; k < 3; k++) {
    r->rdir[k] = 1.0f / zero_fix(dir[k]); /* rcp_safe: rcpps + Newton in the reference */
    r->org_rdir[k] = org[k] * r->rdir[k];
    r->near_is_upper[k] = r->rdir[k] >= 0.0f ? 0 : 1;
  }
  r->tnear = tnear;
  r->tfar = tfar;
}

static unsigned node4_intersect_fast(const node4* n, const travray_fast* r, float tNear[4]) {
  unsigned mask = 0;
  for (int i = 0; i < 4; i++) {
    const float tnx = fmaf(n->b[0 + r->near_is_upper[0]][i], r->rdir[0], -r->org_rdir[0]);
    const float tny = fmaf(n->b[2 + r->near_is_upper[1]][i], r->rdir[1], -r->org_rdir[1]);
    const float tnz = fmaf(n->b[4 + r->near_is_upper[2]][i], r->rdir[2], -r->org_rdir[2]);
    const float tfx = fmaf(n->b[1 - r->near_is_upper[0]][i], r->rdir[0], -r->org_rdir[0]);
    const float tfy = fmaf(n->b[3 - r->near_is_upper[1]][i], r->rdir[1], -r->org_rdir[1]);
    const float tfz = fmaf(n->b[5 - r->near_is_upper[2]][i], r->rdir[2], -r->org_rdir[2]);
    const float tn = maxi_f(maxi_f(tnx, tny), maxi_f(tnz, r->tnear));
    const float tf = mini_f(mini_f(tfx, tfy), mini_f(tfz, r->tfar));
    tNear[i] = tn;
    if (!(ibits(tn) > ibits(tf))) mask |= 1u << i;
  }
  return mask;
}

/* FastInstanceIntersectorN::intersect1 / occluded1 (instance_intersector.cpp:45-83): xfmPoint / xfmVector of
 * AffineSpace3fa are nested madd (common/math/affinespace.h:110-111, linearspace3.h:168-169) = FMA on AVX2 */
static void instance_ray(const cbvh_instance* in, const float o[3], const float d[3], float lo[3], float ld[3]) {
  const float* m = in->world2local;
  for (int k = 0; k < 3; k++) {
    lo[k] = fmaf(o[0], m[0 + k], fmaf(o[1], m[3 + k], fmaf(o[2], m[6 + k], m[9 + k])));
    ld[k] = fmaf(d[0], m[0 + k], fmaf(d[1], m[3 + k], d[2] * m[6 + k]));
  }
}

static void inst_traverse(const cbvh_inst_scene* s, cbvh_rayhit* rh, int any_hit, cbvh_stats* st, collector* col) {
  stack_item stack[256];
  int sp = 1;
  stack[0].ref = s->root;
  stack[0].dist = 0xff800000u;
  const float org[3] = {rh->ray.org_x, rh->ray.org_y, rh->ray.org_z};
  const float dir[3] = {rh->ray.dir_x, rh->ray.dir_y, rh->ray.dir_z};
  travray_fast tr;
  travray_fast_init(&tr, org, dir, sse_max(rh->ray.tnear, 0.0f), sse_max(rh->ray.tfar, 0.0f));
  while (sp > 0) {
    sp--;
    uint64_t cur = stack[sp].ref;
    if (!any_hit && bitsf(stack[sp].dist) > rh->ray.tfar) continue;
    for (;;) {
      if (cur & REF_TY_LEAF) break;
      const uint8_t* np = s->base + (cur & ~REF_ALIGN_MASK);
      node4 n;
      load_aligned_node(np, &n);
      float tNear[4];
      unsigned mask = node4_intersect_fast(&n, &tr, tNear);
      if (mask == 0) goto pop;
      unsigned c[4];
      int nh = 0;
      for (unsigned r = 0; r < 4; r++)
        if (mask & (1u << r)) c[nh++] = r;
      if (any_hit) { /* traverseAnyHit */
        for (int i = 0; i + 1 < nh; i++) stack[sp++].ref = ldq(np + AN_CHILDREN + 8 * c[i]);
        cur = ldq(np + AN_CHILDREN + 8 * c[nh - 1]);
        continue;
      }
      const uint64_t c0 = ldq(np + AN_CHILDREN + 8 * c[0]);
      if (nh == 1) { cur = c0; continue; }
      const uint32_t d0 = fbits(tNear[c[0]]);
      const uint64_t c1 = ldq(np + AN_CHILDREN + 8 * c[1]);
      const uint32_t d1 = fbits(tNear[c[1]]);
      if (nh == 2) {
        if (d0 < d1) { stack[sp].ref = c1; stack[sp].dist = d1; sp++; cur = c0; }
        else         { stack[sp].ref = c0; stack[sp].dist = d0; sp++; cur = c1; }
        continue;
      }
      stack[sp].ref = c0; stack[sp].dist = d0; sp++;
      stack[sp].ref = c1; stack[sp].dist = d1; sp++;
      stack[sp].ref = ldq(np + AN_CHILDREN + 8 * c[2]); stack[sp].dist = fbits(tNear[c[2]]); sp++;
      if (nh == 3) {
        stack_item *s1 = &stack[sp - 1], *s2 = &stack[sp - 2], *s3 = &stack[sp - 3];
        if (s2->dist < s1->dist) xchg(s2, s1);
        if (s3->dist < s2->dist) xchg(s3, s2);
        if (s2->dist < s1->dist) xchg(s2, s1);
        cur = stack[sp - 1].ref; sp--;
        continue;
      }
      stack[sp].ref = ldq(np + AN_CHILDREN + 8 * c[3]); stack[sp].dist = fbits(tNear[c[3]]); sp++;
      {
        stack_item *s1 = &stack[sp - 1], *s2 = &stack[sp - 2], *s3 = &stack[sp - 3], *s4 = &stack[sp - 4];
        if (s2->dist < s1->dist) xchg(s2, s1);
        if (s4->dist < s3->dist) xchg(s4, s3);
        if (s3->dist < s1->dist) xchg(s3, s1);
        if (s4->dist < s2->dist) xchg(s4, s2);
        if (s3->dist < s2->dist) xchg(s3, s2);
        cur = stack[sp - 1].ref; sp--;
      }
    }
    if (cur == REF_TY_LEAF) continue;
    { /* leaf: `num` Object primitives {geomID, primID} (ArrayIntersector1, intersector_iterators.h:27-45) */
      const unsigned num = (unsigned)((cur & REF_ALIGN_MASK) - REF_TY_LEAF);
      const uint8_t* prims = s->base + (cur & ~REF_ALIGN_MASK);
      for (unsigned i = 0; i < num; i++) {
        const uint32_t geomID = ldu(prims + 8 * i);
        if (geomID >= s->n_geoms || s->geom_to_instance[geomID] == 0xffffffffu) continue;
        const cbvh_instance* in = &s->instances[s->geom_to_instance[geomID]];
        cbvh_rayhit local = *rh;
        float lo[3], ld[3];
        instance_ray(in, org, dir, lo, ld);
        local.ray.org_x = lo[0]; local.ray.org_y = lo[1]; local.ray.org_z = lo[2];
        local.ray.dir_x = ld[0]; local.ray.dir_y = ld[1]; local.ray.dir_z = ld[2];
        if (any_hit) {
          cbvh_occluded1(&s->objects[in->object], &local.ray, st);
          if (local.ray.tfar < 0.0f) { rh->ray.tfar = -INFINITY; return; } /* object_intersector.h:61, bvh_intersector1.cpp:199 */
        } else {
          if (col) { top_intersect(&s->objects[in->object], &local, in->geomID, NULL, col); continue; }
          cbvh_intersect1(&s->objects[in->object], &local, in->geomID, st);
          rh->ray.tfar = local.ray.tfar;
          rh->hit = local.hit;
        }
      }
      tr.tfar = rh->ray.tfar;
    }
  pop:;
  }
}

void cbvh_inst_intersect1(const cbvh_inst_scene* scene, cbvh_rayhit* rh, cbvh_stats* stats) { inst_traverse(scene, rh, 0, stats, NULL); }
void cbvh_inst_candidates1(const cbvh_inst_scene* scene, const cbvh_rayhit* rh, cbvh_candidates* out) {
  cbvh_rayhit tmp = *rh;
  out->t_best = out->t_second = INFINITY;
  out->prim_best = out->prim_second = 0xffffffffu;
  collector c = {out};
  inst_traverse(scene, &tmp, 0, NULL, &c);
}
void cbvh_inst_occluded1(const cbvh_inst_scene* scene, cbvh_ray* ray, cbvh_stats* stats) {
  if (ray->tfar < 0.0f) return;
  cbvh_rayhit tmp;
  memset(&tmp, 0, sizeof(tmp));
  tmp.ray = *ray;
  inst_traverse(scene, &tmp, 1, stats, NULL);
  ray->tfar = tmp.ray.tfar;
}

/* ---- threaded drivers ---- */
typedef struct job {
  const cbvh_inst_scene* inst;
  const cbvh_scene* scene;
  char* base;
  size_t n, stride;
  int mode; /* 0 intersect, 1 occluded, 2 candidates */
  volatile size_t* next;
  cbvh_stats stats;
  cbvh_candidates* cand;
} job;

static void* worker(void* arg) {
  job* j = (job*)arg;
  const size_t chunk = 4096;
  for (;;) {
    const size_t b = __sync_fetch_and_add(j->next, chunk);
    if (b >= j->n) break;
    const size_t e = b + chunk < j->n ? b + chunk : j->n;
    for (size_t i = b; i < e; i++) {
      char* p = j->base + i * j->stride;
      if (j->mode == 0) {
        cbvh_rayhit* rh = (cbvh_rayhit*)p;
        if (rh->ray.tnear > rh->ray.tfar) continue;
        cbvh_intersect1(j->scene, rh, 0xffffffffu, &j->stats);
      } else if (j->mode == 1) {
        cbvh_ray* r = (cbvh_ray*)p;
        if (r->tnear > r->tfar) continue;
        cbvh_occluded1(j->scene, r, &j->stats);
      } else if (j->mode == 2) {
        cbvh_candidates1(j->scene, (const cbvh_rayhit*)p, &j->cand[i]);
      } else if (j->mode == 3) {
        cbvh_rayhit* rh = (cbvh_rayhit*)p;
        if (rh->ray.tnear > rh->ray.tfar) continue;
        cbvh_inst_intersect1(j->inst, rh, &j->stats);
      } else if (j->mode == 4) {
        cbvh_ray* r = (cbvh_ray*)p;
        if (r->tnear > r->tfar) continue;
        cbvh_inst_occluded1(j->inst, r, &j->stats);
      } else {
        cbvh_inst_candidates1(j->inst, (const cbvh_rayhit*)p, &j->cand[i]);
      }
    }
  }
  return NULL;
}

static double run_ex(const cbvh_scene* scene, const cbvh_inst_scene* inst, void* base, size_t n, size_t stride, int nthreads,
                     int mode, cbvh_stats* stats, cbvh_candidates* cand) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  volatile size_t next = 0;
  job* jobs = (job*)calloc((size_t)nthreads, sizeof(job));
  pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int i = 0; i < nthreads; i++) {
    jobs[i].scene = scene; jobs[i].inst = inst; jobs[i].base = (char*)base; jobs[i].n = n; jobs[i].stride = stride;
    jobs[i].mode = mode; jobs[i].next = &next; jobs[i].cand = cand;
    if (nthreads > 1) pthread_create(&th[i], NULL, worker, &jobs[i]);
  }
  if (nthreads == 1) worker(&jobs[0]);
  else for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    for (int i = 0; i < nthreads; i++) {
      stats->rays += jobs[i].stats.rays; stats->top_nodes += jobs[i].stats.top_nodes;
      stats->blocks += jobs[i].stats.blocks; stats->cnodes += jobs[i].stats.cnodes;
      stats->cells += jobs[i].stats.cells; stats->hits += jobs[i].stats.hits;
    }
  }
  free(jobs);
  free(th);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

static double run(const cbvh_scene* scene, void* base, size_t n, size_t stride, int nthreads, int mode, cbvh_stats* stats,
                  cbvh_candidates* cand) {
  return run_ex(scene, NULL, base, n, stride, nthreads, mode, stats, cand);
}
double cbvh_inst_intersectN(const cbvh_inst_scene* scene, void* rayhits, size_t n, size_t stride, int nthreads, cbvh_stats* stats) {
  return run_ex(NULL, scene, rayhits, n, stride, nthreads, 3, stats, NULL);
}
double cbvh_inst_occludedN(const cbvh_inst_scene* scene, void* rays, size_t n, size_t stride, int nthreads, cbvh_stats* stats) {
  return run_ex(NULL, scene, rays, n, stride, nthreads, 4, stats, NULL);
}
void cbvh_inst_candidatesN(const cbvh_inst_scene* scene, const void* rayhits, size_t n, size_t stride, int nthreads,
                           cbvh_candidates* out) {
  run_ex(NULL, scene, (void*)rayhits, n, stride, nthreads, 5, NULL, out);
}
double cbvh_intersectN(const cbvh_scene* scene, void* rayhits, size_t n, size_t stride, int nthreads, cbvh_stats* stats) {
  return run(scene, rayhits, n, stride, nthreads, 0, stats, NULL);
}
double cbvh_occludedN(const cbvh_scene* scene, void* rays, size_t n, size_t stride, int nthreads, cbvh_stats* stats) {
  return run(scene, rays, n, stride, nthreads, 1, stats, NULL);
}
void cbvh_candidatesN(const cbvh_scene* scene, const void* rayhits, size_t n, size_t stride, int nthreads,
                      cbvh_candidates* out) {
  run(scene, (void*)rayhits, n, stride, nthreads, 2, NULL, out);
}
