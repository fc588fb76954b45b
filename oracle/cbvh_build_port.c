/* oracle/cbvh_build_port.c -- CPU restatement of the reference's compressed-block CONSTRUCTION.
 *
 * TEST INFRASTRUCTURE ONLY (see cbvh_port.h).  Restates, for one block, everything the reference's
 *   CompressedBVH<node_t,use_grid,use_leaf,use_box>::CompressedBVH   kernels/geometry/compressed.h:49-337
 * does AFTER the vertex grid has been evaluated and the local frame chosen:
 *   homography            ComputeLinearEstimate / project            kernels/geometry/compressed_help.h:54-90
 *   monotonicity test, fall-back scaling                             compressed.h:147-210
 *   leaf boxes + bottom-up merge   build_box_hierarchy               compressed.h:381-405
 *   top-down quantization  Node::setAABB / getAABB, lookUpIdx        kernels/geometry/compressed_node.h:47-55,424-486
 *   conservative world bounds, frustum box                           compressed.h:237-292
 *   pizza-box leaves  quantTris::estimateExtent / setZ, refitTriangle kernels/geometry/compressed_leaf.h:116-236
 * Inputs are what stays on the host in this design: the displaced vertex grid (evalGrid, subdivision + displacement
 * callbacks) and the block frame `space` = inverse([vx vy vz]) with its `world` matrix (compressed.h:91-126; the frame
 * uses normalize() = rsqrtss + Newton, which no other machine reproduces bit for bit, so it is an input, not a result).
 *
 * Arithmetic: the reference builder is the AVX (non-FMA) build (bvh4_factory.cpp:375-377), every operation below is a
 * separately rounded IEEE float/double operation in the reference's order; Eigen calls follow oracle/eigen_shim/Eigen/Dense
 * (the stand-in the compiled reference is built with here: Eigen3 is un-vendored and absent).
 *
 * Pinned against: the blocks the compiled reference commits (tests/test_build_cpu.py: nodes, leaves, proj, frustum box,
 * extent byte for byte; world bounds against the parent BVH4 node within 1e-5 because `world` is re-derived there).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "cbvh_port.h"

static inline float fmin_e(float a, float b) { return a < b ? a : b; } /* embree min(float,float), common/math/math.h:177 */
static inline float fmax_e(float a, float b) { return a < b ? b : a; } /* common/math/math.h:194 */

static const float T_BORDER[8] = {0.000f, 0.005f, 0.010f, 0.050f, 0.100f, 0.200f, 0.400f, 0.600f}; /* compressed_node.h:22-39 */
static const float T_MID[8] = {0.00f, 0.40f, 0.48f, 0.49f, 0.50f, 0.51f, 0.52f, 0.60f};
static const float T_Z[4] = {0.f, 0.25f, 0.5f, 0.75f}; /* LookupTable<uni,2>: i / 4 (compressed_node.h:68-78) */

static unsigned look_up_idx(const float* table, unsigned n, float val) { /* compressed_node.h:47-55 */
  unsigned ret = 0;
  for (unsigned i = 0; i < n; i++) {
    if (table[i] <= val) ret = i;
    else break;
  }
  return ret;
}

typedef struct box3 { float lo[3], hi[3]; } box3;
static void box_empty(box3* b) { for (int k = 0; k < 3; k++) { b->lo[k] = INFINITY; b->hi[k] = -INFINITY; } }
static void box_extend_p(box3* b, const float p[3]) { for (int k = 0; k < 3; k++) { b->lo[k] = fmin_e(b->lo[k], p[k]); b->hi[k] = fmax_e(b->hi[k], p[k]); } }
static void box_extend_b(box3* b, const box3* o) { for (int k = 0; k < 3; k++) { b->lo[k] = fmin_e(b->lo[k], o->lo[k]); b->hi[k] = fmax_e(b->hi[k], o->hi[k]); } }
static void box_merge(box3* r, const box3* a, const box3* b) { for (int k = 0; k < 3; k++) { r->lo[k] = fmin_e(a->lo[k], b->lo[k]); r->hi[k] = fmax_e(a->hi[k], b->hi[k]); } }

/* xfmPoint(LinearSpace3f s, Vec3f a) = a.x*s.vx + (a.y*s.vy + a.z*s.vz)   (common/math/linearspace3.h:168); m = vx,vy,vz */
static void xfm_point(const float m[9], const float a[3], float out[3]) {
  for (int k = 0; k < 3; k++) out[k] = a[0] * m[k] + (a[1] * m[3 + k] + a[2] * m[6 + k]);
}

/* project (compressed_help.h:86-90) with Matrix3f * Vector3f of the stand-in: (m0*x + m1*y) + m2*1 ; p row-major */
static void project_p(const float p[9], const float a[3], float out[3]) {
  const float X = (p[0] * a[0] + p[1] * a[1]) + p[2] * 1.0f;
  const float Y = (p[3] * a[0] + p[4] * a[1]) + p[5] * 1.0f;
  const float W = (p[6] * a[0] + p[7] * a[1]) + p[8] * 1.0f;
  out[0] = X / W;
  out[1] = Y / W;
  out[2] = a[2];
}

static void mat3_mul(const float a[9], const float b[9], float r[9]) { /* Matrix3f::operator* of the stand-in, row-major here */
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r[3 * i + j] = (a[3 * i + 0] * b[0 + j] + a[3 * i + 1] * b[3 + j]) + a[3 * i + 2] * b[6 + j];
}

static void mat3_inverse(const float a[9], float r[9]) { /* Matrix3f::inverse of the stand-in (adjugate / determinant) */
#define A(i, j) a[3 * (i) + (j)]
  const float c00 = A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1);
  const float c10 = A(0, 2) * A(2, 1) - A(0, 1) * A(2, 2);
  const float c20 = A(0, 1) * A(1, 2) - A(0, 2) * A(1, 1);
  const float det = (c00 * A(0, 0) + c10 * A(1, 0)) + c20 * A(2, 0);
  const float id = 1.0f / det;
  r[0] = c00 * id; r[1] = c10 * id; r[2] = c20 * id;
  r[3] = (A(1, 2) * A(2, 0) - A(1, 0) * A(2, 2)) * id;
  r[4] = (A(0, 0) * A(2, 2) - A(0, 2) * A(2, 0)) * id;
  r[5] = (A(0, 2) * A(1, 0) - A(0, 0) * A(1, 2)) * id;
  r[6] = (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0)) * id;
  r[7] = (A(0, 1) * A(2, 0) - A(0, 0) * A(2, 1)) * id;
  r[8] = (A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0)) * id;
#undef A
}

/* ComputeLinearEstimate (compressed_help.h:54-84): 8x8 system in double, complete pivoting as in the stand-in's
 * FullPivLU::solve; result row-major with M(2,2) = 1. */
static void linear_estimate(const float src[4][2], const float dst[4][2], float M[9]) {
  double A[8][8], rhs[8], y[8], x[8];
  int colperm[8];
  for (int i = 0; i < 4; i++) {
    const float* p = dst[i];
    const float* q = src[i];
    const double r0[8] = {q[0], q[1], 1.f, 0.f, 0.f, 0.f, -q[0] * p[0], -q[1] * p[0]}; /* products in float, then widened */
    const double r1[8] = {0.f, 0.f, 0.f, q[0], q[1], 1.f, -q[0] * p[1], -q[1] * p[1]};
    memcpy(A[i], r0, sizeof(r0));
    memcpy(A[4 + i], r1, sizeof(r1));
  }
  for (int i = 0; i < 2; i++)
    for (int k = 0; k < 4; k++) rhs[i * 4 + k] = dst[k][i];
  const int n = 8;
  for (int i = 0; i < n; i++) colperm[i] = i;
  int rank = n;
  for (int k = 0; k < n; k++) {
    int pr = k, pc = k;
    double best = 0.0;
    for (int i = k; i < n; i++)
      for (int j = k; j < n; j++)
        if (fabs(A[i][j]) > best) { best = fabs(A[i][j]); pr = i; pc = j; }
    if (best == 0.0) { rank = k; break; }
    if (pr != k) {
      for (int j = 0; j < n; j++) { const double t = A[k][j]; A[k][j] = A[pr][j]; A[pr][j] = t; }
      const double t = rhs[k]; rhs[k] = rhs[pr]; rhs[pr] = t;
    }
    if (pc != k) {
      for (int i = 0; i < n; i++) { const double t = A[i][k]; A[i][k] = A[i][pc]; A[i][pc] = t; }
      const int t = colperm[k]; colperm[k] = colperm[pc]; colperm[pc] = t;
    }
    for (int i = k + 1; i < n; i++) {
      const double f = A[i][k] / A[k][k];
      if (f == 0.0) continue;
      for (int j = k; j < n; j++) A[i][j] -= f * A[k][j];
      rhs[i] -= f * rhs[k];
    }
  }
  for (int i = 0; i < n; i++) y[i] = 0.0;
  for (int i = rank - 1; i >= 0; i--) {
    double s = rhs[i];
    for (int j = i + 1; j < rank; j++) s -= A[i][j] * y[j];
    y[i] = s / A[i][i];
  }
  for (int i = 0; i < n; i++) x[colperm[i]] = y[i];
  for (int i = 0; i < 8; i++) M[i] = (float)x[i];
  M[8] = 1.f;
}

static uint32_t morton_x(uint32_t c);
static uint32_t morton_y(uint32_t c);
static uint32_t compact(uint32_t x) {
  x &= 0x55555555u;
  x = (x ^ (x >> 1)) & 0x33333333u;
  x = (x ^ (x >> 2)) & 0x0f0f0f0fu;
  x = (x ^ (x >> 4)) & 0x00ff00ffu;
  x = (x ^ (x >> 8)) & 0x0000ffffu;
  return x;
}
static uint32_t morton_x(uint32_t c) { return compact(c); }
static uint32_t morton_y(uint32_t c) { return compact(c >> 1); }

/* Node::setAABB (compressed_node.h:424-463): four bytes xz, x, yz, y */
static void set_aabb(uint8_t w[4], const box3* parent, const box3* b00, const box3* b10, const box3* b01, const box3* b11) {
  double xF = 1.0 / (parent->hi[0] - parent->lo[0]);
  double yF = 1.0 / (parent->hi[1] - parent->lo[1]);
  double zF = 1.0 / (parent->hi[2] - parent->lo[2]);
  if (!isfinite(xF)) xF = FLT_MIN;
  if (!isfinite(yF)) yF = FLT_MIN;
  if (!isfinite(zF)) zF = FLT_MIN;
  const float x1 = fminf(b00->lo[0], b01->lo[0]), x2 = fminf(b10->lo[0], b11->lo[0]);
  const float x3 = fmaxf(b00->hi[0], b01->hi[0]), x4 = fmaxf(b10->hi[0], b11->hi[0]);
  const float y1 = fminf(b00->lo[1], b10->lo[1]), y2 = fminf(b01->lo[1], b11->lo[1]);
  const float y3 = fmaxf(b00->hi[1], b10->hi[1]), y4 = fmaxf(b01->hi[1], b11->hi[1]);
  const float z1 = fminf(fminf(b00->lo[2], b10->lo[2]), fminf(b01->lo[2], b11->lo[2]));
  const float z2 = fmaxf(fmaxf(b00->hi[2], b10->hi[2]), fmaxf(b01->hi[2], b11->hi[2]));
  /* lookUpIdx takes `const float&`: the double product is narrowed at the call */
  const unsigned X1 = look_up_idx(T_BORDER, 8, (float)((x1 - parent->lo[0]) * xF));
  const unsigned X2 = look_up_idx(T_MID, 8, (float)((x2 - parent->lo[0]) * xF));
  const unsigned X3 = look_up_idx(T_MID, 8, (float)((parent->hi[0] - x3) * xF));
  const unsigned X4 = look_up_idx(T_BORDER, 8, (float)((parent->hi[0] - x4) * xF));
  const unsigned Y1 = look_up_idx(T_BORDER, 8, (float)((y1 - parent->lo[1]) * yF));
  const unsigned Y2 = look_up_idx(T_MID, 8, (float)((y2 - parent->lo[1]) * yF));
  const unsigned Y3 = look_up_idx(T_MID, 8, (float)((parent->hi[1] - y3) * yF));
  const unsigned Y4 = look_up_idx(T_BORDER, 8, (float)((parent->hi[1] - y4) * yF));
  const unsigned Zmin = look_up_idx(T_Z, 4, (float)((z1 - parent->lo[2]) * zF));
  const unsigned Zmax = look_up_idx(T_Z, 4, (float)((parent->hi[2] - z2) * zF));
  w[0] = (uint8_t)(X1 << 5 | X2 << 2 | Zmin); /* compressed_node.h:264-293 */
  w[1] = (uint8_t)(X3 << 5 | X4 << 2);
  w[2] = (uint8_t)(Y1 << 5 | Y2 << 2 | Zmax);
  w[3] = (uint8_t)(Y3 << 5 | Y4 << 2);
}

/* Node::getAABB (compressed_node.h:465-486): re-expanded child box `loc` */
static void get_aabb(const uint8_t w[4], const box3* parent, int loc, box3* out) {
  const unsigned X1 = w[0] >> 5, X2 = (w[0] >> 2) & 7, Zmin = w[0] & 3, X3 = w[1] >> 5, X4 = (w[1] >> 2) & 7;
  const unsigned Y1 = w[2] >> 5, Y2 = (w[2] >> 2) & 7, Zmax = w[2] & 3, Y3 = w[3] >> 5, Y4 = (w[3] >> 2) & 7;
  float mn[3], mx[3], dim[3];
  for (int k = 0; k < 3; k++) dim[k] = parent->hi[k] - parent->lo[k];
  mn[2] = T_Z[Zmin];
  mx[2] = 1.f - T_Z[Zmax];
  mn[0] = (loc & 1) ? T_MID[X2] : T_BORDER[X1];
  mx[0] = (loc & 1) ? 1.f - T_BORDER[X4] : 1.f - T_MID[X3];
  mn[1] = (loc & 2) ? T_MID[Y2] : T_BORDER[Y1];
  mx[1] = (loc & 2) ? 1.f - T_BORDER[Y4] : 1.f - T_MID[Y3];
  for (int k = 0; k < 3; k++) {
    out->lo[k] = mn[k] * dim[k] + parent->lo[k];
    out->hi[k] = mx[k] * dim[k] + parent->lo[k];
  }
}

/* baseLeaf::refitTriangle (compressed_leaf.h:116-168): height of the plane through p1,p2,p3 above (p.x, p.y) */
static float refit_triangle(float px, float py, const float p1[3], const float p2[3], const float p3[3]) {
  float tmp = p1[0], tt = p1[1], tmp2 = p1[2];
  const float a = tmp - p2[0];
  const float f = tmp - p3[0];
  const float j = tmp - px;
  const float b = tt - p2[1];
  const float e = tt - p3[1];
  const float k = tt - py;
  const float c = tmp2 - p2[2];
  tt = tmp2 - p3[2];
  const float l = tmp2 - 0.f;
  const float g = 0.f, h = 0.f;
  float gamma = 1.f;
  tmp = e * gamma - h * tt;
  tmp2 = g * tt - f * gamma;
  float M = a * tmp;
  tmp = f * h - e * g;
  M += b * tmp2;
  tmp2 = a * k - j * b;
  M += c * tmp;
  tmp = j * c - a * l;
  tt *= tmp2;
  M = 1.f / M;
  tmp2 = b * l - k * c;
  tt += e * tmp;
  tt += f * tmp2;
  tt *= -M;
  return tt;
}

static void corner_heights(const box3* pb, const float v1[3], const float v2[3], const float v3[3], const float v4[3], float z[4]) {
  z[0] = refit_triangle(pb->lo[0], pb->lo[1], v1, v2, v3);
  z[1] = refit_triangle(pb->hi[0], pb->lo[1], v1, v2, v4);
  z[2] = refit_triangle(pb->lo[0], pb->hi[1], v1, v3, v4);
  z[3] = refit_triangle(pb->hi[0], pb->hi[1], v2, v3, v4);
}

void cbvh_build_block(const cbvh_build_in* in, cbvh_build_out* out) {
  const unsigned C = (unsigned)in->comp_level;
  const unsigned width = (1u << C) + 1u, height = width;
  const unsigned cells = 1u << (2 * C);
  const float* v = in->grid;
  const float* space = in->space;
  memset(out, 0, sizeof(*out));

  /* local corners (compressed.h:129-133) and the monotonicity test + local bounds (:148-166) */
  const unsigned i00 = 0, i10 = width - 1, i01 = width * (height - 1), i11 = width * height - 1;
  float c00[3], c10[3], c01[3], c11[3];
  xfm_point(space, v + 3 * i00, c00);
  xfm_point(space, v + 3 * i10, c10);
  xfm_point(space, v + 3 * i01, c01);
  xfm_point(space, v + 3 * i11, c11);
  float source[4][2] = {{c00[0], c00[1]}, {c10[0], c10[1]}, {c01[0], c01[1]}, {c11[0], c11[1]}};
  const float target[4][2] = {{-1.f, -1.f}, {1.f, -1.f}, {-1.f, 1.f}, {1.f, 1.f}};
  int patch_ok = in->leaf_kind != CBVH_LEAF_GRID; /* `!use_grid && true` (:148) */
  box3 lBox;
  box_empty(&lBox);
  for (unsigned y = 0; y + 1 < height; y++)
    for (unsigned x = 0; x + 1 < width; x++) {
      float a00[3], a10[3], a01[3], a11[3];
      xfm_point(space, v + 3 * (y * width + x), a00);
      xfm_point(space, v + 3 * (y * width + x + 1), a10);
      xfm_point(space, v + 3 * ((y + 1) * width + x), a01);
      xfm_point(space, v + 3 * ((y + 1) * width + x + 1), a11);
      box_extend_p(&lBox, a00); box_extend_p(&lBox, a10); box_extend_p(&lBox, a01); box_extend_p(&lBox, a11);
      if (a00[0] > a10[0] || a01[0] > a11[0] || a00[1] > a01[1] || a10[1] > a11[1]) patch_ok = 0;
    }
  float proj[9];
  box3 pBox;
  box_empty(&pBox);
  if (patch_ok) { /* :171-187 */
    linear_estimate(source, target, proj);
    for (unsigned i = 0; i < width * height; i++) {
      float l[3], p[3];
      xfm_point(space, v + 3 * i, l);
      project_p(proj, l, p);
      if (!isfinite(p[0]) || !isfinite(p[1]) || p[0] < -1.5f || p[0] > 1.5f || p[1] < -1.5f || p[1] > 1.5f) { patch_ok = 0; break; }
      box_extend_p(&pBox, p);
    }
  }
  if (patch_ok) { /* :189-199 */
    float s2[4][2] = {{pBox.lo[0], pBox.lo[1]}, {pBox.hi[0], pBox.lo[1]}, {pBox.lo[0], pBox.hi[1]}, {pBox.hi[0], pBox.hi[1]}};
    float scale[9], tmp[9];
    linear_estimate(s2, target, scale);
    mat3_mul(scale, proj, tmp);
    memcpy(proj, tmp, sizeof(proj));
  } else { /* :200-210 */
    float s2[4][2] = {{lBox.lo[0], lBox.lo[1]}, {lBox.hi[0], lBox.lo[1]}, {lBox.lo[0], lBox.hi[1]}, {lBox.hi[0], lBox.hi[1]}};
    linear_estimate(s2, target, proj);
  }
  out->patch_ok = patch_ok;
  memcpy(out->proj, proj, sizeof(proj));
  float iproj[9];
  mat3_inverse(proj, iproj);

  /* build_box_hierarchy (:381-405): level 0 = cells in Morton order, then merges up to the root; stored heap-like:
   * level L (0 = root) starts at (4^L-1)/3 */
  static __thread box3 heap[(4 * 1024 * 4 - 1) / 3 + 1]; /* up to C = 6 */
  static __thread float pv[(64 + 1) * (64 + 1)][3];       /* projected vertices */
  for (unsigned i = 0; i < width * height; i++) {
    float l[3];
    xfm_point(space, v + 3 * i, l);
    project_p(proj, l, pv[i]);
  }
  const unsigned elems = (cells - 1) / 3;
  for (unsigned i = 0; i < cells; i++) {
    const unsigned x = morton_x(i), y = morton_y(i);
    box3* b = &heap[elems + i];
    box_empty(b);
    box_extend_p(b, pv[y * width + x]);
    box_extend_p(b, pv[y * width + x + 1]);
    box_extend_p(b, pv[(y + 1) * width + x]);
    box_extend_p(b, pv[(y + 1) * width + x + 1]);
  }
  for (int n = (int)elems - 1; n >= 0; n--) { /* merge(merge(a,b),merge(c,d)) (common/math/bbox.h:134-136) */
    box3 ab, cd;
    box_merge(&ab, &heap[4 * n + 1], &heap[4 * n + 2]);
    box_merge(&cd, &heap[4 * n + 3], &heap[4 * n + 4]);
    box_merge(&heap[n], &ab, &cd);
  }

  /* top-down quantization against the RE-EXPANDED parents (:226-268) */
  box3 projBox, bounds;
  box_empty(&projBox);
  box_empty(&bounds);
  const float* world = in->world;
  for (unsigned n = 0; n < elems; n++) {
    uint8_t* w = out->nodes + 4 * n;
    set_aabb(w, &heap[n], &heap[4 * n + 1], &heap[4 * n + 2], &heap[4 * n + 3], &heap[4 * n + 4]);
    for (int m = 0; m < 4; m++) {
      box3 comp;
      get_aabb(w, &heap[n], m, &comp);
      heap[4 * n + 1 + m] = comp;
      if (4 * n + 1 + m >= elems) { /* last inner level: the children are the leaf boxes */
        box3 tmp;
        box_empty(&tmp);
        for (int corner = 0; corner < 8; corner++) { /* :243-250 (order irrelevant: min/max) */
          const float p[3] = {(corner & 1) ? comp.hi[0] : comp.lo[0], (corner & 2) ? comp.hi[1] : comp.lo[1], (corner & 4) ? comp.hi[2] : comp.lo[2]};
          float q[3];
          project_p(iproj, p, q);
          box_extend_p(&tmp, q);
        }
        box_extend_b(&projBox, &comp);
        for (int corner = 0; corner < 8; corner++) { /* :256-263: the list has (u,u,u) twice and no (l,l,u) */
          if (corner == 4) continue;
          const float p[3] = {(corner & 1) ? tmp.hi[0] : tmp.lo[0], (corner & 2) ? tmp.hi[1] : tmp.lo[1], (corner & 4) ? tmp.hi[2] : tmp.lo[2]};
          float q[3];
          xfm_point(world, p, q);
          box_extend_p(&bounds, q);
        }
      }
    }
  }
  for (int k = 0; k < 3; k++) { out->bounds[k] = bounds.lo[k]; out->bounds[3 + k] = bounds.hi[k]; }

  /* frustum box (:273-292) */
  {
    const float a00[3] = {projBox.lo[0], projBox.lo[1], projBox.lo[2]}, a10[3] = {projBox.hi[0], projBox.lo[1], projBox.lo[2]};
    const float a01[3] = {projBox.lo[0], projBox.hi[1], projBox.hi[2]}, a11[3] = {projBox.hi[0], projBox.hi[1], projBox.hi[2]};
    float p00[3], p10[3], p01[3], p11[3];
    project_p(iproj, a00, p00); project_p(iproj, a10, p10); project_p(iproj, a01, p01); project_p(iproj, a11, p11);
    out->box[0] = projBox.lo[2]; out->box[1] = projBox.hi[2];
    out->box[2] = p00[0]; out->box[3] = p00[1]; out->box[4] = p10[0]; out->box[5] = p10[1];
    out->box[6] = p01[0]; out->box[7] = p01[1]; out->box[8] = p11[0]; out->box[9] = p11[1];
  }

  /* pizza-box leaves (:296-326) */
  if (in->leaf_kind == CBVH_LEAF_LEAF) {
    float extent = 0.f;
    for (unsigned i = 0; i < cells; i++) {
      const unsigned x = morton_x(i), y = morton_y(i);
      const box3* pb = &heap[elems + i];
      float z[4];
      corner_heights(pb, pv[y * width + x], pv[y * width + x + 1], pv[(y + 1) * width + x], pv[(y + 1) * width + x + 1], z);
      float t[4];
      for (int q = 0; q < 4; q++) t[q] = fmaxf(fmaxf(z[q] - pb->hi[2], 0.f), fabsf(fminf(z[q] - pb->lo[2], 0.f)));
      const double zF = pb->hi[2] - pb->lo[2];
      const float e = zF == 0.0 ? 0.f : (float)(fmaxf(fmaxf(t[0], t[1]), fmaxf(t[2], t[3])) / zF);
      extent = fmaxf(extent, e);
    }
    extent = fminf(extent, 1.0f); /* MAX_EXTENT */
    out->extent = extent;
    for (unsigned i = 0; i < cells; i++) {
      const unsigned x = morton_x(i), y = morton_y(i);
      const box3* pb = &heap[elems + i];
      float z[4];
      corner_heights(pb, pv[y * width + x], pv[y * width + x + 1], pv[(y + 1) * width + x], pv[(y + 1) * width + x + 1], z);
      const double zF = pb->hi[2] - pb->lo[2];
      unsigned code[4] = {0, 0, 0, 0};
      if (zF != 0.0) {
        /* z.setZ(z1 - (lower.z - extent*zF), ..., (1 + 2 extent) * zF): float arguments narrowed from double expressions */
        const float zf = (float)((1.0f + 2.0f * extent) * zF);
        const float rcpF = 16.f / zf;
        for (int q = 0; q < 4; q++) {
          const float rel = (float)(z[q] - (pb->lo[2] - extent * zF));
          code[q] = (unsigned)(unsigned char)fmaxf(0.f, fminf(15.f, rel * rcpF));
        }
      }
      out->leaves[2 * i] = (uint8_t)((code[0] << 4) | (code[1] & 0x0f));
      out->leaves[2 * i + 1] = (uint8_t)((code[2] << 4) | (code[3] & 0x0f));
    }
  }
}

/* many blocks, pthreads */
#include <pthread.h>
typedef struct bjob { const cbvh_build_in* in; cbvh_build_out* out; size_t n; volatile size_t* next; } bjob;
static void* bworker(void* arg) {
  bjob* j = (bjob*)arg;
  for (;;) {
    const size_t i = __sync_fetch_and_add(j->next, 1);
    if (i >= j->n) break;
    cbvh_build_block(&j->in[i], &j->out[i]);
  }
  return NULL;
}
void cbvh_build_blocks(const cbvh_build_in* in, cbvh_build_out* out, size_t n, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 64) nthreads = 64;
  volatile size_t next = 0;
  bjob j = {in, out, n, &next};
  pthread_t th[64];
  for (int i = 1; i < nthreads; i++) pthread_create(&th[i], NULL, bworker, &j);
  bworker(&j);
  for (int i = 1; i < nthreads; i++) pthread_join(th[i], NULL);
}
