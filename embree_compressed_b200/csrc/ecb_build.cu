// ecb_build.cu -- construction of compressed blocks on the GPU (the step in front of the traversal path).
//
// Builds blocks from their displaced vertex grids and frames, directly in the device format of ecb_format.h
// (one CTA per block for the per-vertex / per-cell / per-node work, one thread per block for the two 8x8 solves).  Restates what CompressedBVH::CompressedBVH (kernels/geometry/compressed.h:128-337) does once the grid
// has been evaluated and the frame chosen:
//   homography            ComputeLinearEstimate / project             kernels/geometry/compressed_help.h:54-90
//   monotonicity test, fall-back scaling                              compressed.h:147-210
//   leaf boxes + bottom-up merge (build_box_hierarchy)                compressed.h:381-405
//   top-down quantization against the RE-EXPANDED parents: Node::setAABB / getAABB / lookUpIdx
//                                                                     kernels/geometry/compressed_node.h:47-55,424-486
//   conservative world bounds and frustum box                         compressed.h:237-292
//   pizza-box leaves: quantTris::estimateExtent / setZ, refitTriangle kernels/geometry/compressed_leaf.h:116-236
// What stays on the host: subdivision + displacement (evalGrid with the application's callbacks) and the 3x3 frame
// (compressed.h:91-126; its normalize() is rsqrtss + Newton in the reference, not reproducible bit for bit elsewhere).
//
// Arithmetic contract: as the reference's AVX builder -- separately rounded IEEE float/double operations in the reference's
// order (-fmad=false, IEEE division); min/max reductions are exact, so their parallel order does not matter.  The 8x8
// system is solved in double with complete pivoting, the algorithm of the Eigen stand-in the compiled reference uses here
// (oracle/eigen_shim; Eigen3 is an un-vendored, unpinned dependency of the reference).  Bit-identical to the scalar oracle
// oracle/cbvh_build_port.c, which is pinned byte for byte against the blocks the compiled reference commits.
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

#include "ecb_format.h"
#include "ecb_kernels.h"

namespace {

constexpr int kBuildThreads = 128;
#ifndef ECB_BUILD_THREADS_C3
#define ECB_BUILD_THREADS_C3 32u  // CTA size for C <= 3 (81 vertices, 64 cells per block): one warp, barriers cost nothing (measured 629 -> 385 us vs 64)
#endif

__device__ __forceinline__ float min_e(float a, float b) { return a < b ? a : b; }  // embree min/max (common/math/math.h:177,194)
__device__ __forceinline__ float max_e(float a, float b) { return a < b ? b : a; }

__constant__ float c_border[8] = {0.000f, 0.005f, 0.010f, 0.050f, 0.100f, 0.200f, 0.400f, 0.600f};
__constant__ float c_mid[8] = {0.00f, 0.40f, 0.48f, 0.49f, 0.50f, 0.51f, 0.52f, 0.60f};
__constant__ float c_z[4] = {0.f, 0.25f, 0.5f, 0.75f};

// lookUpIdx (compressed_node.h:47-55): index of the last entry of the leading run with table[i] <= val, 0 if there is none.
// The tables are increasing, so that run is everything <= val and a branch-free count gives the same index (NaN: 0).
template <unsigned N>
__device__ __forceinline__ unsigned look_up_idx(const float* table, float val) {
  unsigned count = 0;
#pragma unroll
  for (unsigned i = 0; i < N; i++) count += table[i] <= val ? 1u : 0u;
  return count ? count - 1u : 0u;
}

__device__ __forceinline__ uint32_t compact(uint32_t x) {
  x &= 0x55555555u;
  x = (x ^ (x >> 1)) & 0x33333333u;
  x = (x ^ (x >> 2)) & 0x0f0f0f0fu;
  x = (x ^ (x >> 4)) & 0x00ff00ffu;
  x = (x ^ (x >> 8)) & 0x0000ffffu;
  return x;
}

struct Box3 {
  float lo[3], hi[3];
};
__device__ __forceinline__ void box_empty(Box3& b) {
#pragma unroll
  for (int k = 0; k < 3; k++) { b.lo[k] = __int_as_float(0x7f800000); b.hi[k] = __int_as_float(0xff800000); }
}
__device__ __forceinline__ void box_extend_p(Box3& b, const float p[3]) {
#pragma unroll
  for (int k = 0; k < 3; k++) { b.lo[k] = min_e(b.lo[k], p[k]); b.hi[k] = max_e(b.hi[k], p[k]); }
}
__device__ __forceinline__ void box_extend_b(Box3& b, const Box3& o) {
#pragma unroll
  for (int k = 0; k < 3; k++) { b.lo[k] = min_e(b.lo[k], o.lo[k]); b.hi[k] = max_e(b.hi[k], o.hi[k]); }
}

// xfmPoint(LinearSpace3f, Vec3f) = a.x*vx + (a.y*vy + a.z*vz)   (common/math/linearspace3.h:168, no FMA in the AVX build)
__device__ __forceinline__ void xfm_point(const float* m, const float a[3], float out[3]) {
#pragma unroll
  for (int k = 0; k < 3; k++) out[k] = a[0] * m[k] + (a[1] * m[3 + k] + a[2] * m[6 + k]);
}
// project (compressed_help.h:86-90); p row-major
__device__ __forceinline__ void project_p(const float* p, const float a[3], float out[3]) {
  const float X = (p[0] * a[0] + p[1] * a[1]) + p[2] * 1.0f;
  const float Y = (p[3] * a[0] + p[4] * a[1]) + p[5] * 1.0f;
  const float W = (p[6] * a[0] + p[7] * a[1]) + p[8] * 1.0f;
  out[0] = X / W;
  out[1] = Y / W;
  out[2] = a[2];
}

__device__ void mat3_inverse(const float* a, float* r) {
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

// ComputeLinearEstimate (compressed_help.h:54-84): maps the four source points to (+-1,+-1); one thread, double precision
__device__ void linear_estimate(const float src[4][2], float* M) {
  const float dst[4][2] = {{-1.f, -1.f}, {1.f, -1.f}, {-1.f, 1.f}, {1.f, 1.f}};
  double A[8][8], rhs[8], y[8], x[8];
  int colperm[8];
  for (int i = 0; i < 4; i++) {
    const float q0 = src[i][0], q1 = src[i][1], p0 = dst[i][0], p1 = dst[i][1];
    A[i][0] = q0; A[i][1] = q1; A[i][2] = 1.0; A[i][3] = 0.0; A[i][4] = 0.0; A[i][5] = 0.0;
    A[i][6] = (double)(-q0 * p0); A[i][7] = (double)(-q1 * p0);  // products in float, then widened
    A[4 + i][0] = 0.0; A[4 + i][1] = 0.0; A[4 + i][2] = 0.0; A[4 + i][3] = q0; A[4 + i][4] = q1; A[4 + i][5] = 1.0;
    A[4 + i][6] = (double)(-q0 * p1); A[4 + i][7] = (double)(-q1 * p1);
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

// Node::setAABB (compressed_node.h:424-463) -> the 4-byte node word xz | x<<8 | yz<<16 | y<<24
__device__ uint32_t set_aabb(const Box3& parent, const Box3& b00, const Box3& b10, const Box3& b01, const Box3& b11) {
  double xF = 1.0 / (double)(parent.hi[0] - parent.lo[0]);
  double yF = 1.0 / (double)(parent.hi[1] - parent.lo[1]);
  double zF = 1.0 / (double)(parent.hi[2] - parent.lo[2]);
  if (!isfinite(xF)) xF = FLT_MIN;
  if (!isfinite(yF)) yF = FLT_MIN;
  if (!isfinite(zF)) zF = FLT_MIN;
  const float x1 = fminf(b00.lo[0], b01.lo[0]), x2 = fminf(b10.lo[0], b11.lo[0]);
  const float x3 = fmaxf(b00.hi[0], b01.hi[0]), x4 = fmaxf(b10.hi[0], b11.hi[0]);
  const float y1 = fminf(b00.lo[1], b10.lo[1]), y2 = fminf(b01.lo[1], b11.lo[1]);
  const float y3 = fmaxf(b00.hi[1], b10.hi[1]), y4 = fmaxf(b01.hi[1], b11.hi[1]);
  const float z1 = fminf(fminf(b00.lo[2], b10.lo[2]), fminf(b01.lo[2], b11.lo[2]));
  const float z2 = fmaxf(fmaxf(b00.hi[2], b10.hi[2]), fmaxf(b01.hi[2], b11.hi[2]));
  const unsigned X1 = look_up_idx<8>(c_border, (float)((double)(x1 - parent.lo[0]) * xF));
  const unsigned X2 = look_up_idx<8>(c_mid, (float)((double)(x2 - parent.lo[0]) * xF));
  const unsigned X3 = look_up_idx<8>(c_mid, (float)((double)(parent.hi[0] - x3) * xF));
  const unsigned X4 = look_up_idx<8>(c_border, (float)((double)(parent.hi[0] - x4) * xF));
  const unsigned Y1 = look_up_idx<8>(c_border, (float)((double)(y1 - parent.lo[1]) * yF));
  const unsigned Y2 = look_up_idx<8>(c_mid, (float)((double)(y2 - parent.lo[1]) * yF));
  const unsigned Y3 = look_up_idx<8>(c_mid, (float)((double)(parent.hi[1] - y3) * yF));
  const unsigned Y4 = look_up_idx<8>(c_border, (float)((double)(parent.hi[1] - y4) * yF));
  const unsigned Zmin = look_up_idx<4>(c_z, (float)((double)(z1 - parent.lo[2]) * zF));
  const unsigned Zmax = look_up_idx<4>(c_z, (float)((double)(parent.hi[2] - z2) * zF));
  const uint32_t xz = X1 << 5 | X2 << 2 | Zmin, xb = X3 << 5 | X4 << 2, yz = Y1 << 5 | Y2 << 2 | Zmax, yb = Y3 << 5 | Y4 << 2;
  return xz | xb << 8 | yz << 16 | yb << 24;
}

// Node::getAABB (compressed_node.h:465-486)
__device__ void get_aabb(uint32_t w, const Box3& parent, int loc, Box3& out) {
  const unsigned xz = w & 0xff, xb = (w >> 8) & 0xff, yz = (w >> 16) & 0xff, yb = w >> 24;
  float mn[3], mx[3];
  mn[2] = c_z[xz & 3];
  mx[2] = 1.f - c_z[yz & 3];
  mn[0] = (loc & 1) ? c_mid[(xz >> 2) & 7] : c_border[xz >> 5];
  mx[0] = (loc & 1) ? 1.f - c_border[(xb >> 2) & 7] : 1.f - c_mid[xb >> 5];
  mn[1] = (loc & 2) ? c_mid[(yz >> 2) & 7] : c_border[yz >> 5];
  mx[1] = (loc & 2) ? 1.f - c_border[(yb >> 2) & 7] : 1.f - c_mid[yb >> 5];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float dim = parent.hi[k] - parent.lo[k];
    out.lo[k] = mn[k] * dim + parent.lo[k];
    out.hi[k] = mx[k] * dim + parent.lo[k];
  }
}

// baseLeaf::refitTriangle (compressed_leaf.h:116-168) with o = (px,py,0), d = (0,0,1)
__device__ float refit_triangle(float px, float py, const float* p1, const float* p2, const float* p3) {
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
  const float gamma = 1.f;
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

__device__ __forceinline__ void corner_heights(const Box3& pb, const float* v1, const float* v2, const float* v3, const float* v4, float z[4]) {
  z[0] = refit_triangle(pb.lo[0], pb.lo[1], v1, v2, v3);
  z[1] = refit_triangle(pb.hi[0], pb.lo[1], v1, v2, v4);
  z[2] = refit_triangle(pb.lo[0], pb.hi[1], v1, v3, v4);
  z[3] = refit_triangle(pb.hi[0], pb.hi[1], v2, v3, v4);
}

// block-wide exact min/max reduction of a box (result valid in every thread)
__device__ void reduce_box(Box3& b, float* scratch /* 6 x warps */) {
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    for (int o = 16; o > 0; o >>= 1) {
      b.lo[k] = min_e(b.lo[k], __shfl_xor_sync(0xffffffffu, b.lo[k], o));
      b.hi[k] = max_e(b.hi[k], __shfl_xor_sync(0xffffffffu, b.hi[k], o));
    }
  }
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) { scratch[6 * warp + k] = b.lo[k]; scratch[6 * warp + 3 + k] = b.hi[k]; }
  }
  __syncthreads();
  box_empty(b);
  for (unsigned w = 0; w < nw; w++) {
#pragma unroll
    for (int k = 0; k < 3; k++) { b.lo[k] = min_e(b.lo[k], scratch[6 * w + k]); b.hi[k] = max_e(b.hi[k], scratch[6 * w + 3 + k]); }
  }
  __syncthreads();
}

// Per-block scratch that carries the (tiny) serial part between the phases.
struct BuildTmp {
  float lBox[6];      // bounds of the local vertices
  float pBox[6];      // bounds of the vertices under the first homography estimate
  float corners[8];   // local xy of the four corner vertices
  float proj[9];      // first estimate, later the final homography (row-major)
  float iproj[9];
  int ok;             // the projective mapping is (still) applicable
};

struct BuildArgs {
  const float* grids;      // n x W*W x 3
  const float* space;      // n x 9
  const float* world;      // n x 9
  const float* uv;         // n x 4
  const uint32_t* geomIDs;
  const uint32_t* primIDs;
  uint8_t* blocks;         // n x stride, device format
  float* bounds;           // n x 6
  uint32_t* patch_ok;      // n (optional)
  BuildTmp* tmp;           // n
  uint32_t n, C, leaf_kind, stride, leaf_offset;
};

// The construction runs in five launches so that the two 8x8 solves per block -- serial, double precision, ~90 % of the
// instructions when one thread of a CTA does them while the others wait -- run ONE THREAD PER BLOCK across the whole
// grid instead (phases B and D); the vertex / cell / node parallel work is phases A, C and E (one CTA per block).

// local coordinates of every vertex of block `blk` into pv (nv x 3)
__device__ __forceinline__ void local_vertices(const BuildArgs& a, unsigned blk, unsigned nv, const float* s_space, float* pv) {
  const float* v = a.grids + (size_t)blk * nv * 3;
  for (unsigned i = threadIdx.x; i < nv; i += blockDim.x) {
    const float p[3] = {v[3 * i], v[3 * i + 1], v[3 * i + 2]};
    float l[3];
    xfm_point(s_space, p, l);
    pv[3 * i] = l[0]; pv[3 * i + 1] = l[1]; pv[3 * i + 2] = l[2];
  }
}

// Phase A: monotonicity test per cell and bounds of the local vertices (compressed.h:148-166)
__global__ void __launch_bounds__(kBuildThreads) ecb_build_phase_a(BuildArgs a) {
  extern __shared__ float smem[];
  const unsigned C = a.C, W = (1u << C) + 1u, nv = W * W, cells = 1u << (2 * C);
  float* pv = smem;
  float* scratch = smem + 3 * nv;
  __shared__ float s_space[9];
  __shared__ int s_flag;
  const unsigned tid = threadIdx.x, T = blockDim.x;
  for (unsigned blk = blockIdx.x; blk < a.n; blk += gridDim.x) {
    if (tid < 9) s_space[tid] = a.space[(size_t)blk * 9 + tid];
    if (tid == 0) s_flag = 0;
    __syncthreads();
    local_vertices(a, blk, nv, s_space, pv);
    __syncthreads();
    Box3 lBox;
    box_empty(lBox);
    for (unsigned i = tid; i < nv; i += T) box_extend_p(lBox, pv + 3 * i);
    bool bad = false;
    for (unsigned i = tid; i < cells; i += T) {
      const unsigned x = i % (W - 1), y = i / (W - 1);
      const float* v00 = pv + 3 * (y * W + x);
      const float* v10 = v00 + 3;
      const float* v01 = v00 + 3 * W;
      const float* v11 = v01 + 3;
      if (v00[0] > v10[0] || v01[0] > v11[0] || v00[1] > v01[1] || v10[1] > v11[1]) bad = true;
    }
    if (bad) s_flag = 1;
    reduce_box(lBox, scratch);  // its barriers publish s_flag
    if (tid == 0) {
      BuildTmp& t = a.tmp[blk];
#pragma unroll
      for (int k = 0; k < 3; k++) { t.lBox[k] = lBox.lo[k]; t.lBox[3 + k] = lBox.hi[k]; }
      const unsigned idx[4] = {0, W - 1, W * (W - 1), W * W - 1};
      for (int c = 0; c < 4; c++) { t.corners[2 * c] = pv[3 * idx[c]]; t.corners[2 * c + 1] = pv[3 * idx[c] + 1]; }
      t.ok = (a.leaf_kind != 0 /* `!use_grid && true` (:148) */ && s_flag == 0) ? 1 : 0;
    }
    __syncthreads();
  }
}

// Phase B: first homography estimate from the four corners, one thread per block (compressed.h:171)
__global__ void __launch_bounds__(128) ecb_build_phase_b(BuildArgs a) {
  const unsigned blk = blockIdx.x * blockDim.x + threadIdx.x;
  if (blk >= a.n) return;
  BuildTmp& t = a.tmp[blk];
  if (!t.ok) return;
  const float src[4][2] = {{t.corners[0], t.corners[1]}, {t.corners[2], t.corners[3]}, {t.corners[4], t.corners[5]}, {t.corners[6], t.corners[7]}};
  float M[9];
  linear_estimate(src, M);
  for (int i = 0; i < 9; i++) t.proj[i] = M[i];
}

// Phase C: the first estimate applied to every vertex: range test and projected bounds (compressed.h:173-186)
__global__ void __launch_bounds__(kBuildThreads) ecb_build_phase_c(BuildArgs a) {
  extern __shared__ float smem[];
  const unsigned C = a.C, W = (1u << C) + 1u, nv = W * W;
  float* pv = smem;
  float* scratch = smem + 3 * nv;
  __shared__ float s_space[9], s_proj[9];
  __shared__ int s_flag;
  const unsigned tid = threadIdx.x, T = blockDim.x;
  for (unsigned blk = blockIdx.x; blk < a.n; blk += gridDim.x) {
    if (!a.tmp[blk].ok) continue;  // uniform per CTA
    if (tid < 9) { s_space[tid] = a.space[(size_t)blk * 9 + tid]; s_proj[tid] = a.tmp[blk].proj[tid]; }
    if (tid == 0) s_flag = 0;
    __syncthreads();
    local_vertices(a, blk, nv, s_space, pv);
    __syncthreads();
    Box3 pBox;
    box_empty(pBox);
    bool broke = false;
    for (unsigned i = tid; i < nv; i += T) {
      float p[3];
      project_p(s_proj, pv + 3 * i, p);
      if (!isfinite(p[0]) || !isfinite(p[1]) || p[0] < -1.5f || p[0] > 1.5f || p[1] < -1.5f || p[1] > 1.5f) broke = true;
      box_extend_p(pBox, p);
    }
    if (broke) s_flag = 1;
    reduce_box(pBox, scratch);
    if (tid == 0) {
      BuildTmp& t = a.tmp[blk];
#pragma unroll
      for (int k = 0; k < 3; k++) { t.pBox[k] = pBox.lo[k]; t.pBox[3 + k] = pBox.hi[k]; }
      if (s_flag) t.ok = 0;
    }
    __syncthreads();
  }
}

// Phase D: final homography and its inverse, one thread per block (compressed.h:189-214)
__global__ void __launch_bounds__(128) ecb_build_phase_d(BuildArgs a) {
  const unsigned blk = blockIdx.x * blockDim.x + threadIdx.x;
  if (blk >= a.n) return;
  BuildTmp& t = a.tmp[blk];
  float proj[9];
  if (t.ok) {
    const float s2[4][2] = {{t.pBox[0], t.pBox[1]}, {t.pBox[3], t.pBox[1]}, {t.pBox[0], t.pBox[4]}, {t.pBox[3], t.pBox[4]}};
    float scale[9];
    linear_estimate(s2, scale);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) proj[3 * i + j] = (scale[3 * i] * t.proj[j] + scale[3 * i + 1] * t.proj[3 + j]) + scale[3 * i + 2] * t.proj[6 + j];
  } else {
    const float s2[4][2] = {{t.lBox[0], t.lBox[1]}, {t.lBox[3], t.lBox[1]}, {t.lBox[0], t.lBox[4]}, {t.lBox[3], t.lBox[4]}};
    linear_estimate(s2, proj);
  }
  float iproj[9];
  mat3_inverse(proj, iproj);
  for (int i = 0; i < 9; i++) { t.proj[i] = proj[i]; t.iproj[i] = iproj[i]; }
  if (a.patch_ok) a.patch_ok[blk] = t.ok ? 1u : 0u;
}

// Phase E: projected vertices, box hierarchy, quantization, bounds, frustum box, leaves -> the block in the device format
__global__ void __launch_bounds__(kBuildThreads) ecb_build_phase_e(BuildArgs a) {
  extern __shared__ float smem[];
  const unsigned C = a.C, W = (1u << C) + 1u, nv = W * W, cells = 1u << (2 * C), elems = (cells - 1) / 3;
  float* pv = smem;                                   // nv x 3: local, then projected vertices
  Box3* heap = reinterpret_cast<Box3*>(smem + 3 * nv);  // elems + cells boxes, heap order (children of n: 4n+1..4n+4)
  float* scratch = reinterpret_cast<float*>(heap + elems + cells);  // 6 x warps
  __shared__ float s_space[9], s_world[9], s_proj[9], s_iproj[9];
  __shared__ float s_extent;
  const unsigned tid = threadIdx.x, T = blockDim.x;

  for (unsigned blk = blockIdx.x; blk < a.n; blk += gridDim.x) {
    const float* v = a.grids + (size_t)blk * nv * 3;
    uint8_t* out = a.blocks + (size_t)blk * a.stride;
    if (tid < 9) {
      s_space[tid] = a.space[(size_t)blk * 9 + tid];
      s_world[tid] = a.world[(size_t)blk * 9 + tid];
      s_proj[tid] = a.tmp[blk].proj[tid];
      s_iproj[tid] = a.tmp[blk].iproj[tid];
    }
    __syncthreads();
    local_vertices(a, blk, nv, s_space, pv);
    __syncthreads();

    // projected vertices, leaf boxes in Morton order, bottom-up merge (build_box_hierarchy, :381-405)
    for (unsigned i = tid; i < nv; i += T) {
      float p[3];
      project_p(s_proj, pv + 3 * i, p);
      pv[3 * i] = p[0]; pv[3 * i + 1] = p[1];
    }
    __syncthreads();
    for (unsigned i = tid; i < cells; i += T) {
      const unsigned x = compact(i), y = compact(i >> 1);
      Box3 b;
      box_empty(b);
      box_extend_p(b, pv + 3 * (y * W + x));
      box_extend_p(b, pv + 3 * (y * W + x + 1));
      box_extend_p(b, pv + 3 * ((y + 1) * W + x));
      box_extend_p(b, pv + 3 * ((y + 1) * W + x + 1));
      heap[elems + i] = b;
    }
    __syncthreads();
    for (int level = (int)C - 1; level >= 0; level--) {
      const unsigned first = ((1u << (2 * level)) - 1) / 3, count = 1u << (2 * level);
      for (unsigned k = tid; k < count; k += T) {
        const unsigned n = first + k;
        Box3 ab = heap[4 * n + 1], cd = heap[4 * n + 3];
        box_extend_b(ab, heap[4 * n + 2]);  // merge(merge(a,b),merge(c,d)) (common/math/bbox.h:134-136)
        box_extend_b(cd, heap[4 * n + 4]);
        box_extend_b(ab, cd);
        heap[n] = ab;
      }
      __syncthreads();
    }

    // top-down quantization; every child is replaced by its re-expanded box before its own children are quantized (:226-268)
    Box3 projBox, bounds;
    box_empty(projBox);
    box_empty(bounds);
    uint32_t* node_words = reinterpret_cast<uint32_t*>(out + ECB_BLOCK_NODES_OFFSET);
    for (unsigned level = 0; level < C; level++) {
      const unsigned first = ((1u << (2 * level)) - 1) / 3, count = 1u << (2 * level);
      for (unsigned k = tid; k < count; k += T) {
        const unsigned n = first + k;
        const Box3 parent = heap[n];
        const uint32_t w = set_aabb(parent, heap[4 * n + 1], heap[4 * n + 2], heap[4 * n + 3], heap[4 * n + 4]);
        if (n == 0) reinterpret_cast<EcbBlockHeader*>(out)->root_word = w;
        else node_words[n - 1] = w;
        for (int m = 0; m < 4; m++) {
          Box3 comp;
          get_aabb(w, parent, m, comp);
          heap[4 * n + 1 + m] = comp;
          if (level + 1 == C) {  // the children are the leaf boxes: conservative bounds from every re-expanded leaf box
            Box3 tmp;
            box_empty(tmp);
#pragma unroll
            for (int corner = 0; corner < 8; corner++) {
              const float p[3] = {(corner & 1) ? comp.hi[0] : comp.lo[0], (corner & 2) ? comp.hi[1] : comp.lo[1], (corner & 4) ? comp.hi[2] : comp.lo[2]};
              float q[3];
              project_p(s_iproj, p, q);
              box_extend_p(tmp, q);
            }
            box_extend_b(projBox, comp);
#pragma unroll
            for (int corner = 0; corner < 8; corner++) {
              if (corner == 4) continue;  // the reference's list has (u,u,u) twice and no (l,l,u) (compressed.h:256-263)
              const float p[3] = {(corner & 1) ? tmp.hi[0] : tmp.lo[0], (corner & 2) ? tmp.hi[1] : tmp.lo[1], (corner & 4) ? tmp.hi[2] : tmp.lo[2]};
              float q[3];
              xfm_point(s_world, p, q);
              box_extend_p(bounds, q);
            }
          }
        }
      }
      __syncthreads();
    }
    reduce_box(projBox, scratch);
    reduce_box(bounds, scratch);

    // header: frame, homography, frustum box (:273-292), uv, ids
    EcbBlockHeader* hdr = reinterpret_cast<EcbBlockHeader*>(out);
    if (tid < 9) { hdr->space[tid] = s_space[tid]; hdr->proj[tid] = s_proj[tid]; }
    if (tid < 4) hdr->uv[tid] = a.uv[(size_t)blk * 4 + tid];
    if (tid == 0) {
      const float a00[3] = {projBox.lo[0], projBox.lo[1], projBox.lo[2]}, a10[3] = {projBox.hi[0], projBox.lo[1], projBox.lo[2]};
      const float a01[3] = {projBox.lo[0], projBox.hi[1], projBox.hi[2]}, a11[3] = {projBox.hi[0], projBox.hi[1], projBox.hi[2]};
      float p00[3], p10[3], p01[3], p11[3];
      project_p(s_iproj, a00, p00); project_p(s_iproj, a10, p10); project_p(s_iproj, a01, p01); project_p(s_iproj, a11, p11);
      hdr->box[0] = projBox.lo[2]; hdr->box[1] = projBox.hi[2];
      hdr->box[2] = p00[0]; hdr->box[3] = p00[1]; hdr->box[4] = p10[0]; hdr->box[5] = p10[1];
      hdr->box[6] = p01[0]; hdr->box[7] = p01[1]; hdr->box[8] = p11[0]; hdr->box[9] = p11[1];
      hdr->geomID = a.geomIDs[blk];
      hdr->primID = a.primIDs[blk];
      if (a.leaf_kind != 1) hdr->extent = 0.f;
#pragma unroll
      for (int k = 0; k < 3; k++) { a.bounds[(size_t)blk * 6 + k] = bounds.lo[k]; a.bounds[(size_t)blk * 6 + 3 + k] = bounds.hi[k]; }
    }

    if (a.leaf_kind == 1) {  // pizza-box leaves (:296-326)
      float ext = 0.f;
      for (unsigned i = tid; i < cells; i += T) {
        const unsigned x = compact(i), y = compact(i >> 1);
        const Box3 pb = heap[elems + i];
        float z[4];
        corner_heights(pb, pv + 3 * (y * W + x), pv + 3 * (y * W + x + 1), pv + 3 * ((y + 1) * W + x), pv + 3 * ((y + 1) * W + x + 1), z);
        float t[4];
#pragma unroll
        for (int q = 0; q < 4; q++) t[q] = fmaxf(fmaxf(z[q] - pb.hi[2], 0.f), fabsf(fminf(z[q] - pb.lo[2], 0.f)));
        const double zF = (double)(pb.hi[2] - pb.lo[2]);
        const float e = zF == 0.0 ? 0.f : (float)((double)fmaxf(fmaxf(t[0], t[1]), fmaxf(t[2], t[3])) / zF);
        ext = fmaxf(ext, e);
      }
      for (int o = 16; o > 0; o >>= 1) ext = fmaxf(ext, __shfl_xor_sync(0xffffffffu, ext, o));
      __syncthreads();
      if ((tid & 31u) == 0) scratch[tid >> 5] = ext;
      __syncthreads();
      if (tid == 0) {
        float e = 0.f;
        for (unsigned w = 0; w < (T >> 5); w++) e = fmaxf(e, scratch[w]);
        e = fminf(e, 1.0f);  // MAX_EXTENT
        s_extent = e;
        hdr->extent = e;
      }
      __syncthreads();
      const float extent = s_extent;
      unsigned short* leaves = reinterpret_cast<unsigned short*>(out + a.leaf_offset);
      for (unsigned i = tid; i < cells; i += T) {
        const unsigned x = compact(i), y = compact(i >> 1);
        const Box3 pb = heap[elems + i];
        float z[4];
        corner_heights(pb, pv + 3 * (y * W + x), pv + 3 * (y * W + x + 1), pv + 3 * ((y + 1) * W + x), pv + 3 * ((y + 1) * W + x + 1), z);
        const double zF = (double)(pb.hi[2] - pb.lo[2]);
        unsigned code[4] = {0, 0, 0, 0};
        if (zF != 0.0) {
          const float zf = (float)((double)(1.0f + 2.0f * extent) * zF);
          const float rcpF = 16.f / zf;
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const float rel = (float)((double)z[q] - ((double)pb.lo[2] - (double)extent * zF));
            code[q] = (unsigned)(unsigned char)fmaxf(0.f, fminf(15.f, rel * rcpF));
          }
        }
        const unsigned z12 = ((code[0] << 4) | (code[1] & 0x0f)) & 0xff, z34 = ((code[2] << 4) | (code[3] & 0x0f)) & 0xff;
        leaves[i] = (unsigned short)(z12 | (z34 << 8));
      }
    } else if (a.leaf_kind == 0) {  // exact leaves: the displaced grid itself (:329-334)
      float* grid = reinterpret_cast<float*>(out + a.leaf_offset);
      for (unsigned i = tid; i < 3 * nv; i += T) grid[i] = v[i];
    }
    __syncthreads();
  }
}

}  // namespace

cudaError_t ecb_launch_build(const EcbBuildLaunch& b, int device, cudaStream_t stream) {
  if (b.n == 0) return cudaSuccess;
  const unsigned W = (1u << b.comp_level) + 1u, cells = 1u << (2 * b.comp_level), elems = (cells - 1) / 3;
  const size_t smem_v = (size_t)3 * W * W * 4 + 6 * (kBuildThreads / 32) * 4;
  const size_t smem_e = (size_t)3 * W * W * 4 + (size_t)(elems + cells) * sizeof(Box3) + 6 * (kBuildThreads / 32) * 4;
  cudaError_t e = cudaFuncSetAttribute(ecb_build_phase_e, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_e);
  if (e != cudaSuccess) return e;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  // a block of 2^C x 2^C cells has (2^C+1)^2 vertices: 81 / 289 / 1089 for C = 3 / 4 / 5
  const unsigned threads = b.comp_level <= 3 ? ECB_BUILD_THREADS_C3 : (unsigned)kBuildThreads;
  const unsigned per_sm = 2048u / threads < 16u ? 2048u / threads : 16u;
  const unsigned grid = b.n < (unsigned)sms * per_sm ? b.n : (unsigned)sms * per_sm;
  const unsigned grid_t = (b.n + 127u) / 128u;
  BuildArgs a;
  a.grids = b.grids; a.space = b.space; a.world = b.world; a.uv = b.uv; a.geomIDs = b.geomIDs; a.primIDs = b.primIDs;
  a.blocks = b.blocks; a.bounds = b.bounds; a.patch_ok = b.patch_ok; a.tmp = static_cast<BuildTmp*>(b.tmp);
  a.n = b.n; a.C = b.comp_level; a.leaf_kind = b.leaf_kind; a.stride = b.stride; a.leaf_offset = b.leaf_offset;
  ecb_build_phase_a<<<grid, threads, smem_v, stream>>>(a);
  if (b.leaf_kind != 0) {  // cGrid never uses the projective mapping: no first estimate to compute or test
    ecb_build_phase_b<<<grid_t, 128, 0, stream>>>(a);
    ecb_build_phase_c<<<grid, threads, smem_v, stream>>>(a);
  }
  ecb_build_phase_d<<<grid_t, 128, 0, stream>>>(a);
  ecb_build_phase_e<<<grid, threads, smem_e, stream>>>(a);
  return cudaGetLastError();
}

size_t ecb_build_tmp_bytes(uint32_t n) { return (size_t)n * sizeof(BuildTmp); }
