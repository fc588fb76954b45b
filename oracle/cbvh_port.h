/* oracle/cbvh_port.h -- CPU restatement ("port") of the reference's compressed-BVH traversal.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product (embree_compressed_b200/) never does.
 *
 * The port is plain scalar C over the REFERENCE's own committed memory layout (128-byte BVH4
 * AlignedNodes with tagged 64-bit child refs, 184-byte CompressedBVH block headers followed by
 * 4-byte quantized nodes and 2-byte leaf heights / 12-byte grid vertices).  Every address is
 * `base + (ref & ~15)`: base = 0 walks a live reference scene in this process, base = buffer
 * walks a relocated image (oracle "RIMG", see rimg.h).
 *
 * Pinned against: the compiled reference itself (oracle/_ref, rtcIntersect1 / rtcOccluded1 on the
 * same rays) and the golden fixtures under tests/golden/ that were generated from it.  The
 * reference's own tests hold no vectors for this path.
 *
 * Arithmetic: fp32, no contraction (-ffp-contract=off), IEEE division / sqrt where the
 * reference uses rcpss/rsqrtss + one Newton step (common/math/math.h:60-97), fmaf() only where
 * the reference's AVX2 build has an explicit FMA (xfmPoint/xfmVector: common/math/linearspace3.h:168-169).
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { CBVH_LEAF_GRID = 0, CBVH_LEAF_LEAF = 1, CBVH_LEAF_BOX = 2 };

typedef struct cbvh_scene {
  const uint8_t* base;      /* 0 for live reference memory, image payload otherwise */
  uint64_t root;            /* BVH4 NodeRef (kernels/bvh/bvh.h:150-400) */
  int32_t leaf_kind;        /* CBVH_LEAF_* : which CompressedBVHIntersector1 variant (compressed.h:777-779) */
  int32_t proj_col_major;   /* storage order of Eigen::Matrix3f proj in the block header */
  uint64_t tri_root;        /* 0, or the root NodeRef of the scene's BVH4 over Triangle4v blocks (mixed scenes) */
} cbvh_scene;

/* include/embree3/rtcore_ray.h:26-64 */
typedef struct cbvh_ray {
  float org_x, org_y, org_z, tnear, dir_x, dir_y, dir_z, time, tfar;
  uint32_t mask, id, flags;
} cbvh_ray;
typedef struct cbvh_hit {
  float Ng_x, Ng_y, Ng_z, u, v;
  uint32_t primID, geomID, instID;
} cbvh_hit;
typedef struct cbvh_rayhit {
  cbvh_ray ray;
  cbvh_hit hit;
} cbvh_rayhit;

/* per-ray work counters (N_top, N_blk, N_cnode, N_cell) */
typedef struct cbvh_stats {
  uint64_t rays, top_nodes, blocks, cnodes, cells, hits;
} cbvh_stats;

/* optional second-nearest-candidate record used by the tests to compute the near-tie set */
typedef struct cbvh_candidates {
  float t_best, t_second;        /* nearest and second nearest accepted leaf candidates (world t), +inf if none */
  uint32_t prim_best, prim_second;
} cbvh_candidates;

void cbvh_intersect1(const cbvh_scene* scene, cbvh_rayhit* rh, uint32_t instID, cbvh_stats* stats);
void cbvh_occluded1(const cbvh_scene* scene, cbvh_ray* ray, cbvh_stats* stats);

/* all rays of a strided array (inactive rays tnear>tfar skipped), nthreads pthreads; returns wall seconds */
double cbvh_intersectN(const cbvh_scene* scene, void* rayhits, size_t n, size_t stride, int nthreads, cbvh_stats* stats);
double cbvh_occludedN(const cbvh_scene* scene, void* rays, size_t n, size_t stride, int nthreads, cbvh_stats* stats);

/* Like cbvh_intersect1 but never shrinks the search interval: records the two nearest leaf candidates. */
void cbvh_candidates1(const cbvh_scene* scene, const cbvh_rayhit* rh, cbvh_candidates* out);
void cbvh_candidatesN(const cbvh_scene* scene, const void* rayhits, size_t n, size_t stride, int nthreads, cbvh_candidates* out);

/* ---- instanced scenes --------------------------------------------------------------------------
 * Top level: the reference's user-geometry BVH4 (NON-robust node test, leaves = arrays of Object{geomID,primID});
 * each instance transforms the ray with world2local, sets instID = its geomID and runs cbvh_intersect1 /
 * cbvh_occluded1 on the instanced scene (kernels/geometry/instance_intersector.cpp:45-83). */
typedef struct cbvh_instance {
  float world2local[12];   /* vx, vy, vz, p */
  uint32_t geomID;
  uint32_t object;         /* index into cbvh_inst_scene::objects */
} cbvh_instance;

typedef struct cbvh_inst_scene {
  const uint8_t* base;
  uint64_t root;                 /* root NodeRef of the user-geometry BVH4 */
  const cbvh_instance* instances;
  uint32_t n_instances;
  uint32_t reserved;
  const uint32_t* geom_to_instance; /* geomID -> index into instances (0xFFFFFFFF if not an instance) */
  uint32_t n_geoms;
  uint32_t reserved2;
  const cbvh_scene* objects;
} cbvh_inst_scene;

void cbvh_inst_intersect1(const cbvh_inst_scene* scene, cbvh_rayhit* rh, cbvh_stats* stats);
void cbvh_inst_occluded1(const cbvh_inst_scene* scene, cbvh_ray* ray, cbvh_stats* stats);
double cbvh_inst_intersectN(const cbvh_inst_scene* scene, void* rayhits, size_t n, size_t stride, int nthreads, cbvh_stats* stats);
void cbvh_inst_candidates1(const cbvh_inst_scene* scene, const cbvh_rayhit* rh, cbvh_candidates* out);
void cbvh_inst_candidatesN(const cbvh_inst_scene* scene, const void* rayhits, size_t n, size_t stride, int nthreads, cbvh_candidates* out);
double cbvh_inst_occludedN(const cbvh_inst_scene* scene, void* rays, size_t n, size_t stride, int nthreads, cbvh_stats* stats);

/* ---- block construction, oracle/cbvh_build_port.c ---------------------------------------------------
 * One compressed block from its displaced vertex grid and its frame; restates CompressedBVH::CompressedBVH
 * (kernels/geometry/compressed.h:128-337).  comp_level <= 5. */
typedef struct cbvh_build_in {
  const float* grid;     /* (2^C+1)^2 x 3 floats, row-major y*width+x (compressed.h:69-77) */
  float space[9];        /* LinearSpace3f world->local: vx, vy, vz columns (compressed.h:125-126) */
  float world[9];        /* LinearSpace3f(vx,vy,vz) it was inverted from (only the world bounds need it) */
  int32_t leaf_kind;     /* CBVH_LEAF_* */
  int32_t comp_level;    /* C */
} cbvh_build_in;

typedef struct cbvh_build_out {
  float proj[9];         /* homography, row-major */
  float box[10];         /* frustum box */
  float extent;          /* cLeaf only */
  float bounds[6];       /* conservative world bounds: lower xyz, upper xyz */
  int32_t patch_ok;      /* the projective mapping was applicable */
  uint8_t nodes[4 * 341];   /* xz, x, yz, y per inner node (low two bits of x and y are zero here, unspecified in the reference) */
  uint8_t leaves[2 * 1024]; /* z12, z34 per cell (cLeaf only) */
} cbvh_build_out;

void cbvh_build_block(const cbvh_build_in* in, cbvh_build_out* out);
void cbvh_build_blocks(const cbvh_build_in* in, cbvh_build_out* out, size_t n, int nthreads);

#ifdef __cplusplus
}
#endif
