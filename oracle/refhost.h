/* oracle/refhost.h -- C interface of the reference harness (TEST INFRASTRUCTURE ONLY).
 *
 * oracle/refhost.cpp is compiled by oracle/build_ref.py against the UNMODIFIED reference
 * sources into oracle/_ref/libembree3_ref.so.  It drives the reference through its public rtc*
 * API (scene construction, rtcIntersect1 / rtcOccluded1 loops = the parity oracle and the CPU
 * baseline) and exposes the one thing a host application would hand to the B200 plugin at
 * commit time: the root NodeRef of the committed compressed BVH4 (INTEGRATION.md shows the
 * equivalent 10-line accessor a maintainer adds to the reference).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct RefScene RefScene;

/* sizeof/offsetof of the reference structs the product's importer hard-codes (probe for tests) */
typedef struct RefLayout {
  uint32_t sizeof_block;      /* sizeof(CompressedBVH<compressedNonUniform>)  kernels/geometry/compressed.h:408-433 */
  uint32_t off_geomID, off_primID, off_uv, off_rcp_edges, off_nodes, off_leaves, off_grid, off_extent;
  uint32_t off_elems, off_grid_width, off_space, off_proj, off_box;
  uint32_t proj_col_major;    /* 1 if the compiled Eigen::Matrix3f stores column-major (real Eigen default) */
  uint32_t sizeof_node, sizeof_leaf, sizeof_vec3f, sizeof_aligned_node, sizeof_rayhit;
  uint32_t off_an_children, off_an_lower_x;
} RefLayout;
void ref_layout(RefLayout* out);

/* leaf: "grid" | "leaf" | "box" (device config subdiv_accel=bvh4.compressed.<leaf>), or "eager" for stock bvh4.grid.eager */
RefScene* ref_scene_new(const char* leaf, unsigned subdivLevel, unsigned compLevel, int build_threads);
void ref_scene_free(RefScene*);

/* Adds one Catmull-Clark subdivision mesh.  displacement: 0 none, 1 closed-form "ridged sines" (see refhost.cpp). */
unsigned ref_scene_add_subdiv(RefScene*, const float* xyz, unsigned nverts, const unsigned* indices, unsigned nindices,
                              const unsigned* faces, unsigned nfaces, int displacement, float disp_scale);
/* the displaced cube of tutorials/displacement_geometry (8 verts / 6 quads) */
unsigned ref_scene_add_cube(RefScene*, int displacement, float disp_scale);
/* Wavefront OBJ (v / f only, all objects merged into ONE subdiv mesh like the reference viewer does) */
int ref_scene_add_obj(RefScene*, const char* path);
/* two-triangle ground plane of tutorials/displacement_geometry (y=-2, +-10) */
unsigned ref_scene_add_ground_plane(RefScene*);
/* instance of another committed scene with a 3x4 column-major affine transform (vx,vy,vz,p) */
unsigned ref_scene_add_instance(RefScene*, RefScene* object, const float* xfm12);

void ref_scene_commit(RefScene*);
int ref_scene_error(RefScene*);                 /* sticky RTCError of the device */
void ref_scene_bounds(RefScene*, float* lower_upper6);
void* ref_scene_handle(RefScene*);              /* the RTCScene */

/* Root NodeRef (tagged pointer, kernels/bvh/bvh.h:150-400) of the BVH4 whose leaves are CompressedBVH blocks.
 * Returns 0 when the scene has no such accel.  *n_accels = number of accels in the scene's AccelN (1 = direct). */
uint64_t ref_scene_compressed_root(RefScene*, int* n_accels);

/* rtcIntersect1 / rtcOccluded1 applied to every ray (the only entry points that work on the compressed accel,
 * SURVEY 3.3), spread over nthreads std::threads in 4096-ray chunks.  Returns wall seconds. */
double ref_intersect(RefScene*, void* rayhits, size_t n, size_t stride, int nthreads);
double ref_occluded(RefScene*, void* rays, size_t n, size_t stride, int nthreads);

/* For every hit in rayhits: Ns = cross(dPdv,dPdu) via rtcInterpolate1 (tutorials/viewer/viewer_device.cpp:284-292);
 * writes 3 floats per ray (zeros for misses). */
void ref_hit_normals(RefScene*, const void* rayhits, size_t n, size_t stride, float* normals3);

/* Relocatable image ("RIMG") of the committed compressed accel: every AlignedNode and CompressedBVH block copied
 * verbatim (reference layout) into one buffer, pointers replaced by byte offsets from the start of the buffer with
 * their tag bits kept.  Layout of the 96-byte header: see rimg.h.  Returns the image size (call with cap=0 to size). */
size_t ref_scene_dump_image(RefScene*, void* buf, size_t cap);

/* Instanced scenes: root NodeRef of the user-geometry BVH4 and, for each instance geometry (up to cap), its geomID,
 * world2local (vx,vy,vz,p = 12 floats) and the RTCScene it instances.  Returns the number of instances, <0 if none. */
int ref_scene_instance_info(RefScene*, uint64_t* root_out, unsigned cap, unsigned* geomIDs, float* world2local12, void** object_scenes);

/* Mixed scenes: root of the BVH4 over Triangle4v blocks (0 if none); positions of the triangle and the compressed accel
 * in the scene's AccelN list (the order rtcIntersect1 visits them, kernels/common/acceln.cpp:51-56). */
uint64_t ref_scene_triangle_root(RefScene*, int* tri_index, int* subdiv_index, unsigned* sizeof_block);

int ref_hardware_threads(void);

#ifdef __cplusplus
}
#endif
