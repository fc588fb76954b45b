// oracle/refhost.cpp -- harness around the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY; see refhost.h).
//
// Compiled by oracle/build_ref.py together with the reference's own translation units.  It includes the
// reference's internal headers for exactly two purposes:
//   (1) ref_scene_compressed_root(): read `bvh->root` of the committed compressed accel, the way the
//       reference's own tutorials/bvh_access/bvh_access.cpp:199-219 walks a committed scene;
//   (2) ref_layout(): report sizeof/offsetof of the structs the product's importer hard-codes, so a test can
//       prove the hard-coded layout (include/ecb200.h) matches the compiled reference.
// Everything else goes through the public rtc* API.
#include "refhost.h"
#include "rimg.h"

#include <embree3/rtcore.h>

#include "kernels/common/default.h"
#include "kernels/common/scene.h"
#include "kernels/common/accel.h"
#include "kernels/common/acceln.h"
#include "kernels/bvh/bvh.h"
#include "kernels/geometry/subdivpatch1cached.h"
#include "kernels/geometry/grid_soa.h"
#include "kernels/geometry/compressed.h"
#include "kernels/geometry/object.h"
#include "kernels/geometry/trianglev.h"
#include "kernels/common/scene_instance.h"

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {

struct Mesh {
  std::vector<float> xyzw;          // stride 16 bytes (reference tutorials share Vec3fa buffers)
  std::vector<unsigned> indices;
  std::vector<unsigned> faces;
  std::vector<float> levels;
  float disp_scale = 0.f;
};

std::mutex g_mutex;
std::map<std::string, RTCDevice> g_devices;

RTCDevice get_device(const std::string& leaf, int threads) {
  std::lock_guard<std::mutex> lock(g_mutex);
  std::string accel = leaf.compare(0, 5, "eager") == 0 ? "bvh4.grid." + leaf : "bvh4.compressed." + leaf;
  // set_affinity=0: with the internal tasking system the default pins the calling thread (kernels/common/state.cpp:123-128)
  // object_accel=bvh4.object: instances go into a BVH4 (the default on AVX hardware is a BVH8, kernels/common/scene.cpp:531-555)
  // tri_accel=bvh4.triangle4v: triangle meshes of a ROBUST scene go into a BVH4 of Triangle4v blocks with the Pluecker
  // intersector (the default on AVX hardware is the BVH8 flavour of the same thing, kernels/common/scene.cpp:158-164)
  std::string cfg = "subdiv_accel=" + accel + ",object_accel=bvh4.object,tri_accel=bvh4.triangle4v,set_affinity=0,start_threads=0";
  if (threads > 0) cfg += ",threads=" + std::to_string(threads);
  auto it = g_devices.find(cfg);
  if (it != g_devices.end()) return it->second;
  RTCDevice d = rtcNewDevice(cfg.c_str());
  g_devices[cfg] = d;
  return d;
}

// Closed-form stand-in for the tutorial's Perlin-noise displacement (displacement_geometry_device.cpp:89-97):
// same octave structure  sum_{freq=1,2,..,32} 1.4 n^2 / freq  with n a smooth product of sines instead of noise().
inline float displacement_value(float x, float y, float z) {
  float dN = 0.0f;
  for (float freq = 1.0f; freq < 40.0f; freq *= 2.0f) {
    const float n = std::sin(freq * (1.7f * x + 0.3f)) * std::sin(freq * (1.3f * y + 1.1f)) * std::sin(freq * (1.9f * z + 2.3f));
    dN += 1.4f * n * n / freq;
  }
  return dN;
}

void displacement_cb(const struct RTCDisplacementFunctionNArguments* args) {
  const Mesh* mesh = static_cast<const Mesh*>(args->geometryUserPtr);
  const float s = mesh ? mesh->disp_scale : 1.0f;
  for (unsigned i = 0; i < args->N; i++) {
    const float d = s * displacement_value(args->P_x[i], args->P_y[i], args->P_z[i]);
    args->P_x[i] += d * args->Ng_x[i];
    args->P_y[i] += d * args->Ng_y[i];
    args->P_z[i] += d * args->Ng_z[i];
  }
}

}  // namespace

struct RefScene {
  RTCDevice device = nullptr;
  RTCScene scene = nullptr;
  std::vector<std::unique_ptr<Mesh>> meshes;  // shared buffers must outlive the scene
  int leaf_kind = -1;                          // 0 grid, 1 leaf, 2 box (-1: stock accel)
  unsigned S = 0, C = 0;
};

REF_API void ref_layout(RefLayout* o) {
  using namespace embree;
  typedef isa::compressed::CompressedBVH<isa::compressed::compressedNonUniform> Block;
  memset(o, 0, sizeof(*o));
  o->sizeof_block = sizeof(Block);
#define OFF(m) (uint32_t)((size_t) & reinterpret_cast<char const volatile&>(((Block*)0x1000)->m) - 0x1000)
  o->off_geomID = OFF(geomID);
  o->off_primID = OFF(primID);
  o->off_uv = OFF(uv);
  o->off_rcp_edges = OFF(rcp_edges);
  o->off_nodes = OFF(nodes);
  o->off_leaves = OFF(leaves);
  o->off_grid = OFF(grid);
  o->off_extent = OFF(extent);
  o->off_elems = OFF(elems);
  o->off_grid_width = OFF(grid_width);
  o->off_space = OFF(space);
  o->off_proj = OFF(proj);
  o->off_box = OFF(box);
#undef OFF
  Eigen::Matrix3f m;
  m << 0.f, 1.f, 2.f, 3.f, 4.f, 5.f, 6.f, 7.f, 8.f;  // (r,c) = 3r+c
  const float* raw = reinterpret_cast<const float*>(&m);
  o->proj_col_major = (raw[1] == 3.f) ? 1u : 0u;
  o->sizeof_node = sizeof(isa::compressed::compressedNonUniform);
  o->sizeof_leaf = sizeof(isa::compressed::leaf_t);
  o->sizeof_vec3f = sizeof(Vec3f);
  o->sizeof_aligned_node = sizeof(BVH4::AlignedNode);
  o->sizeof_rayhit = sizeof(RTCRayHit);
  o->off_an_children = (uint32_t)((size_t) & reinterpret_cast<char const volatile&>(((BVH4::AlignedNode*)0x1000)->children) - 0x1000);
  o->off_an_lower_x = (uint32_t)((size_t) & reinterpret_cast<char const volatile&>(((BVH4::AlignedNode*)0x1000)->lower_x) - 0x1000);
}

REF_API RefScene* ref_scene_new(const char* leaf, unsigned S, unsigned C, int build_threads) {
  RefScene* s = new RefScene;
  s->device = get_device(leaf, build_threads);
  s->scene = rtcNewScene(s->device);
  // tutorials/displacement_geometry/displacement_geometry_device.cpp:186 ; the compressed intersectors are
  // instantiated robust-only anyway (kernels/bvh/bvh_intersector1_bvh4.cpp:63-78)
  rtcSetSceneFlags(s->scene, RTC_SCENE_FLAG_ROBUST);
  rtcSetSceneLevels(s->scene, S, C);
  // `leaf` may carry extra device options after a comma, e.g. "leaf,max_isa=sse4.2" (used to measure how far the
  // reference's own ISA code paths drift from each other on the same rays)
  const std::string kind = std::string(leaf).substr(0, std::string(leaf).find(','));
  s->leaf_kind = kind == "grid" ? 0 : kind == "leaf" ? 1 : kind == "box" ? 2 : -1;
  s->S = S;
  s->C = C;
  return s;
}

REF_API void ref_scene_free(RefScene* s) {
  if (!s) return;
  rtcReleaseScene(s->scene);
  delete s;
}

REF_API unsigned ref_scene_add_subdiv(RefScene* s, const float* xyz, unsigned nverts, const unsigned* indices,
                                      unsigned nindices, const unsigned* faces, unsigned nfaces, int displacement,
                                      float disp_scale) {
  std::unique_ptr<Mesh> m(new Mesh);
  m->xyzw.resize(4 * (size_t)nverts + 4, 0.f);
  for (unsigned i = 0; i < nverts; i++)
    for (int k = 0; k < 3; k++) m->xyzw[4 * (size_t)i + k] = xyz[3 * (size_t)i + k];
  m->indices.assign(indices, indices + nindices);
  m->faces.assign(faces, faces + nfaces);
  m->levels.assign(nindices, 1.0f);  // ignored: the compressed builder forces 2^S (bvh_builder_subdiv.cpp:771-775)
  m->disp_scale = disp_scale;
  RTCGeometry geom = rtcNewGeometry(s->device, RTC_GEOMETRY_TYPE_SUBDIVISION);
  rtcSetSharedGeometryBuffer(geom, RTC_BUFFER_TYPE_VERTEX, 0, RTC_FORMAT_FLOAT3, m->xyzw.data(), 0, 16, nverts);
  rtcSetSharedGeometryBuffer(geom, RTC_BUFFER_TYPE_INDEX, 0, RTC_FORMAT_UINT, m->indices.data(), 0, 4, nindices);
  rtcSetSharedGeometryBuffer(geom, RTC_BUFFER_TYPE_FACE, 0, RTC_FORMAT_UINT, m->faces.data(), 0, 4, nfaces);
  rtcSetSharedGeometryBuffer(geom, RTC_BUFFER_TYPE_LEVEL, 0, RTC_FORMAT_FLOAT, m->levels.data(), 0, 4, nindices);
  if (displacement) {
    rtcSetGeometryUserData(geom, m.get());
    rtcSetGeometryDisplacementFunction(geom, displacement_cb);
  }
  rtcCommitGeometry(geom);
  unsigned id = rtcAttachGeometry(s->scene, geom);
  rtcReleaseGeometry(geom);
  s->meshes.push_back(std::move(m));
  return id;
}

REF_API unsigned ref_scene_add_cube(RefScene* s, int displacement, float disp_scale) {
  // tutorials/displacement_geometry/displacement_geometry_device.cpp:37-66
  static const float v[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, -1, 1}, {-1, -1, 1}, {-1, 1, -1}, {1, 1, -1}, {1, 1, 1}, {-1, 1, 1}};
  static const unsigned idx[24] = {0, 4, 5, 1, 1, 5, 6, 2, 2, 6, 7, 3, 0, 3, 7, 4, 4, 7, 6, 5, 0, 1, 2, 3};
  static const unsigned faces[6] = {4, 4, 4, 4, 4, 4};
  return ref_scene_add_subdiv(s, &v[0][0], 8, idx, 24, faces, 6, displacement, disp_scale);
}

REF_API int ref_scene_add_obj(RefScene* s, const char* path) {
  FILE* f = fopen(path, "r");
  if (!f) return -1;
  std::vector<float> xyz;
  std::vector<unsigned> indices, faces;
  char line[4096];
  while (fgets(line, sizeof(line), f)) {
    if (line[0] == 'v' && (line[1] == ' ' || line[1] == '\t')) {
      float x, y, z;
      if (sscanf(line + 2, "%f %f %f", &x, &y, &z) == 3) { xyz.push_back(x); xyz.push_back(y); xyz.push_back(z); }
    } else if (line[0] == 'f' && (line[1] == ' ' || line[1] == '\t')) {
      unsigned n = 0;
      char* p = line + 2;
      while (*p) {
        while (*p == ' ' || *p == '\t') p++;
        if (*p == 0 || *p == '\n' || *p == '\r') break;
        long i = strtol(p, &p, 10);
        if (i < 0) i = (long)(xyz.size() / 3) + i + 1;
        indices.push_back((unsigned)(i - 1));
        n++;
        while (*p && *p != ' ' && *p != '\t' && *p != '\n' && *p != '\r') p++;  // skip /vt/vn
      }
      if (n >= 3) faces.push_back(n); else indices.resize(indices.size() - n);
    }
  }
  fclose(f);
  if (faces.empty()) return -2;
  return (int)ref_scene_add_subdiv(s, xyz.data(), (unsigned)(xyz.size() / 3), indices.data(), (unsigned)indices.size(),
                                   faces.data(), (unsigned)faces.size(), 0, 0.f);
}

REF_API unsigned ref_scene_add_ground_plane(RefScene* s) {
  // tutorials/displacement_geometry/displacement_geometry_device.cpp:156-177
  RTCGeometry geom = rtcNewGeometry(s->device, RTC_GEOMETRY_TYPE_TRIANGLE);
  float* v = (float*)rtcSetNewGeometryBuffer(geom, RTC_BUFFER_TYPE_VERTEX, 0, RTC_FORMAT_FLOAT3, 16, 4);
  const float vv[4][3] = {{-10, -2, -10}, {-10, -2, 10}, {10, -2, -10}, {10, -2, 10}};
  for (int i = 0; i < 4; i++) { v[4 * i] = vv[i][0]; v[4 * i + 1] = vv[i][1]; v[4 * i + 2] = vv[i][2]; v[4 * i + 3] = 0.f; }
  unsigned* t = (unsigned*)rtcSetNewGeometryBuffer(geom, RTC_BUFFER_TYPE_INDEX, 0, RTC_FORMAT_UINT3, 12, 2);
  const unsigned tt[6] = {0, 1, 2, 1, 3, 2};
  memcpy(t, tt, sizeof(tt));
  rtcCommitGeometry(geom);
  unsigned id = rtcAttachGeometry(s->scene, geom);
  rtcReleaseGeometry(geom);
  return id;
}

REF_API unsigned ref_scene_add_instance(RefScene* s, RefScene* object, const float* xfm12) {
  RTCGeometry geom = rtcNewGeometry(s->device, RTC_GEOMETRY_TYPE_INSTANCE);
  rtcSetGeometryInstancedScene(geom, object->scene);
  rtcSetGeometryTransform(geom, 0, RTC_FORMAT_FLOAT3X4_COLUMN_MAJOR, xfm12);
  rtcCommitGeometry(geom);
  unsigned id = rtcAttachGeometry(s->scene, geom);
  rtcReleaseGeometry(geom);
  return id;
}

REF_API void ref_scene_commit(RefScene* s) { rtcCommitScene(s->scene); }
REF_API int ref_scene_error(RefScene* s) { return (int)rtcGetDeviceError(s->device); }
REF_API void* ref_scene_handle(RefScene* s) { return s->scene; }

REF_API void ref_scene_bounds(RefScene* s, float* b) {
  RTCBounds bb;
  rtcGetSceneBounds(s->scene, &bb);
  b[0] = bb.lower_x; b[1] = bb.lower_y; b[2] = bb.lower_z;
  b[3] = bb.upper_x; b[4] = bb.upper_y; b[5] = bb.upper_z;
}

REF_API uint64_t ref_scene_triangle_root(RefScene* s, int* tri_index, int* subdiv_index, unsigned* sizeof_block);

REF_API uint64_t ref_scene_compressed_root(RefScene* s, int* n_accels) {
  using namespace embree;
  // same access pattern as the reference's tutorials/bvh_access/bvh_access.cpp:199-219
  AccelData* accel = ((Accel*)s->scene)->intersectors.ptr;
  if (n_accels) *n_accels = 0;
  if (!accel) return 0;
  if (accel->type == AccelData::TY_BVH4) {
    if (n_accels) *n_accels = 1;
    BVH4* bvh = (BVH4*)accel;
    if (bvh->primTy != &SubdivPatch1Cached::type) return 0;  // factories tag compressed accels so (bvh4_factory.cpp:1653-1675)
    return (uint64_t)(size_t)bvh->root;
  }
  if (accel->type == AccelData::TY_ACCELN) {
    AccelN* an = (AccelN*)accel;
    if (n_accels) *n_accels = (int)an->validAccels.size();  // non-empty accels: what AccelN::intersect loops over
    for (size_t i = 0; i < an->accels.size(); i++) {
      AccelData* a = an->accels[i]->intersectors.ptr;
      if (a && a->type == AccelData::TY_BVH4) {
        BVH4* bvh = (BVH4*)a;
        if (bvh->primTy == &SubdivPatch1Cached::type && bvh->numPrimitives) return (uint64_t)(size_t)bvh->root;
      }
    }
  }
  return 0;
}

namespace {
template <typename F>
double run_chunks(size_t n, int nthreads, F&& body) {
  if (nthreads < 1) nthreads = 1;
  const size_t chunk = 4096;
  std::atomic<size_t> next(0);
  auto worker = [&]() {
    for (;;) {
      const size_t b = next.fetch_add(chunk);
      if (b >= n) break;
      const size_t e = b + chunk < n ? b + chunk : n;
      body(b, e);
    }
  };
  const auto t0 = std::chrono::steady_clock::now();
  if (nthreads == 1) {
    worker();
  } else {
    std::vector<std::thread> th;
    for (int i = 0; i < nthreads; i++) th.emplace_back(worker);
    for (auto& t : th) t.join();
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
}  // namespace

REF_API double ref_intersect(RefScene* s, void* rayhits, size_t n, size_t stride, int nthreads) {
  RTCScene scene = s->scene;
  char* base = static_cast<char*>(rayhits);
  return run_chunks(n, nthreads, [=](size_t b, size_t e) {
    for (size_t i = b; i < e; i++) {
      RTCRayHit* rh = reinterpret_cast<RTCRayHit*>(base + i * stride);
      if (rh->ray.tnear > rh->ray.tfar) continue;  // inactive stream ray (man/man3/rtcIntersect1M.3embree3)
      RTCIntersectContext ctx;
      rtcInitIntersectContext(&ctx);
      rtcIntersect1(scene, &ctx, rh);
    }
  });
}

REF_API double ref_occluded(RefScene* s, void* rays, size_t n, size_t stride, int nthreads) {
  RTCScene scene = s->scene;
  char* base = static_cast<char*>(rays);
  return run_chunks(n, nthreads, [=](size_t b, size_t e) {
    for (size_t i = b; i < e; i++) {
      RTCRay* r = reinterpret_cast<RTCRay*>(base + i * stride);
      if (r->tnear > r->tfar) continue;
      RTCIntersectContext ctx;
      rtcInitIntersectContext(&ctx);
      rtcOccluded1(scene, &ctx, r);
    }
  });
}

REF_API void ref_hit_normals(RefScene* s, const void* rayhits, size_t n, size_t stride, float* normals3) {
  RTCScene scene = s->scene;
  const char* base = static_cast<const char*>(rayhits);
  run_chunks(n, ref_hardware_threads(), [=](size_t b, size_t e) {
    for (size_t i = b; i < e; i++) {
      const RTCRayHit* rh = reinterpret_cast<const RTCRayHit*>(base + i * stride);
      float* N = normals3 + 3 * i;
      N[0] = N[1] = N[2] = 0.f;
      if (rh->hit.geomID == RTC_INVALID_GEOMETRY_ID || rh->hit.instID[0] != RTC_INVALID_GEOMETRY_ID) continue;
      __attribute__((aligned(16))) float dPdu[4], dPdv[4];
      rtcInterpolate1(rtcGetGeometry(scene, rh->hit.geomID), rh->hit.primID, rh->hit.u, rh->hit.v, RTC_BUFFER_TYPE_VERTEX, 0,
                      nullptr, dPdu, dPdv, 3);
      N[0] = dPdv[1] * dPdu[2] - dPdv[2] * dPdu[1];
      N[1] = dPdv[2] * dPdu[0] - dPdv[0] * dPdu[2];
      N[2] = dPdv[0] * dPdu[1] - dPdv[1] * dPdu[0];
    }
  });
}

REF_API int ref_hardware_threads(void) {
  unsigned n = std::thread::hardware_concurrency();
  return n ? (int)n : 1;
}

// ---- relocatable image of the committed compressed accel (rimg.h) ----
namespace {
struct Dumper {
  RefScene* s;
  char* buf;
  size_t cap, pos;
  uint32_t n_nodes = 0, n_blocks = 0;
  bool tri = false;  // walking the Triangle4v BVH4: a leaf is an array of `num` 176-byte blocks
  std::map<uint64_t, uint64_t> seen;

  size_t alloc(size_t bytes) {
    const size_t at = (pos + 15) & ~size_t(15);
    pos = at + bytes;
    return at;
  }
  void put(size_t at, const void* src, size_t bytes) {
    if (buf && at + bytes <= cap) memcpy(buf + at, src, bytes);
  }
  uint64_t visit(uint64_t ref) {
    using namespace embree;
    if (ref == BVH4::emptyNode) return ref;
    const uint64_t tag = ref & 15, ptr = ref & ~uint64_t(15);
    auto it = seen.find(ptr);
    if (it != seen.end()) return it->second | tag;
    if ((ref & BVH4::tyLeaf) && tri) {
      const size_t num = (size_t)tag - BVH4::tyLeaf;
      const size_t at = alloc(num * sizeof(Triangle4v));
      seen[ptr] = at;
      put(at, (const void*)ptr, num * sizeof(Triangle4v));
      return at | tag;
    }
    if (ref & BVH4::tyLeaf) {
      typedef isa::compressed::CompressedBVH<isa::compressed::compressedNonUniform> Block;
      const Block* b = (const Block*)ptr;
      const size_t cells = 3 * (size_t)b->elems + 1;
      const size_t nodes_b = 4 * (size_t)b->elems;
      const size_t leaves_b = s->leaf_kind == 1 ? 2 * cells : 0;
      const size_t grid_b = s->leaf_kind == 0 ? 12 * (size_t)b->grid_width * b->grid_width : 0;
      const size_t at = alloc(sizeof(Block) + nodes_b + leaves_b + grid_b);
      seen[ptr] = at;
      n_blocks++;
      std::vector<char> tmp(sizeof(Block) + nodes_b + leaves_b + grid_b);
      memcpy(tmp.data(), b, sizeof(Block));
      memcpy(tmp.data() + sizeof(Block), b->nodes, nodes_b);
      if (leaves_b) memcpy(tmp.data() + sizeof(Block) + nodes_b, b->leaves, leaves_b);
      if (grid_b) memcpy(tmp.data() + sizeof(Block) + nodes_b + leaves_b, b->grid, grid_b);
      Block* nb = (Block*)tmp.data();
      nb->nodes = (decltype(nb->nodes))(at + sizeof(Block));
      nb->leaves = (decltype(nb->leaves))(at + sizeof(Block) + nodes_b);
      nb->grid = (decltype(nb->grid))(at + sizeof(Block) + nodes_b + leaves_b);
      put(at, tmp.data(), tmp.size());
      return at | tag;
    }
    const BVH4::AlignedNode* n = (const BVH4::AlignedNode*)ptr;
    const size_t at = alloc(sizeof(BVH4::AlignedNode));
    seen[ptr] = at;
    n_nodes++;
    char tmp[sizeof(BVH4::AlignedNode)];
    memcpy(tmp, n, sizeof(tmp));
    for (int i = 0; i < 4; i++) {
      const uint64_t c = visit((uint64_t)(size_t)n->child(i));
      memcpy(tmp + 8 * i, &c, 8);
    }
    put(at, tmp, sizeof(tmp));
    return at | tag;
  }
};
}  // namespace

REF_API size_t ref_scene_dump_image(RefScene* s, void* buf, size_t cap) {
  int n_accels = 0;
  const uint64_t root = ref_scene_compressed_root(s, &n_accels);
  if (!root || s->leaf_kind < 0) return 0;
  const uint64_t tri_root = ref_scene_triangle_root(s, nullptr, nullptr, nullptr);
  Dumper d;
  d.s = s;
  d.buf = (char*)buf;
  d.cap = cap;
  d.pos = sizeof(rimg_header);
  rimg_header h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, RIMG_MAGIC, 8);
  h.version = 1;
  h.leaf_kind = (uint32_t)s->leaf_kind;
  h.subdiv_level = s->S;
  h.comp_level = s->C;
  RefLayout lay;
  ref_layout(&lay);
  h.proj_col_major = lay.proj_col_major;
  h.root = d.visit(root);
  h.n_nodes = d.n_nodes;
  h.n_blocks = d.n_blocks;
  if (tri_root) {  // mixed scene: the triangle accel travels in the same image
    d.tri = true;
    h.tri_root = d.visit(tri_root);
  }
  h.total_bytes = (d.pos + 15) & ~size_t(15);
  ref_scene_bounds(s, h.bounds);
  d.put(0, &h, sizeof(h));
  return (size_t)h.total_bytes;
}

// ---- the accessor a maintainer adds to the reference, supplied here for the test build ----
// Everything the plugin needs at commit time, from the reference's own objects: root of the compressed accel
// (same walk as ref_scene_compressed_root), leaf flavour from Device::subdiv_accel (kernels/common/state.h:109),
// S and C from Scene::subdivisionLevel / compressionLevel (kernels/common/scene.h:231-232).
struct RTCCompressedBVHInfo_b200 { uint64_t root; uint32_t leaf_kind, subdiv_level, comp_level, proj_col_major, n_accels; float bounds[6]; };
REF_API int rtcGetCompressedBVH_b200(RTCScene hscene, RTCCompressedBVHInfo_b200* out) {
  using namespace embree;
  Scene* scene = (Scene*)hscene;
  memset(out, 0, sizeof(*out));
  const std::string& a = scene->device->subdiv_accel;
  if (a == "bvh4.compressed.grid") out->leaf_kind = 0;
  else if (a == "bvh4.compressed.leaf") out->leaf_kind = 1;
  else if (a == "bvh4.compressed.box") out->leaf_kind = 2;
  else return -1;
  RefScene tmp;
  tmp.scene = hscene;
  int n = 0;
  out->root = ref_scene_compressed_root(&tmp, &n);
  out->n_accels = (uint32_t)n;
  out->subdiv_level = scene->subdivisionLevel;
  out->comp_level = scene->compressionLevel;
  RefLayout lay;
  ref_layout(&lay);
  out->proj_col_major = lay.proj_col_major;
  RTCBounds b;
  rtcGetSceneBounds(hscene, &b);
  out->bounds[0] = b.lower_x; out->bounds[1] = b.lower_y; out->bounds[2] = b.lower_z;
  out->bounds[3] = b.upper_x; out->bounds[4] = b.upper_y; out->bounds[5] = b.upper_z;
  return out->root ? 0 : -2;
}

// ---- instancing: what the host hands over for a scene whose geometries are instances ----
// root of the user-geometry BVH4 (leaves = arrays of Object{geomID,primID}, kernels/geometry/object.h) and, per
// instance geometry, its world2local transform, geomID and instanced scene (kernels/common/scene_instance.h:54-56,91).
REF_API int ref_scene_instance_info(RefScene* s, uint64_t* root_out, unsigned cap, unsigned* geomIDs, float* world2local12,
                                    void** object_scenes) {
  using namespace embree;
  Scene* scene = (Scene*)s->scene;
  *root_out = 0;
  AccelData* accel = ((Accel*)s->scene)->intersectors.ptr;
  BVH4* bvh = nullptr;
  if (accel && accel->type == AccelData::TY_BVH4) bvh = (BVH4*)accel;
  else if (accel && accel->type == AccelData::TY_ACCELN) {
    AccelN* an = (AccelN*)accel;
    for (size_t i = 0; i < an->accels.size(); i++) {
      AccelData* a = an->accels[i]->intersectors.ptr;
      if (a && a->type == AccelData::TY_BVH4 && ((BVH4*)a)->primTy == &Object::type) bvh = (BVH4*)a;
    }
  }
  if (!bvh || bvh->primTy != &Object::type) return -1;
  *root_out = (uint64_t)(size_t)bvh->root;
  unsigned n = 0;
  for (size_t i = 0; i < scene->size(); i++) {
    Geometry* g = scene->get(i);
    // an Instance is an AccelSet registered as USER_GEOMETRY (kernels/common/scene_instance.cpp:31-39)
    Instance* inst = g ? dynamic_cast<Instance*>(g) : nullptr;
    if (!inst) continue;
    if (n < cap) {
      geomIDs[n] = (unsigned)i;
      const AffineSpace3fa w = inst->getWorld2Local();
      const float m[12] = {w.l.vx.x, w.l.vx.y, w.l.vx.z, w.l.vy.x, w.l.vy.y, w.l.vy.z, w.l.vz.x, w.l.vz.y, w.l.vz.z, w.p.x, w.p.y, w.p.z};
      memcpy(world2local12 + 12 * n, m, sizeof(m));
      object_scenes[n] = (void*)inst->object;
    }
    n++;
  }
  return (int)n;
}

// mixed scenes: third accessor of INTEGRATION.md -- root of the BVH4 over Triangle4v blocks that AccelN visits before the
// compressed accel (0 if the scene has no triangle meshes); fails if the triangle accel is not a BVH4<Triangle4v>
REF_API int rtcGetTriangleBVH_b200(RTCScene hscene, uint64_t* root_out) {
  RefScene tmp;
  tmp.scene = hscene;
  int ti = -1, si = -1;
  *root_out = ref_scene_triangle_root(&tmp, &ti, &si, nullptr);
  if (*root_out && si >= 0 && ti > si) return -1;  // the device path assumes triangles first (scene.cpp:650-654)
  return 0;
}

// instancing scenes: second accessor of INTEGRATION.md (instance-level BVH4 root + per-instance data)
struct RTCInstanceBVHInfo_b200 { uint64_t root; uint32_t n_instances; uint32_t reserved; float bounds[6]; };
REF_API int rtcGetInstanceBVH_b200(RTCScene hscene, RTCInstanceBVHInfo_b200* info, unsigned cap, unsigned* geomIDs, float* world2local12,
                                   RTCScene* objects) {
  RefScene tmp;
  tmp.scene = hscene;
  memset(info, 0, sizeof(*info));
  const int n = ref_scene_instance_info(&tmp, &info->root, cap, geomIDs, world2local12, (void**)objects);
  if (n < 0) return n;
  info->n_instances = (uint32_t)n;
  ref_scene_bounds(&tmp, info->bounds);
  return 0;
}

// mixed scenes: root of the BVH4 over Triangle4v blocks, and the position of both accels in the AccelN list
REF_API uint64_t ref_scene_triangle_root(RefScene* s, int* tri_index, int* subdiv_index, unsigned* sizeof_block) {
  using namespace embree;
  if (tri_index) *tri_index = -1;
  if (subdiv_index) *subdiv_index = -1;
  if (sizeof_block) *sizeof_block = (unsigned)sizeof(Triangle4v);
  AccelData* accel = ((Accel*)s->scene)->intersectors.ptr;
  uint64_t root = 0;
  if (accel && accel->type == AccelData::TY_BVH4) {
    BVH4* bvh = (BVH4*)accel;
    if (bvh->primTy == &Triangle4v::type) { root = (uint64_t)(size_t)bvh->root; if (tri_index) *tri_index = 0; }
    return root;
  }
  if (accel && accel->type == AccelData::TY_ACCELN) {
    AccelN* an = (AccelN*)accel;
    for (size_t i = 0; i < an->accels.size(); i++) {
      AccelData* a = an->accels[i]->intersectors.ptr;
      if (!a || a->type != AccelData::TY_BVH4) continue;
      BVH4* bvh = (BVH4*)a;
      if (bvh->primTy == &Triangle4v::type) { root = (uint64_t)(size_t)bvh->root; if (tri_index) *tri_index = (int)i; }
      if (bvh->primTy == &SubdivPatch1Cached::type && subdiv_index) *subdiv_index = (int)i;
    }
  }
  return root;
}
