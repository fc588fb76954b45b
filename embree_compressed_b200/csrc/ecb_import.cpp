// ecb_import.cpp -- converts the reference's committed compressed BVH4 into the device scene format.
//
// The importer reads the reference's memory directly through a hard-coded description of its layout (no reference
// headers are included; the layout is probe-verified against the compiled reference by
// tests/test_oracle_cpu.py::test_layout_probe_matches_importer_constants):
//   BVH4 NodeRef            kernels/bvh/bvh.h:107-140,150-400   low 4 bits = type, tyLeaf = 8, emptyNode = 8
//   BVH4::AlignedNode       kernels/bvh/bvh.h:403-430,588-594   children[4] @0 (8 B each), 6 x vfloat4 bounds @32
//   leaf                    encodeTypedLeaf(cbvh,1): one CompressedBVH per leaf (kernels/bvh/bvh_builder_subdiv.cpp:727-728)
//   CompressedBVH<node_t>   kernels/geometry/compressed.h:408-433,340-376  184-byte header, then nodes/leaves/grid
// Host only; runs without a GPU.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <new>
#include <stdexcept>

#include "ecb_scene.h"

namespace {

// ---- reference layout ----
constexpr uint64_t kAlignMask = 15, kTyLeaf = 8, kEmpty = 8;
constexpr size_t kAN_Children = 0, kAN_Bounds = 32, kAN_Size = 128;
constexpr size_t kB_GeomID = 0, kB_PrimID = 4, kB_UV = 8, kB_Nodes = 32, kB_Leaves = 40, kB_Grid = 48, kB_Extent = 56,
                 kB_Elems = 60, kB_GridWidth = 64, kB_Space = 68, kB_Proj = 104, kB_Box = 140, kB_Size = 184;

template <typename T>
T rd(const uint8_t* p) {
  T v;
  memcpy(&v, p, sizeof(T));
  return v;
}

// Range checks of a relocated image: with base != NULL and image_bytes != 0 every offset read from the image must stay
// inside it (a truncated or corrupt image is rejected, never dereferenced).  Live memory (base == NULL) is the host
// application's own committed BVH and is trusted like the reference trusts it.  `nodes_left` bounds the number of nodes an
// import may visit: every distinct node occupies 128 bytes of the image, so a reference cycle (which would otherwise fan
// out 4^depth) runs out of budget instead of memory.
struct Range {
  const uint8_t* base;
  uint64_t bytes;  // 0 = unchecked
  uint64_t nodes_left;
  Range(const void* b, uint64_t n) : base(static_cast<const uint8_t*>(b)), bytes(b ? n : 0), nodes_left(b && n ? n / 128 + 1 : (1ull << 28)) {}
  bool ok(uint64_t off, uint64_t len) const { return bytes == 0 || (off <= bytes && len <= bytes - off); }
};

struct Importer {
  const ecbImportDesc& d;
  ecbScene& s;
  const uint8_t* base;
  Range range;
  uint32_t max_depth = 0;
  uint64_t ref_bytes = 0;
  std::string err;

  Importer(const ecbImportDesc& desc, ecbScene& scene)
      : d(desc), s(scene), base(static_cast<const uint8_t*>(desc.base)), range(desc.base, desc.image_bytes) {}

  const uint8_t* at(uint64_t ref) const { return base + (ref & ~kAlignMask); }
  uint32_t outside() {
    err = "image offset out of range (truncated or corrupt image)";
    return ECB_REF_EMPTY;
  }

  uint32_t import_block(uint64_t ref) {
    if (!range.ok(ref & ~kAlignMask, kB_Size)) return outside();
    const uint8_t* b = at(ref);
    const uint32_t elems = rd<uint32_t>(b + kB_Elems);
    const uint32_t width = rd<uint32_t>(b + kB_GridWidth);
    if (elems != s.elems || width != s.grid_width) {
      err = "block with elems=" + std::to_string(elems) + " grid_width=" + std::to_string(width) +
            " does not match compression level " + std::to_string(d.comp_level);
      return ECB_REF_EMPTY;
    }
    const size_t index = s.blocks.size() / s.block_stride;
    s.blocks.resize(s.blocks.size() + s.block_stride, 0);
    uint8_t* out = s.blocks.data() + index * s.block_stride;
    EcbBlockHeader h;
    memset(&h, 0, sizeof(h));
    for (int i = 0; i < 9; i++) h.space[i] = rd<float>(b + kB_Space + 4 * i);
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) h.proj[3 * r + c] = rd<float>(b + kB_Proj + 4 * (d.proj_col_major ? 3 * c + r : 3 * r + c));
    for (int i = 0; i < 10; i++) h.box[i] = rd<float>(b + kB_Box + 4 * i);
    for (int i = 0; i < 4; i++) h.uv[i] = rd<float>(b + kB_UV + 4 * i);
    h.extent = rd<float>(b + kB_Extent);
    h.geomID = rd<uint32_t>(b + kB_GeomID);
    h.primID = rd<uint32_t>(b + kB_PrimID);
    const size_t n_cells = 3 * (size_t)elems + 1;
    if (!range.ok(rd<uint64_t>(b + kB_Nodes), 4 * (uint64_t)elems) ||
        (d.leaf_kind == ECB_LEAF_LEAF && !range.ok(rd<uint64_t>(b + kB_Leaves), 2 * (uint64_t)n_cells)) ||
        (d.leaf_kind == ECB_LEAF_GRID && !range.ok(rd<uint64_t>(b + kB_Grid), 12 * (uint64_t)width * width))) {
      s.blocks.resize(s.blocks.size() - s.block_stride);
      return outside();
    }
    const uint8_t* nodes = base + rd<uint64_t>(b + kB_Nodes);
    h.root_word = rd<uint32_t>(nodes);
    memcpy(out, &h, sizeof(h));
    if (elems > 1) memcpy(out + ECB_BLOCK_NODES_OFFSET, nodes + 4, 4 * (size_t)(elems - 1));
    const size_t cells = 3 * (size_t)elems + 1;
    if (d.leaf_kind == ECB_LEAF_LEAF) {
      memcpy(out + s.leaf_offset, base + rd<uint64_t>(b + kB_Leaves), 2 * cells);
      ref_bytes += 2 * cells;
    } else if (d.leaf_kind == ECB_LEAF_GRID) {
      memcpy(out + s.leaf_offset, base + rd<uint64_t>(b + kB_Grid), 12 * (size_t)width * width);
      ref_bytes += 12 * (size_t)width * width;
    }
    ref_bytes += kB_Size + 4 * (size_t)elems;
    return ECB_REF_LEAF | (uint32_t)index;
  }

  // pre-order DFS keeps subtrees (and the blocks below them) contiguous in memory
  uint32_t import_ref(uint64_t ref, uint32_t depth) {
    if (!err.empty()) return ECB_REF_EMPTY;
    if (ref == kEmpty) return ECB_REF_EMPTY;
    if (ref & kTyLeaf) {
      const uint64_t num = (ref & kAlignMask) - kTyLeaf;
      if (num != 1) {
        err = "leaf with " + std::to_string(num) + " blocks (the compressed builder emits exactly 1)";
        return ECB_REF_EMPTY;
      }
      if (depth > max_depth) max_depth = depth;
      return import_block(ref);
    }
    if ((ref & kAlignMask) != 0) {
      err = "node type " + std::to_string(ref & kAlignMask) + " is not an AlignedNode (outside the compressed path)";
      return ECB_REF_EMPTY;
    }
    if (depth > 512) {
      err = "top-level BVH deeper than 512";
      return ECB_REF_EMPTY;
    }
    if (!range.ok(ref, kAN_Size)) return outside();
    if (range.nodes_left-- == 0) {
      err = "more nodes than the image can hold (reference cycle)";
      return ECB_REF_EMPTY;
    }
    const uint8_t* n = at(ref);
    const uint32_t index = (uint32_t)s.top.size();
    s.top.emplace_back();
    ref_bytes += kAN_Size;
    EcbTopNode node;
    memcpy(node.bounds, n + kAN_Bounds, 96);
    for (int i = 0; i < 4; i++) node.child[i] = import_ref(rd<uint64_t>(n + kAN_Children + 8 * i), depth + 1);
    s.top[index] = node;
    return index;
  }
};

}  // namespace

ecbError ecb_import_bvh4(const ecbImportDesc* desc, ecbScene** out) {
  if (!desc || !out) return ECB_ERROR_INVALID_ARGUMENT;
  if (desc->leaf_kind > ECB_LEAF_BOX) {
    ecb_set_error("unknown leaf kind");
    return ECB_ERROR_INVALID_ARGUMENT;
  }
  if (desc->comp_level < 1 || desc->comp_level > ECB_MAX_C) {
    ecb_set_error("compression level must be in 1.." + std::to_string(ECB_MAX_C));
    return ECB_ERROR_INVALID_ARGUMENT;
  }
  ecbScene* s = new (std::nothrow) ecbScene;
  if (!s) return ECB_ERROR_OUT_OF_MEMORY;
  const uint32_t C = desc->comp_level;
  const uint32_t cells = 1u << (2 * C);
  s->elems = (cells - 1) / 3;
  s->grid_width = (1u << C) + 1;
  s->rcp_edges = 1.0f / (float)(1u << C);
  ecb_block_layout(C, desc->leaf_kind, &s->leaf_offset, &s->block_stride);

  Importer imp(*desc, *s);
  try {
    s->root = imp.import_ref(desc->root, 0);
  } catch (const std::bad_alloc&) {
    delete s;
    return ECB_ERROR_OUT_OF_MEMORY;
  }
  if (!imp.err.empty()) {
    ecb_set_error(imp.err);
    delete s;
    return ECB_ERROR_UNSUPPORTED;
  }
  if (3 * imp.max_depth + 4 > ECB_MAX_TOP_STACK) {
    ecb_set_error("top-level BVH depth " + std::to_string(imp.max_depth) + " exceeds the traversal stack");
    delete s;
    return ECB_ERROR_UNSUPPORTED;
  }
  ecbSceneInfo& in = s->info;
  in.leaf_kind = desc->leaf_kind;
  in.subdiv_level = desc->subdiv_level;
  in.comp_level = C;
  in.n_top_nodes = (uint32_t)s->top.size();
  in.n_blocks = (uint32_t)(s->blocks.size() / s->block_stride);
  in.top_depth = imp.max_depth;
  in.block_stride = s->block_stride;
  in.top_node_bytes = sizeof(EcbTopNode);
  in.device_bytes = s->top.size() * sizeof(EcbTopNode) + s->blocks.size();
  in.reference_bytes = imp.ref_bytes;
  in.micro_triangles = 2ull * cells * in.n_blocks;
  memcpy(in.bounds, desc->bounds, sizeof(in.bounds));
  in.device = -1;
  *out = s;
  return ECB_SUCCESS;
}

// ---- instancing scene: BVH4 over Object{geomID,primID} arrays (kernels/geometry/object.h, 8 bytes per primitive) ----
namespace {
struct InstImporter {
  const ecbInstanceDesc& d;
  ecbScene& s;
  const uint8_t* base;
  Range range;
  std::vector<uint32_t> geom_to_instance;
  uint32_t max_depth = 0;
  std::string err;
  InstImporter(const ecbInstanceDesc& desc, ecbScene& scene)
      : d(desc), s(scene), base(static_cast<const uint8_t*>(desc.base)), range(desc.base, desc.image_bytes) {}

  uint32_t import_ref(uint64_t ref, uint32_t depth) {
    if (!err.empty() || ref == kEmpty) return ECB_REF_EMPTY;
    if (ref & kTyLeaf) {
      const uint32_t num = (uint32_t)((ref & kAlignMask) - kTyLeaf);
      if (num > 7) { err = "instance leaf with more than 7 primitives"; return ECB_REF_EMPTY; }
      if (!range.ok(ref & ~kAlignMask, 8 * (uint64_t)num)) { err = "image offset out of range (truncated or corrupt image)"; return ECB_REF_EMPTY; }
      const uint8_t* prims = base + (ref & ~kAlignMask);
      const uint32_t first = (uint32_t)s.items.size();
      uint32_t kept = 0;
      for (uint32_t i = 0; i < num; i++) {
        const uint32_t geomID = rd<uint32_t>(prims + 8 * i);
        if (geomID >= geom_to_instance.size() || geom_to_instance[geomID] == 0xFFFFFFFFu) {
          err = "user geometry " + std::to_string(geomID) + " is not an instance handed to ecbSceneImportInstances";
          return ECB_REF_EMPTY;
        }
        s.items.push_back(geom_to_instance[geomID]);
        kept++;
      }
      if (first >= (1u << 27)) { err = "too many instance primitives"; return ECB_REF_EMPTY; }
      if (depth > max_depth) max_depth = depth;
      return ECB_REF_LEAF | (first << 3) | kept;
    }
    if ((ref & kAlignMask) != 0) { err = "instance BVH node is not an AlignedNode (use object_accel=bvh4.object)"; return ECB_REF_EMPTY; }
    if (depth > 512) { err = "instance BVH deeper than 512"; return ECB_REF_EMPTY; }
    if (!range.ok(ref, kAN_Size) || range.nodes_left-- == 0) { err = "image offset out of range or reference cycle (truncated or corrupt image)"; return ECB_REF_EMPTY; }
    const uint8_t* n = base + (ref & ~kAlignMask);
    const uint32_t index = (uint32_t)s.top.size();
    s.top.emplace_back();
    EcbTopNode node;
    memcpy(node.bounds, n + kAN_Bounds, 96);
    for (int i = 0; i < 4; i++) node.child[i] = import_ref(rd<uint64_t>(n + kAN_Children + 8 * i), depth + 1);
    s.top[index] = node;
    return index;
  }
};
}  // namespace

static ecbError import_instances_impl(const ecbInstanceDesc* desc, ecbScene** out);
ecbError ecb_import_instances(const ecbInstanceDesc* desc, ecbScene** out) {
  try {
    return import_instances_impl(desc, out);
  } catch (const std::exception&) {  // std::bad_alloc from a huge geomID / instance count: never across the C ABI
    ecb_set_error("out of memory while importing the instancing scene");
    return ECB_ERROR_OUT_OF_MEMORY;
  }
}
static ecbError import_instances_impl(const ecbInstanceDesc* desc, ecbScene** out) {
  if (!desc || !out || (desc->n_instances && (!desc->geomIDs || !desc->world2local || !desc->objects))) return ECB_ERROR_INVALID_ARGUMENT;
  ecbScene* s = new (std::nothrow) ecbScene;
  if (!s) return ECB_ERROR_OUT_OF_MEMORY;
  s->instanced = true;
  struct Guard {  // an exception below must not leak the half-built scene or the references it took
    ecbScene* s;
    ~Guard() {
      if (!s) return;
      for (ecbScene* o : s->objects) o->refs--;
      delete s;
    }
  } guard{s};
  InstImporter imp(*desc, *s);
  uint32_t max_geom = 0;
  for (uint32_t i = 0; i < desc->n_instances; i++) max_geom = desc->geomIDs[i] > max_geom ? desc->geomIDs[i] : max_geom;
  imp.geom_to_instance.assign((size_t)max_geom + 1, 0xFFFFFFFFu);
  for (uint32_t i = 0; i < desc->n_instances; i++) {
    ecbScene* obj = desc->objects[i];
    if (!obj || obj->instanced || obj->has_triangles()) {
      ecb_set_error("instance " + std::to_string(i) + " does not refer to an imported compressed-only (non-instanced, no triangle accel) scene");
      return ECB_ERROR_INVALID_ARGUMENT;  // the guard releases the scene
    }
    uint32_t oi = 0;
    for (; oi < s->objects.size(); oi++)
      if (s->objects[oi] == obj) break;
    if (oi == s->objects.size()) {
      obj->refs++;
      s->objects.push_back(obj);
    }
    EcbInstance in;
    memset(&in, 0, sizeof(in));
    memcpy(in.world2local, desc->world2local + 12 * (size_t)i, 48);
    in.geomID = desc->geomIDs[i];
    in.object = oi;
    imp.geom_to_instance[desc->geomIDs[i]] = (uint32_t)s->instances.size();
    s->instances.push_back(in);
  }
  s->root = imp.import_ref(desc->root, 0);
  if (imp.err.empty() && 3 * imp.max_depth + 4 > ECB_MAX_TOP_STACK) imp.err = "instance BVH too deep for the traversal stack";
  if (!imp.err.empty()) {
    ecb_set_error(imp.err);
    return ECB_ERROR_UNSUPPORTED;  // the guard releases the scene
  }
  ecbSceneInfo& in = s->info;
  memset(&in, 0, sizeof(in));
  in.n_top_nodes = (uint32_t)s->top.size();
  in.top_depth = imp.max_depth;
  in.top_node_bytes = sizeof(EcbTopNode);
  in.device_bytes = s->top.size() * sizeof(EcbTopNode) + s->items.size() * 4 + s->instances.size() * sizeof(EcbInstance);
  for (const EcbInstance& i : s->instances) {
    in.micro_triangles += s->objects[i.object]->info.micro_triangles;
    in.n_blocks += s->objects[i.object]->info.n_blocks;
  }
  for (ecbScene* o : s->objects) in.device_bytes += o->info.device_bytes;
  in.leaf_kind = s->objects.empty() ? 0 : s->objects[0]->info.leaf_kind;
  memcpy(in.bounds, desc->bounds, sizeof(in.bounds));
  in.device = -1;
  guard.s = nullptr;
  *out = s;
  return ECB_SUCCESS;
}

// ---- mixed scenes: BVH4 over Triangle4v blocks (kernels/geometry/trianglev.h) ----
namespace {
constexpr size_t kTri_Size = 176;
struct TriImporter {
  ecbScene& s;
  const uint8_t* base;
  Range range;
  uint32_t max_depth = 0;
  std::string err;
  TriImporter(const ecbTriangleDesc& desc, ecbScene& scene) : s(scene), base(static_cast<const uint8_t*>(desc.base)), range(desc.base, desc.image_bytes) {}

  uint32_t import_ref(uint64_t ref, uint32_t depth) {
    if (!err.empty() || ref == kEmpty) return ECB_REF_EMPTY;
    if (ref & kTyLeaf) {
      const uint32_t num = (uint32_t)((ref & kAlignMask) - kTyLeaf);
      const uint32_t first = (uint32_t)s.tri_blocks.size();
      if (first >= (1u << 27)) { err = "too many triangle blocks"; return ECB_REF_EMPTY; }
      if (!range.ok(ref & ~kAlignMask, kTri_Size * (uint64_t)num)) { err = "image offset out of range (truncated or corrupt image)"; return ECB_REF_EMPTY; }
      const uint8_t* blocks = base + (ref & ~kAlignMask);
      for (uint32_t i = 0; i < num; i++) {
        EcbTriBlock b;
        memcpy(&b, blocks + kTri_Size * i, kTri_Size);
        s.tri_blocks.push_back(b);
      }
      if (depth > max_depth) max_depth = depth;
      return ECB_REF_LEAF | (first << 3) | num;
    }
    if ((ref & kAlignMask) != 0) { err = "triangle BVH node is not an AlignedNode (use tri_accel=bvh4.triangle4v)"; return ECB_REF_EMPTY; }
    if (depth > 512) { err = "triangle BVH deeper than 512"; return ECB_REF_EMPTY; }
    if (!range.ok(ref, kAN_Size) || range.nodes_left-- == 0) { err = "image offset out of range or reference cycle (truncated or corrupt image)"; return ECB_REF_EMPTY; }
    const uint8_t* n = base + (ref & ~kAlignMask);
    const uint32_t index = (uint32_t)s.tri_top.size();
    s.tri_top.emplace_back();
    EcbTopNode node;
    memcpy(node.bounds, n + kAN_Bounds, 96);
    for (int i = 0; i < 4; i++) node.child[i] = import_ref(rd<uint64_t>(n + kAN_Children + 8 * i), depth + 1);
    s.tri_top[index] = node;
    return index;
  }
};
}  // namespace

ecbError ecb_import_triangles(const ecbTriangleDesc* desc, ecbScene* s) {
  if (!desc || !s || !desc->root) return ECB_ERROR_INVALID_ARGUMENT;
  if (s->instanced || s->resident() || s->has_triangles() || s->refs.load() > 1) {
    ecb_set_error("triangles attach to a freshly imported, non-instancing scene that is not yet resident or instanced");
    return ECB_ERROR_INVALID_OPERATION;
  }
  TriImporter imp(*desc, *s);
  uint32_t root;
  try {
    root = imp.import_ref(desc->root, 0);
  } catch (const std::bad_alloc&) {
    s->tri_top.clear();
    s->tri_blocks.clear();
    return ECB_ERROR_OUT_OF_MEMORY;
  }
  if (imp.err.empty() && 3 * imp.max_depth + 4 > ECB_MAX_TOP_STACK) imp.err = "triangle BVH too deep for the traversal stack";
  if (!imp.err.empty()) {
    ecb_set_error(imp.err);
    s->tri_top.clear();
    s->tri_blocks.clear();
    return ECB_ERROR_UNSUPPORTED;
  }
  s->tri_root = root;
  s->tri_depth = imp.max_depth;
  s->info.device_bytes += s->tri_top.size() * sizeof(EcbTopNode) + s->tri_blocks.size() * sizeof(EcbTriBlock);
  s->info.reference_bytes += s->tri_top.size() * kAN_Size + s->tri_blocks.size() * kTri_Size;
  return ECB_SUCCESS;
}
