// ecb_scene.h -- host-side scene object behind the opaque ecbScene handle.
#pragma once
#include <atomic>
#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ecb200.h"
#include "ecb_format.h"

struct ecbScene {
  // converted scene (host copy, pointer-free)
  std::vector<EcbTopNode> top;
  std::vector<uint8_t> blocks;
  ecbSceneInfo info{};
  uint32_t root = ECB_REF_EMPTY;
  uint32_t block_stride = 0, leaf_offset = 0, elems = 0, grid_width = 0;
  float rcp_edges = 0.f;

  // device residency: one Residency per CUDA device the scene has been uploaded to (ecbSceneUpload / ecbSceneUploadMulti);
  // res[0] is the primary one (lazy uploads and single-device calls use it).  `mutex` guards the vector.
  mutable std::mutex mutex;
  std::vector<struct Residency*> res;
  std::atomic<struct Residency*> primary{nullptr};
  bool resident() const { return primary.load(std::memory_order_acquire) != nullptr; }

  // mixed scenes: the triangle accel that AccelN visits before the compressed accel
  std::vector<EcbTopNode> tri_top;
  std::vector<EcbTriBlock> tri_blocks;
  uint32_t tri_root = ECB_REF_EMPTY;
  uint32_t tri_depth = 0;
  bool has_triangles() const { return tri_root != ECB_REF_EMPTY; }

  // instanced scenes: instance-level BVH4 (in `top`), Object items, instance records, retained object scenes
  bool instanced = false;
  std::vector<uint32_t> items;
  std::vector<EcbInstance> instances;
  std::vector<ecbScene*> objects;
  std::atomic<int> refs{1};

};

// byte layout of one block in the device format for compression level C and leaf kind (ecb_format.h)
inline void ecb_block_layout(uint32_t C, uint32_t leaf_kind, uint32_t* leaf_offset, uint32_t* stride) {
  const uint32_t cells = 1u << (2 * C), elems = (cells - 1) / 3, width = (1u << C) + 1;
  size_t off = ECB_BLOCK_NODES_OFFSET + 4 * (size_t)(elems - 1);
  off = (off + 15) & ~size_t(15);
  *leaf_offset = (uint32_t)off;
  if (leaf_kind == ECB_LEAF_LEAF) off += 2 * (size_t)cells;
  if (leaf_kind == ECB_LEAF_GRID) off += 12 * (size_t)width * width;
  *stride = (uint32_t)((off + 15) & ~size_t(15));
}

void ecb_set_error(const std::string& msg);
ecbError ecb_import_bvh4(const ecbImportDesc* desc, ecbScene** out);
ecbError ecb_import_instances(const ecbInstanceDesc* desc, ecbScene** out);
ecbError ecb_import_triangles(const ecbTriangleDesc* desc, ecbScene* scene);
