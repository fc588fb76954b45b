// ecb_format.h -- device scene format of the B200 compressed-BVH traversal (shared by host importer and kernels).
//
// Converted from the reference's committed memory (128-byte BVH4::AlignedNode with 64-bit tagged child pointers,
// kernels/bvh/bvh.h:433-594; 184-byte CompressedBVH header + payload, kernels/geometry/compressed.h:340-376,408-433)
// into two flat, 16-byte aligned, pointer-free arrays:
//
//   top nodes  : 112 B each  = 6 x float4 child bounds (same SoA order as the reference) + 4 x 32-bit child refs
//   blocks     : fixed stride per scene = 144 B header | quantized nodes 1..elems-1 (4 B each, 16-B aligned so that
//                the four children words of ANY node are one aligned 16-byte load) | leaf payload (2 B per cell for
//                cLeaf, 12 B per vertex for cGrid, nothing for cBox), padded to 16 B.
//
// Per-scene constants (C, elems, grid width, rcp_edges) are NOT repeated per block.
#pragma once
#include <stdint.h>

#define ECB_REF_EMPTY 0xFFFFFFFFu  // empty child (reference: BVH4::emptyNode, bvh.h:127)
#define ECB_REF_LEAF 0x80000000u   // bit 31: child is a block; low 31 bits = block index
#define ECB_MAX_TOP_STACK 96       // entries of the per-ray top-level stack (importer checks 3*depth+1 <= this)
#define ECB_MAX_C 5                // compression levels supported by the kernels (reference tutorials clamp to 1..4)

struct EcbTopNode {
  float bounds[6][4];  // lower_x[4], upper_x[4], lower_y[4], upper_y[4], lower_z[4], upper_z[4]
  uint32_t child[4];
};
static_assert(sizeof(EcbTopNode) == 112, "top node is 7 x 16 bytes");

struct EcbBlockHeader {
  float space[9];   // LinearSpace3f columns vx,vy,vz: space[3*c+k]   (compressed.h:425)
  float proj[9];    // homography, row-major proj[3*r+c]              (compressed.h:429, layout normalised)
  float box[10];    // zmin, zmax, 4 frustum corners xy               (compressed.h:432, :277-292)
  float uv[4];      // uv0.x, uv0.y, extent.x, extent.y               (compressed.h:414, :84-86)
  float extent;     // cLeaf height-range extension                   (compressed.h:420, :312-313)
  uint32_t geomID;
  uint32_t primID;
  uint32_t root_word;  // quantized node 0: bytes xz | x<<8 | yz<<16 | y<<24  (compressed_node.h:264-268)
};
static_assert(sizeof(EcbBlockHeader) == 144, "block header is 9 x 16 bytes");

#define ECB_BLOCK_NODES_OFFSET 144u  // nodes 1..elems-1 start here (node n at +4*(n-1))

// Everything a kernel needs; passed by value.
struct EcbDeviceScene {
  const EcbTopNode* top;
  const uint8_t* blocks;
  uint32_t root;          // top-level root ref (node index, leaf ref, or ECB_REF_EMPTY)
  uint32_t n_top, n_blocks;
  uint32_t block_stride;  // bytes
  uint32_t leaf_offset;   // byte offset of the leaf payload inside a block
  uint32_t leaf_kind;     // ecbLeafKind
  uint32_t comp_level;    // C
  uint32_t elems;         // (4^C-1)/3 inner nodes per block
  uint32_t grid_width;    // 2^C+1
  float rcp_edges;        // 1/2^C
  uint32_t top_depth;     // depth of the top-level BVH4 (sizes the shared-memory stack of the quad kernel: 3*depth+1 entries)
  // experiment (-DECB_PRETEST builds with ECB200_PRETEST=1 only, NULL otherwise): per top-level node 48 bytes = the six planes of
  // its four children as bfloat16, rounded OUTWARD.  A ray that misses all four widened boxes cannot hit the fp32 ones
  // (IEEE subtraction, multiplication and min / max are monotone), so the 112-byte node need not be fetched.
  const void* pre;
};

// ---- instanced scenes: a BVH4 over instances of object scenes ----
// Instance-level nodes reuse EcbTopNode; a leaf ref is ECB_REF_LEAF | first_item << 3 | count (count <= 7 Object
// primitives per leaf, kernels/bvh/bvh.h:140); items[] holds the instance index of every Object primitive.
struct EcbInstance {
  float world2local[12];  // vx, vy, vz, p of Instance::world2local0 (kernels/common/scene_instance.h:91)
  uint32_t geomID;        // reported as hit.instID
  uint32_t object;        // index into EcbDeviceInstScene::objects
  uint32_t pad[2];
};
static_assert(sizeof(EcbInstance) == 64, "instance record is 4 x 16 bytes");

struct EcbDeviceInstScene {
  const EcbTopNode* nodes;
  const uint32_t* items;
  const EcbInstance* instances;
  const EcbDeviceScene* objects;
  uint32_t root;
  uint32_t n_nodes, n_items, n_instances, n_objects;
  int32_t uniform_leaf_kind;  // leaf kind shared by all object scenes, or -1 (selects the kernel instantiation)
};

// ---- triangle meshes of a mixed scene: the host's BVH4 over Triangle4v blocks ----
// Nodes reuse EcbTopNode; a leaf ref is ECB_REF_LEAF | first_block << 3 | count (count <= 7 blocks per leaf,
// kernels/bvh/bvh.h:140,380-384).  A block is the reference's TriangleMv<4> verbatim (kernels/geometry/trianglev.h):
// already pointer-free, SoA and a multiple of 16 bytes.
struct EcbTriBlock {
  float v0[3][4], v1[3][4], v2[3][4];  // x[4], y[4], z[4] of each vertex
  uint32_t geomID[4];                  // 0xFFFFFFFF = unused lane (its vertices are zero)
  uint32_t primID[4];
};
static_assert(sizeof(EcbTriBlock) == 176, "triangle block is 11 x 16 bytes");

struct EcbDeviceTriScene {
  const EcbTopNode* nodes;
  const EcbTriBlock* blocks;
  uint32_t root;  // ECB_REF_EMPTY: the scene has no triangle accel
  uint32_t n_nodes, n_blocks;
};

// native blob header (ecbSceneSave / ecbSceneLoad)
struct EcbBlobHeader {
  char magic[8];  // "ECB200S1"
  uint32_t version, leaf_kind, subdiv_level, comp_level;
  uint32_t n_top, n_blocks, block_stride, leaf_offset;
  uint32_t root, top_depth, elems, grid_width;
  float rcp_edges;
  float bounds[6];
  uint32_t reserved;
  uint64_t reference_bytes;
  uint64_t top_bytes, block_bytes;
};
// optional trailer of a blob: the triangle accel of a mixed scene (nodes, then blocks)
struct EcbBlobTriTrailer {
  char magic[8];  // "ECB200T1"
  uint32_t n_nodes, n_blocks, root, depth;
};
