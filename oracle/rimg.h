/* oracle/rimg.h -- header of a relocatable image ("RIMG") of the reference's committed compressed accel.
 *
 * An RIMG is the reference's own memory, byte for byte (128-byte BVH4::AlignedNode, 184-byte CompressedBVH header
 * followed by its 4-byte nodes and 2-byte leaves / 12-byte grid vertices), copied into one buffer with every
 * pointer (child NodeRefs, Block::nodes/leaves/grid) replaced by its byte offset from the start of the buffer, tag
 * bits kept.  Produced by oracle/refhost.cpp:ref_scene_dump_image (test infrastructure) and consumed by
 *   - the oracle port (cbvh_scene{base = image, root = header.root}),
 *   - the product's importer ecbSceneImportBVH4 (same base/root pair) -- exactly the call a host application makes
 *     on live memory with base = 0.
 * The header is plain data; this file contains no reference code.
 */
#pragma once
#include <stdint.h>

#define RIMG_MAGIC "ECBRIMG1"

typedef struct rimg_header {
  char magic[8];
  uint32_t version;        /* 1 */
  uint32_t leaf_kind;      /* 0 grid, 1 leaf, 2 box */
  uint32_t subdiv_level;   /* S */
  uint32_t comp_level;     /* C */
  uint32_t proj_col_major; /* Eigen::Matrix3f storage order of the build that produced the image */
  uint32_t n_nodes;        /* top-level AlignedNodes */
  uint32_t n_blocks;       /* CompressedBVH blocks */
  uint32_t reserved0;
  uint64_t root;           /* relocated root NodeRef */
  uint64_t total_bytes;
  float bounds[6];         /* scene bounds lower xyz, upper xyz */
  uint64_t tri_root;       /* 0, or the relocated root NodeRef of the scene's BVH4 over Triangle4v blocks (176 B each,
                              kernels/geometry/trianglev.h) -- mixed scenes, SURVEY 8f-1 */
  uint32_t reserved1[2];
} rimg_header;             /* 96 bytes */
