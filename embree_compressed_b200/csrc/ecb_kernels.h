// ecb_kernels.h -- launch interface between the C-ABI host layer and the CUDA kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ecb_format.h"

#ifndef ECB_THREADS
#define ECB_THREADS 128        // threads per CTA
#endif
#ifndef ECB_BLOCKS_PER_SM
#define ECB_BLOCKS_PER_SM 8    // persistent grid = SMs x this
#endif
#ifndef ECB_MIN_BLOCKS
#define ECB_MIN_BLOCKS 5       // __launch_bounds__ occupancy target (caps registers per thread at 96)
#endif

#ifndef ECB_CHUNK
#define ECB_CHUNK 32           // ray indices a warp takes from the global queue per atomic (measured: 32 < 64 < 128)
#endif
#ifndef ECB_REFILL_IDLE_MIN
#define ECB_REFILL_IDLE_MIN 8  // idle lanes are given new rays once at least this many are idle
#endif
#ifndef ECB_BLOCK_MIN
#define ECB_BLOCK_MIN 16       // lanes wait at their block until this many have one (or no other work is left)
#endif

// One launch: where rays are read, where hits are written (may be the same array, may be mapped host memory).
struct EcbLaunch {
  const uint8_t* in;              // RTCRayHit / RTCRay records (org,tnear,dir,time,tfar in the first 36 bytes)
  size_t in_stride;
  // byte offset of tfar inside an input record: 32 (RTCRay), or 28 for the 32-byte PACKED records the host-buffer entry
  // points upload ({org.xyz, tnear}, {dir.xyz, tfar}: the fields this path reads; `time` is not one of them).  The compressed-accel
  // kernels (intersect / occluded) and the instancing kernel honour it.
  uint32_t tfar_offset = 32;
  // SoA ray input (rtcIntersectNp / rtcOccludedNp, include/embree3/rtcore_ray.h:189-206): when soa[0] != NULL the closest-hit
  // and any-hit kernels of the compressed accel read ray i from eight arrays instead of `in` -- org_x, org_y, org_z, tnear,
  // dir_x, dir_y, dir_z, tfar -- i.e. consecutive lanes load consecutive words (fully coalesced)
  const float* soa[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  uint8_t* out;                   // RTCRayHit / RTCRay records that receive tfar + hit
  size_t out_stride;
  uint64_t M;
  uint32_t instID;
  uint32_t any_hit_exact;         // 1: shadow-ray mode (first accepted leaf hit terminates the ray, only tfar = -inf is written)
  // compact result mode (host-buffer entry points): when hit_geom != NULL the kernels never touch `out`; per ray they write
  // hit_geom[i] = geomID of the hit (0xFFFFFFFF: miss / inactive; any-hit queries: 0 = occluded) and, for a hit,
  // hit_rec[i] = {tfar, u, v, primID}.  20 bytes per ray go back over PCIe instead of the 80-byte record.
  float4* hit_rec;
  uint32_t* hit_geom;
  // mixed scenes in compact result mode (triangle accel first, then the compressed accel on the same staging buffer): the
  // triangle pass writes the tfar of its hit back into the INPUT record (which the second pass reads), its hit into the
  // result arrays and the real normal into hit_ng (3 floats per ray); the second pass runs with keep_results = 1, i.e. it
  // does not reset hit_geom when it picks up a ray, and writes the dummy normal (1,0,0) with its own hits
  float* hit_ng = nullptr;
  uint32_t keep_results = 0;
  uint32_t* hit_inst = nullptr;    // instancing scenes in compact result mode: instID[0] of the hit (geomID of the instance)
  // batching thresholds of the lane kernel (0 = the built-in ECB_BLOCK_MIN / ECB_REFILL_IDLE_MIN): lanes wait at their block until
  // block_min lanes have one; idle lanes are refilled once refill_min are idle.  Coherent rays want full block batches;
  // incoherent ones (few rays reach a block, most end after one or two node tests) want their lanes back sooner.
  uint32_t block_min = 0, refill_min = 0;
  uint32_t* block_out;            // optional: index of the block of the final hit per ray
  unsigned long long* counters;   // optional: instrumented variant (6 counters)
  unsigned long long* queue;      // device word used as the dynamic ray-fetch cursor (zeroed by the launcher)
};

#ifndef ECB_DEFAULT_KERNEL_MODE
#define ECB_DEFAULT_KERNEL_MODE 0  // closest-hit kernel of ecb_launch_intersect: 0 = one lane per ray, 1 = four lanes per ray
#endif
int ecb_kernel_mode();
void ecb_set_kernel_mode(int mode);  // 0 lane, 1 quad, -1 back to the default / ECB200_KERNEL
cudaError_t ecb_launch_intersect(const EcbDeviceScene& sc, int device, const EcbLaunch& L, cudaStream_t stream);
cudaError_t ecb_launch_occluded(const EcbDeviceScene& sc, int device, const EcbLaunch& L, cudaStream_t stream);
cudaError_t ecb_launch_instanced(const EcbDeviceInstScene& is, int device, const EcbLaunch& L, bool any_hit, cudaStream_t stream);
// mixed scenes: the triangle accel is traversed first, in place (L.in == L.out), then the compressed accel on the same rays
cudaError_t ecb_launch_triangles(const EcbDeviceTriScene& ts, int device, const EcbLaunch& L, bool any_hit, cudaStream_t stream);

// -DECB_BOUNDS_CHECK builds: violation counters of the kernels' index checks (8 classes, see ecb_kernels.cu); enabled = 0 otherwise
cudaError_t ecb_bounds_check_read(unsigned long long out[8], int* enabled);
// measurement support: read bandwidth of a working set of `bytes` streamed `passes` times (GB/s, best of 5)
cudaError_t ecb_probe_read_bandwidth(size_t bytes, int passes, int device, double* gbs_out);

// block construction (csrc/ecb_build.cu): all pointers are device pointers
struct EcbBuildLaunch {
  const float* grids;        // n x (2^C+1)^2 x 3 displaced vertices, row-major
  const float* space;        // n x 9 block frames (world -> local)
  const float* world;        // n x 9 their inverses [vx vy vz]
  const float* uv;           // n x 4 uv of vertex (0,0) and uv extent
  const uint32_t* geomIDs;
  const uint32_t* primIDs;
  uint8_t* blocks;           // n x stride, zero-initialised by the caller; receives blocks in the device format
  float* bounds;             // n x 6 conservative world bounds
  uint32_t* patch_ok;        // optional, n
  void* tmp;                 // ecb_build_tmp_bytes(n) bytes of scratch
  uint32_t n, comp_level, leaf_kind, stride, leaf_offset;
};
size_t ecb_build_tmp_bytes(uint32_t n);
cudaError_t ecb_launch_build(const EcbBuildLaunch& b, int device, cudaStream_t stream);
