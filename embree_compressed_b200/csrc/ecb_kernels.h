// ecb_kernels.h -- launch interface between the C-ABI host layer and the CUDA kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ecb_format.h"

#define ECB_THREADS 128        // threads per CTA
#define ECB_BLOCKS_PER_SM 8    // persistent grid = SMs x this

cudaError_t ecb_launch_intersect(const EcbDeviceScene& sc, int device, void* d_rayhit, uint64_t M, size_t stride, uint32_t instID,
                                 uint32_t* d_block_out, unsigned long long* d_counters, cudaStream_t stream);
cudaError_t ecb_launch_occluded(const EcbDeviceScene& sc, int device, void* d_ray, uint64_t M, size_t stride, cudaStream_t stream);
