// tests/mock_cuda/mock_kernels.cpp -- TEST INFRASTRUCTURE ONLY.  Stand-ins for the kernel launchers of
// embree_compressed_b200/csrc/ecb_kernels.cu with the same launch interface (csrc/ecb_kernels.h): each "launch" is
// enqueued on the mock stream and fills the outputs the real kernel would fill -- in place or as compact result arrays,
// packed or full input records, keep_results -- with fake_trace.h's deterministic results.
#include "../../embree_compressed_b200/csrc/ecb_kernels.h"
#include "fake_trace.h"

namespace {

struct RayIn { float o[3], tnear, d[3], tfar; };
RayIn load_ray(const EcbLaunch& L, uint64_t i) {
  RayIn r;
  if (L.soa[0]) {  // SoA ray input (ecbIntersectNp / ecbOccludedNp)
    for (int k = 0; k < 3; k++) { r.o[k] = L.soa[k][i]; r.d[k] = L.soa[4 + k][i]; }
    r.tnear = L.soa[3][i];
    r.tfar = L.soa[7][i];
    return r;
  }
  const uint8_t* rp = L.in + i * L.in_stride;
  memcpy(r.o, rp, 12);
  memcpy(&r.tnear, rp + 12, 4);
  memcpy(r.d, rp + 16, 12);
  memcpy(&r.tfar, rp + L.tfar_offset, 4);
  return r;
}
const float kNegInf = -__builtin_inff();

void closest_pass(int pass, const EcbLaunch L) {
  for (uint64_t i = 0; i < L.M; i++) {
    const RayIn r = load_ray(L, i);
    if (L.hit_geom && !L.keep_results) L.hit_geom[i] = 0xFFFFFFFFu;
    FakeHit h;
    if (L.any_hit_exact) {  // shadow-ray mode of the closest-hit kernel
      if (fake_any(pass, r.o, r.tnear, r.d, r.tfar)) {
        if (L.hit_geom) L.hit_geom[i] = 0u;
        else memcpy(L.out + i * L.out_stride + 32, &kNegInf, 4);
      }
      continue;
    }
    if (!fake_closest(pass, r.o, r.tnear, r.d, r.tfar, h)) continue;
    if (L.hit_geom) {
      L.hit_rec[i] = make_float4(h.t, h.u, h.v, 0.f);
      memcpy(&L.hit_rec[i].w, &h.primID, 4);
      L.hit_geom[i] = h.geomID;
      if (L.hit_ng) memcpy(L.hit_ng + 3 * i, h.ng, 12);
      if (L.hit_inst) L.hit_inst[i] = h.geomID + 100u;
      if (pass == 1) {  // handed to the next pass
        if (L.soa[0]) const_cast<float*>(L.soa[7])[i] = h.t;
        else memcpy(const_cast<uint8_t*>(L.in) + i * L.in_stride + L.tfar_offset, &h.t, 4);
      }
    } else {
      uint8_t* op = L.out + i * L.out_stride;
      memcpy(op + 32, &h.t, 4);
      memcpy(op + 48, h.ng, 12);
      memcpy(op + 60, &h.u, 4);
      memcpy(op + 64, &h.v, 4);
      memcpy(op + 68, &h.primID, 4);
      memcpy(op + 72, &h.geomID, 4);
      memcpy(op + 76, &L.instID, 4);
    }
  }
}

void any_pass(int pass, const EcbLaunch L) {
  for (uint64_t i = 0; i < L.M; i++) {
    const RayIn r = load_ray(L, i);
    const bool occ = fake_any(pass, r.o, r.tnear, r.d, r.tfar);
    if (L.hit_geom) {
      if (!L.keep_results) L.hit_geom[i] = occ ? 0u : 0xFFFFFFFFu;
      else if (occ) L.hit_geom[i] = 0u;
      if (occ && pass == 1) {
        if (L.soa[0]) const_cast<float*>(L.soa[7])[i] = kNegInf;
        else memcpy(const_cast<uint8_t*>(L.in) + i * L.in_stride + L.tfar_offset, &kNegInf, 4);
      }
    } else if (occ) {
      memcpy(L.out + i * L.out_stride + 32, &kNegInf, 4);
    }
  }
}

}  // namespace

cudaError_t ecb_launch_intersect(const EcbDeviceScene&, int, const EcbLaunch& L, cudaStream_t stream) {
  const EcbLaunch copy = L;
  mock_enqueue(stream, [copy] { closest_pass(0, copy); });
  return cudaSuccess;
}
cudaError_t ecb_launch_occluded(const EcbDeviceScene&, int, const EcbLaunch& L, cudaStream_t stream) {
  const EcbLaunch copy = L;
  mock_enqueue(stream, [copy] { any_pass(0, copy); });
  return cudaSuccess;
}
cudaError_t ecb_launch_triangles(const EcbDeviceTriScene&, int, const EcbLaunch& L, bool any_hit, cudaStream_t stream) {
  const EcbLaunch copy = L;
  mock_enqueue(stream, [copy, any_hit] { any_hit ? any_pass(1, copy) : closest_pass(1, copy); });
  return cudaSuccess;
}
cudaError_t ecb_launch_instanced(const EcbDeviceInstScene&, int, const EcbLaunch& L, bool any_hit, cudaStream_t stream) {
  const EcbLaunch copy = L;
  mock_enqueue(stream, [copy, any_hit] { any_hit ? any_pass(0, copy) : closest_pass(0, copy); });
  return cudaSuccess;
}
cudaError_t ecb_bounds_check_read(unsigned long long out[8], int* enabled) { *enabled = 0; for (int i = 0; i < 8; i++) out[i] = 0; return cudaSuccess; }
int ecb_kernel_mode() { return 0; }
void ecb_set_kernel_mode(int) {}
cudaError_t ecb_probe_read_bandwidth(size_t, int, int, double* g) { *g = 0.0; return cudaErrorNoDevice; }
size_t ecb_build_tmp_bytes(uint32_t) { return 0; }
cudaError_t ecb_launch_build(const EcbBuildLaunch&, int, cudaStream_t) { return cudaErrorNoDevice; }
