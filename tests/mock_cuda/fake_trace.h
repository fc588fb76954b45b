// tests/mock_cuda/fake_trace.h -- TEST INFRASTRUCTURE ONLY.  A deterministic stand-in for "tracing a ray": the result is a
// function of the ray's own fields only, so that the host pipeline test can predict what every record must hold after a
// call without any traversal.  Two passes, like a mixed scene (AccelN order): the triangle pass (pass 1) runs first and
// leaves its tfar in the record, the compressed-accel pass (pass 0) sees that tfar.
#pragma once
#include <cstdint>
#include <cstring>

struct FakeHit {
  float t, u, v, ng[3];
  uint32_t primID, geomID;
};
inline uint32_t fake_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline uint32_t fake_class(const float o[3], const float d[3]) { return (fake_bits(o[0]) ^ (fake_bits(d[1]) >> 3) ^ (fake_bits(o[2]) >> 7)) % 6u; }
inline bool fake_active(float tnear, float tfar) { return !(tnear > tfar); }

// closest hit of one pass.  classes: 0 miss; 1, 2 compressed only; 3 triangles only; 4 both, compressed nearer (replaces the
// triangle hit); 5 both, compressed farther (must lose against the tfar the triangle pass left behind)
inline bool fake_closest(int pass, const float o[3], float tnear, const float d[3], float tfar, FakeHit& h) {
  if (!fake_active(tnear, tfar)) return false;
  const uint32_t k = fake_class(o, d);
  float t;
  if (pass == 1) {
    if (k < 3) return false;
    t = 0.75f;
  } else {
    if (k == 0 || k == 3) return false;
    t = k == 5 ? 0.875f : 0.5f + 0.0625f * (float)(fake_bits(o[1]) & 3u);
  }
  if (!(t < tfar) || t < tnear) return false;
  h.t = t;
  h.u = d[0] * 0.5f;
  h.v = d[2] * 0.25f;
  h.primID = fake_bits(o[1]) & 0xFFFFu;
  h.geomID = pass == 1 ? 7u : (fake_bits(o[2]) & 3u);
  if (pass == 1) { h.ng[0] = d[2]; h.ng[1] = d[0]; h.ng[2] = d[1]; }
  else { h.ng[0] = 1.f; h.ng[1] = 0.f; h.ng[2] = 0.f; }
  return true;
}
// any hit of one pass (rays with tfar < 0 are "already occluded" and skipped, bvh_intersector1.cpp:133)
inline bool fake_any(int pass, const float o[3], float tnear, const float d[3], float tfar) {
  if (!fake_active(tnear, tfar) || tfar < 0.0f) return false;
  const uint32_t k = fake_class(o, d);
  return pass == 1 ? k >= 3 : (k == 1 || k == 2);
}
