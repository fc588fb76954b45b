// shim_demo.cpp -- an "unmodified embree3 application" for the drop-in test: it only speaks the embree3 C API
// (declarations from include/embree3_b200/rtcore_subset.h, which restate the reference's headers) and is linked
// against libembree3_b200.so instead of libembree3.so.
//
//   shim_demo <leaf> <S> <C> <rays.bin> <out.bin> [instanced|plane|streams]
// builds the displaced cube of tutorials/displacement_geometry through rtc* calls (subdivision mesh + displacement
// callback), commits, and traces the RTCRayHit records of rays.bin four ways, writing one result block per way:
//   [0] rtcIntersect1M   [1] rtcIntersect16 packets   [2] rtcIntersect1 per ray (first 64 rays)   [3] rtcOccluded1M
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/embree3_b200/rtcore_subset.h"

static float cube_vertices[8][4] = {{-1, -1, -1, 0}, {1, -1, -1, 0}, {1, -1, 1, 0}, {-1, -1, 1, 0},
                                    {-1, 1, -1, 0},  {1, 1, -1, 0},  {1, 1, 1, 0},  {-1, 1, 1, 0}};
static unsigned cube_indices[24] = {0, 4, 5, 1, 1, 5, 6, 2, 2, 6, 7, 3, 0, 3, 7, 4, 4, 7, 6, 5, 0, 1, 2, 3};
static unsigned cube_faces[6] = {4, 4, 4, 4, 4, 4};

// same closed-form displacement as the reference harness uses for this scene (oracle/refhost.cpp)
static float displacement(float x, float y, float z) {
  float dN = 0.0f;
  for (float freq = 1.0f; freq < 40.0f; freq *= 2.0f) {
    const float n = std::sin(freq * (1.7f * x + 0.3f)) * std::sin(freq * (1.3f * y + 1.1f)) * std::sin(freq * (1.9f * z + 2.3f));
    dN += 1.4f * n * n / freq;
  }
  return dN;
}
static void displacement_cb(const RTCDisplacementFunctionNArguments* a) {
  for (unsigned i = 0; i < a->N; i++) {
    const float d = displacement(a->P_x[i], a->P_y[i], a->P_z[i]);
    a->P_x[i] += d * a->Ng_x[i];
    a->P_y[i] += d * a->Ng_y[i];
    a->P_z[i] += d * a->Ng_z[i];
  }
}
static void on_error(void*, RTCError code, const char* str) { fprintf(stderr, "RTC error %d: %s\n", (int)code, str); }

int main(int argc, char** argv) {
  if (argc < 6) { fprintf(stderr, "usage: shim_demo leaf S C rays.bin out.bin\n"); return 2; }
  const std::string cfg = std::string("subdiv_accel=bvh4.compressed.") + argv[1] + ",object_accel=bvh4.object,tri_accel=bvh4.triangle4v,set_affinity=0,start_threads=0";
  RTCDevice device = rtcNewDevice(cfg.c_str());
  if (!device) { fprintf(stderr, "rtcNewDevice failed (%d)\n", (int)rtcGetDeviceError(nullptr)); return 3; }
  rtcSetDeviceErrorFunction(device, on_error, nullptr);
  RTCScene scene = rtcNewScene(device);
  rtcSetSceneFlags(scene, RTC_SCENE_FLAG_ROBUST);
  rtcSetSceneLevels(scene, (unsigned)atoi(argv[2]), (unsigned)atoi(argv[3]));
  RTCGeometry geom = rtcNewGeometry(device, RTC_GEOMETRY_TYPE_SUBDIVISION);
  const bool plumbing = argc > 6 && std::string(argv[6]) == "plumbing";
  if (plumbing) {  // the same vertices through a buffer object (rtcore_buffer.h) instead of a shared pointer
    RTCBuffer vb = rtcNewSharedBuffer(device, (void*)cube_vertices, sizeof(cube_vertices));
    if (!vb || rtcGetBufferData(vb) != (void*)cube_vertices) { fprintf(stderr, "rtcNewSharedBuffer/rtcGetBufferData failed\n"); return 7; }
    rtcSetGeometryBuffer(geom, RTC_BUFFER_TYPE_VERTEX, 0, RTC_FORMAT_FLOAT3, vb, 0, 16, 8);
    rtcReleaseBuffer(vb);
  } else
  rtcSetSharedGeometryBuffer(geom, RTC_BUFFER_TYPE_VERTEX, 0, RTC_FORMAT_FLOAT3, cube_vertices, 0, 16, 8);
  rtcSetSharedGeometryBuffer(geom, RTC_BUFFER_TYPE_INDEX, 0, RTC_FORMAT_UINT, cube_indices, 0, 4, 24);
  rtcSetSharedGeometryBuffer(geom, RTC_BUFFER_TYPE_FACE, 0, RTC_FORMAT_UINT, cube_faces, 0, 4, 6);
  float* level = (float*)rtcSetNewGeometryBuffer(geom, RTC_BUFFER_TYPE_LEVEL, 0, RTC_FORMAT_FLOAT, 4, 24);
  for (int i = 0; i < 24; i++) level[i] = 1.0f;
  rtcSetGeometryDisplacementFunction(geom, displacement_cb);
  rtcCommitGeometry(geom);
  rtcAttachGeometry(scene, geom);
  rtcReleaseGeometry(geom);
  if (argc > 6 && std::string(argv[6]) == "plane") {
    // the tutorial's ground plane (tutorials/displacement_geometry/displacement_geometry_device.cpp:156-177): a triangle mesh in
    // the same scene -> the host builds a second accel and the shim hands both to the device path
    RTCGeometry plane = rtcNewGeometry(device, RTC_GEOMETRY_TYPE_TRIANGLE);
    float* v = (float*)rtcSetNewGeometryBuffer(plane, RTC_BUFFER_TYPE_VERTEX, 0, RTC_FORMAT_FLOAT3, 16, 4);
    const float vv[4][4] = {{-10, -2, -10, 0}, {-10, -2, 10, 0}, {10, -2, -10, 0}, {10, -2, 10, 0}};
    memcpy(v, vv, sizeof(vv));
    unsigned* t = (unsigned*)rtcSetNewGeometryBuffer(plane, RTC_BUFFER_TYPE_INDEX, 0, RTC_FORMAT_UINT3, 12, 2);
    const unsigned tt[6] = {0, 1, 2, 1, 3, 2};
    memcpy(t, tt, sizeof(tt));
    rtcCommitGeometry(plane);
    rtcAttachGeometry(scene, plane);
    rtcReleaseGeometry(plane);
  }
  rtcCommitScene(scene);
  if (RTCError e = rtcGetDeviceError(device)) { fprintf(stderr, "commit failed: %d\n", (int)e); return 4; }
  RTCScene object_scene = nullptr;
  if (argc > 6 && std::string(argv[6]) == "instanced") {
    // two instances of the cube scene (tutorials/instanced_geometry style): the object scene stays alive, the queries go to the top scene
    object_scene = scene;
    scene = rtcNewScene(device);
    const float xfm[2][12] = {{1, 0, 0, 0, 1, 0, 0, 0, 1, -1.6f, 0, 0}, {0.5f, 0, 0, 0, 0.5f, 0, 0, 0, 0.5f, 1.4f, 0.3f, 0}};
    for (int i = 0; i < 2; i++) {
      RTCGeometry inst = rtcNewGeometry(device, RTC_GEOMETRY_TYPE_INSTANCE);
      rtcSetGeometryInstancedScene(inst, object_scene);
      rtcSetGeometryTransform(inst, 0, RTC_FORMAT_FLOAT3X4_COLUMN_MAJOR, xfm[i]);
      rtcCommitGeometry(inst);
      rtcAttachGeometry(scene, inst);
      rtcReleaseGeometry(inst);
    }
    rtcCommitScene(scene);
    if (RTCError e = rtcGetDeviceError(device)) { fprintf(stderr, "instanced commit failed: %d\n", (int)e); return 4; }
  }

  FILE* f = fopen(argv[4], "rb");
  if (!f) return 5;
  fseek(f, 0, SEEK_END);
  const size_t bytes = (size_t)ftell(f);
  fseek(f, 0, SEEK_SET);
  const size_t n = bytes / sizeof(RTCRayHit);
  std::vector<RTCRayHit> rays(n);
  if (fread(rays.data(), sizeof(RTCRayHit), n, f) != n) return 5;
  fclose(f);
  FILE* out = fopen(argv[5], "wb");
  RTCIntersectContext ctx;
  rtcInitIntersectContext(&ctx);

  {  // [0] stream call on pinned memory
    RTCRayHit* r = (RTCRayHit*)rtcAllocRays_b200(bytes);
    memcpy(r, rays.data(), bytes);
    rtcIntersect1M(scene, &ctx, r, (unsigned)n, sizeof(RTCRayHit));
    fwrite(r, sizeof(RTCRayHit), n, out);
    rtcFreeRays_b200(r);
  }
  {  // [1] packets of 16 with every 5th lane masked out
    std::vector<RTCRayHit> r(rays);
    const size_t packets = n / 16 < 32 ? n / 16 : 32;
    for (size_t p = 0; p < packets; p++) {
      RTCRayHit16 pk;
      int valid[16];
      unsigned* soa = (unsigned*)&pk;
      for (int l = 0; l < 16; l++) {
        valid[l] = (l % 5 == 4) ? 0 : -1;
        const unsigned* a = (const unsigned*)&r[p * 16 + l];
        for (int fld = 0; fld < 20; fld++) soa[fld * 16 + l] = a[fld];
      }
      rtcIntersect16(valid, scene, &ctx, &pk);
      for (int l = 0; l < 16; l++) {
        unsigned* a = (unsigned*)&r[p * 16 + l];
        for (int fld = 0; fld < 20; fld++) a[fld] = soa[fld * 16 + l];
      }
    }
    fwrite(r.data(), sizeof(RTCRayHit), packets * 16, out);
  }
  {  // [2] single-ray calls (pageable stack memory)
    const size_t m = n < 64 ? n : 64;
    std::vector<RTCRayHit> r(rays.begin(), rays.begin() + m);
    for (size_t i = 0; i < m; i++) rtcIntersect1(scene, &ctx, &r[i]);
    fwrite(r.data(), sizeof(RTCRayHit), m, out);
  }
  {  // [3] any-hit stream on pageable memory with a 64-byte stride
    struct Padded { RTCRay ray; float pad[4]; };
    std::vector<Padded> r(n);
    for (size_t i = 0; i < n; i++) { r[i].ray = rays[i].ray; for (int k = 0; k < 4; k++) r[i].pad[k] = 7.f; }
    rtcOccluded1M(scene, &ctx, &r[0].ray, (unsigned)n, sizeof(Padded));
    for (size_t i = 0; i < n; i++) fwrite(&r[i], sizeof(Padded), 1, out);
  }
  if (argc > 6 && std::string(argv[6]) == "streams") {
    // the other stream layouts of the API: each must give what rtcIntersect1M gave
    {  // [4] rtcIntersect1Mp: array of pointers, here in reverse order
      std::vector<RTCRayHit> r(rays);
      std::vector<RTCRayHit*> ptr(n);
      for (size_t i = 0; i < n; i++) ptr[i] = &r[n - 1 - i];
      rtcIntersect1Mp(scene, &ctx, ptr.data(), (unsigned)n);
      fwrite(r.data(), sizeof(RTCRayHit), n, out);
    }
    {  // [5] rtcIntersectNM: packets of N = 8 rays in SoA layout, 64 bytes of padding between packets
      const unsigned N = 8;
      const size_t M = n / N, stride = 20 * N * 4 + 64;
      std::vector<unsigned char> buf(M * stride, 0x5a);
      for (size_t i = 0; i < M * N; i++)
        for (int fld = 0; fld < 20; fld++)
          ((unsigned*)(buf.data() + (i / N) * stride))[fld * N + i % N] = ((const unsigned*)&rays[i])[fld];
      rtcIntersectNM(scene, &ctx, (RTCRayHitN*)buf.data(), N, (unsigned)M, stride);
      std::vector<RTCRayHit> r(M * N);
      for (size_t i = 0; i < M * N; i++)
        for (int fld = 0; fld < 20; fld++)
          ((unsigned*)&r[i])[fld] = ((const unsigned*)(buf.data() + (i / N) * stride))[fld * N + i % N];
      fwrite(r.data(), sizeof(RTCRayHit), M * N, out);
      for (size_t p = 0; p < M; p++)
        for (size_t b = 20 * N * 4; b < stride; b++)
          if (buf[p * stride + b] != 0x5a) { fprintf(stderr, "rtcIntersectNM touched the padding\n"); return 7; }
    }
    {  // [6] rtcIntersectNp: one array per field
      std::vector<std::vector<unsigned>> f(20, std::vector<unsigned>(n));
      for (size_t i = 0; i < n; i++)
        for (int fld = 0; fld < 20; fld++) f[fld][i] = ((const unsigned*)&rays[i])[fld];
      RTCRayHitNp np;
      np.ray.org_x = (float*)f[0].data(); np.ray.org_y = (float*)f[1].data(); np.ray.org_z = (float*)f[2].data();
      np.ray.tnear = (float*)f[3].data(); np.ray.dir_x = (float*)f[4].data(); np.ray.dir_y = (float*)f[5].data();
      np.ray.dir_z = (float*)f[6].data(); np.ray.time = (float*)f[7].data(); np.ray.tfar = (float*)f[8].data();
      np.ray.mask = f[9].data(); np.ray.id = f[10].data(); np.ray.flags = f[11].data();
      np.hit.Ng_x = (float*)f[12].data(); np.hit.Ng_y = (float*)f[13].data(); np.hit.Ng_z = (float*)f[14].data();
      np.hit.u = (float*)f[15].data(); np.hit.v = (float*)f[16].data();
      np.hit.primID = f[17].data(); np.hit.geomID = f[18].data(); np.hit.instID[0] = f[19].data();
      rtcIntersectNp(scene, &ctx, &np, (unsigned)n);
      std::vector<RTCRayHit> r(n);
      for (size_t i = 0; i < n; i++)
        for (int fld = 0; fld < 20; fld++) ((unsigned*)&r[i])[fld] = f[fld][i];
      fwrite(r.data(), sizeof(RTCRayHit), n, out);
      // [7] rtcOccludedNp on the same arrays restored to the input rays: tfar only
      for (size_t i = 0; i < n; i++)
        for (int fld = 0; fld < 12; fld++) f[fld][i] = ((const unsigned*)&rays[i])[fld];
      rtcOccludedNp(scene, &ctx, &np.ray, (unsigned)n);
      fwrite(f[8].data(), 4, n, out);
      // [8] rtcOccluded1Mp / rtcOccludedNM (N = 4): tfar only
      std::vector<RTCRay> o(n);
      std::vector<RTCRay*> optr(n);
      for (size_t i = 0; i < n; i++) { o[i] = rays[i].ray; optr[i] = &o[i]; }
      rtcOccluded1Mp(scene, &ctx, optr.data(), (unsigned)n);
      for (size_t i = 0; i < n; i++) fwrite(&o[i].tfar, 4, 1, out);
      const unsigned N4 = 4;
      std::vector<unsigned> soa(12 * n);
      for (size_t i = 0; i < n / N4 * N4; i++)
        for (int fld = 0; fld < 12; fld++) soa[(i / N4) * 12 * N4 + fld * N4 + i % N4] = ((const unsigned*)&rays[i])[fld];
      rtcOccludedNM(scene, &ctx, (RTCRayN*)soa.data(), N4, (unsigned)(n / N4), 12 * N4 * 4);
      for (size_t i = 0; i < n / N4 * N4; i++) fwrite(&soa[(i / N4) * 12 * N4 + 8 * N4 + i % N4], 4, 1, out);
    }
  }
  if (plumbing) {
    // what the device path answers for itself, what it passes on to the host, and what it refuses
    if (rtcGetDeviceProperty(device, RTC_DEVICE_PROPERTY_RAY_STREAM_SUPPORTED) != 1 ||
        rtcGetDeviceProperty(device, RTC_DEVICE_PROPERTY_NATIVE_RAY16_SUPPORTED) != 1 ||
        rtcGetDeviceProperty(device, RTC_DEVICE_PROPERTY_FILTER_FUNCTION_SUPPORTED) != 0 ||
        rtcGetDeviceProperty(device, RTC_DEVICE_PROPERTY_SUBDIVISION_GEOMETRY_SUPPORTED) != 1 ||
        rtcGetDeviceProperty(device, RTC_DEVICE_PROPERTY_VERSION_MAJOR) != 3) { fprintf(stderr, "rtcGetDeviceProperty: unexpected answers\n"); return 8; }
    if (rtcGetDeviceError(device) != RTC_ERROR_NONE) return 9;
    RTCRayHit probe = rays[n / 2 + 80], before = probe;   // a ray that hits (the frame centre)
    RTCIntersectContext fctx;
    rtcInitIntersectContext(&fctx);
    fctx.filter = (RTCFilterFunctionN)on_error;            // any non-null callback: must be refused, ray untouched
    rtcIntersect1(scene, &fctx, &probe);
    if (memcmp(&probe, &before, sizeof(probe)) != 0) { fprintf(stderr, "a query with a filter callback modified the ray\n"); return 10; }
    if (rtcGetDeviceError(device) != RTC_ERROR_INVALID_OPERATION) { fprintf(stderr, "filter callback in the context was not refused\n"); return 11; }
    rtcIntersect1(scene, &ctx, &probe);
    if (probe.hit.geomID == RTC_INVALID_GEOMETRY_ID) { fprintf(stderr, "probe ray missed\n"); return 12; }
    RTCGeometry g0 = rtcGetGeometry(scene, 0);
    rtcSetGeometryIntersectFilterFunction(g0, (RTCFilterFunctionN)on_error);
    if (rtcGetDeviceError(device) != RTC_ERROR_INVALID_OPERATION) { fprintf(stderr, "geometry filter callback was not refused\n"); return 13; }
    rtcSetGeometryIntersectFilterFunction(g0, nullptr);    // clearing one is fine
    rtcSetGeometryMask(g0, 0xFFFFFFFFu);
    if (rtcGetDeviceError(device) != RTC_ERROR_NONE) return 14;
    RTCBVH bvh = rtcNewBVH(device);                        // the stand-alone builder API lives on the host
    if (!bvh) return 15;
    rtcReleaseBVH(bvh);
    printf("plumbing ok\n");
  }
  fclose(out);
  const RTCError e = rtcGetDeviceError(device);
  rtcReleaseScene(scene);
  if (object_scene) rtcReleaseScene(object_scene);
  rtcReleaseDevice(device);
  if (e) { fprintf(stderr, "device error %d\n", (int)e); return 6; }
  printf("shim_demo ok: %zu rays\n", n);
  return 0;
}
