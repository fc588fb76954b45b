// Hand-written equivalent of the file CMake would generate from the reference's kernels/config.h.in
// with the reference's default options (CMakeLists.txt:96-115). Used only by oracle/build_ref.py.
#define EMBREE_FILTER_FUNCTION
#define EMBREE_GEOMETRY_TRIANGLE
#define EMBREE_GEOMETRY_QUAD
#define EMBREE_GEOMETRY_CURVE
#define EMBREE_GEOMETRY_SUBDIVISION
#define EMBREE_GEOMETRY_USER
#define EMBREE_RAY_PACKETS

#define IF_ENABLED_TRIS(x) x
#define IF_ENABLED_QUADS(x) x
#define IF_ENABLED_CURVES(x) x
#define IF_ENABLED_SUBDIV(x) x
#define IF_ENABLED_USER(x) x
