// written by oracle/build_ref.py (stands in for the CMake-configured kernels/config.h)
#define EMBREE_FILTER_FUNCTION
#define EMBREE_GEOMETRY_TRIANGLE
#define EMBREE_RAY_PACKETS

#define EMBREE_CURVE_SELF_INTERSECTION_AVOIDANCE_FACTOR 2.0
#define IF_ENABLED_TRIS(x) x
#define IF_ENABLED_QUADS(x)
#define IF_ENABLED_CURVES_OR_POINTS(x)
#define IF_ENABLED_CURVES(x)
#define IF_ENABLED_POINTS(x)
#define IF_ENABLED_SUBDIV(x)
#define IF_ENABLED_USER(x)
#define IF_ENABLED_INSTANCE(x)
#define IF_ENABLED_GRIDS(x)
