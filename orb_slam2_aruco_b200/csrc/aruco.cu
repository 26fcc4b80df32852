// aruco.cu -- B200-native ArUco marker detector (sm_100a).  Placeholder translation unit: the entry points
// exist so that the C-ABI is complete; the kernels land in the next milestone.
#include "common.h"
using namespace b200;
struct b200_aruco_s { int device; };
extern "C" {
int b200_aruco_create(b200_aruco_t* out, const char*, int, int, int, int) { if (out) *out = nullptr; return fail(B200_EINVAL, "aruco detector %s", "not built yet"); }
int b200_aruco_destroy(b200_aruco_t) { return B200_OK; }
int b200_aruco_max_markers(b200_aruco_t) { return 64; }
int b200_aruco_detect(b200_aruco_t, const uint8_t*, int, int, int, int64_t, int64_t, b200_marker*, int32_t*, void*) { return fail(B200_EINVAL, "aruco detector %s", "not built yet"); }
int b200_aruco_detect_host(b200_aruco_t, const uint8_t*, int, int, int, int64_t, int64_t, b200_marker*, int32_t*) { return fail(B200_EINVAL, "aruco detector %s", "not built yet"); }
}
