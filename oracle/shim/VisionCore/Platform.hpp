// oracle/shim/VisionCore/Platform.hpp -- stand-in for VisionCore (jczarnowski/vision_core @ 924c5333, not in this image).
// TEST INFRASTRUCTURE ONLY (see Eigen/Core in this directory): the subset the reference's hot-path headers touch on
// the HOST.  Target tags only.
#ifndef DFK_SHIM_VC_PLATFORM_
#define DFK_SHIM_VC_PLATFORM_
#include <cstddef>
namespace vc {
struct TargetHost {};
struct TargetDeviceCUDA {};
}  // namespace vc
#endif
