// oracle/shim/VisionCore/CUDAGenerics.hpp -- host build: nothing (shuffles / SharedMemory are inside #ifdef __CUDACC__)
#ifndef DFK_SHIM_VC_CUDAGENERICS_
#define DFK_SHIM_VC_CUDAGENERICS_
#include "Platform.hpp"
#endif
