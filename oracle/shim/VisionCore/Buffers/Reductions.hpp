// oracle/shim/VisionCore/Buffers/Reductions.hpp -- host build: nothing (runReductions / finalizeReduction are device code)
#ifndef DFK_SHIM_VC_REDUCTIONS_
#define DFK_SHIM_VC_REDUCTIONS_
#include "../Platform.hpp"
#endif
