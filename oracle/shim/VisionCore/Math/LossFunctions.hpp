// oracle/shim/VisionCore/Math/LossFunctions.hpp -- included by dense_sfm.h:24, nothing of it is used on the path
#ifndef DFK_SHIM_VC_LOSSFUNCTIONS_
#define DFK_SHIM_VC_LOSSFUNCTIONS_
#endif
