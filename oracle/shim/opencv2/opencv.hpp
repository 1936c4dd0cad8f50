// oracle/shim/opencv2/opencv.hpp -- pinhole_camera_impl.h:21 includes OpenCV for PinholeCamera::FromFile (calibration
// file I/O, never instantiated on the hot path).  Declarations only, so that the template parses.  TEST INFRASTRUCTURE.
#ifndef DFK_SHIM_OPENCV_
#define DFK_SHIM_OPENCV_
#include <string>
namespace cv {
class Mat {
 public:
  template <typename T>
  T& at(int, int);
};
class FileNode {
 public:
  template <typename T>
  void operator>>(T&) const;
};
class FileStorage {
 public:
  enum { READ = 0 };
  FileStorage(const std::string&, int);
  bool isOpened() const;
  FileNode operator[](const char*) const;
};
}  // namespace cv
#endif
