// compat/utils.h -- the pose bookkeeping of the reference's utils.h (reference src/utils.h:36-41,
// src/utils.cpp:57-131): frame-to-frame motion -> accumulated camera pose, with the reference's gates.
// and the KITTI-layout image loaders (src/utils.cpp:172-190).  Host code; the
// trajectory display and the unused gyro loader are not part of this library.
#ifndef UTILS_H
#define UTILS_H
#include "vo_cv.h"
#include <string>

#ifndef VO_HAVE_OPENCV
namespace cv { struct Vec3f { float val[3]; float& operator[](int i) { return val[i]; } const float& operator[](int i) const { return val[i]; } }; }
#endif

// [R|t; 0 0 0 1] -> rigid_body_transformation = its inverse; frame_pose *= that inverse when
// 0.05 < |t| < 10 (otherwise the frame is skipped with a warning).               src/utils.cpp:57-91
void integrateOdometryStereo(int frame_id, cv::Mat& rigid_body_transformation, cv::Mat& frame_pose,
                             const cv::Mat& rotation, const cv::Mat& translation_stereo);
// |R^T R - I|_F < 1e-6                                                             src/utils.cpp:93-102
bool isRotationMatrix(cv::Mat& R);
// x-y-z Euler angles in float, as the reference computes them (sy in float)        src/utils.cpp:107-131
cv::Vec3f rotationMatrixToEulerAngles(cv::Mat& R);

// <filepath>image_0/%06d.png (left) / image_1/%06d.png (right) -> image_color (CV_8UC3, BGR, as imread(IMREAD_COLOR)) and
// image_gray (CV_8UC1, as cvtColor(BGR2GRAY)); decoded by the library's own PNG reader (row N3).  An unreadable file
// throws std::runtime_error (the reference would hand an empty Mat to cvtColor, which throws too).  src/utils.cpp:172-190
void loadImageLeft(cv::Mat& image_color, cv::Mat& image_gray, int frame_id, std::string filepath);
void loadImageRight(cv::Mat& image_color, cv::Mat& image_gray, int frame_id, std::string filepath);

#endif
