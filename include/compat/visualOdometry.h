// compat/visualOdometry.h -- the reference's visualOdometry.h interface (reference
// src/visualOdometry.h:27-46) over the B200 C-ABI.
#ifndef VISUAL_ODOM_H
#define VISUAL_ODOM_H

#include "feature.h"
#include "bucket.h"
#include "Frame.h"
#include <vector>

// reference src/visualOdometry.cpp:81-129
void matchingFeatures(cv::Mat& imageLeft_t0, cv::Mat& imageRight_t0,
                      cv::Mat& imageLeft_t1, cv::Mat& imageRight_t1,
                      FeatureSet& currentVOFeatures,
                      std::vector<cv::Point2f>& pointsLeft_t0,
                      std::vector<cv::Point2f>& pointsRight_t0,
                      std::vector<cv::Point2f>& pointsLeft_t1,
                      std::vector<cv::Point2f>& pointsRight_t1);

// reference src/visualOdometry.cpp:132-193.  mono_rotation=true (5-point essential matrix, not
// executed by the reference's main(), src/main.cpp:181) throws std::runtime_error.
void trackingFrame2Frame(cv::Mat& projMatrl, cv::Mat& projMatrr,
                         std::vector<cv::Point2f>& pointsLeft_t0,
                         std::vector<cv::Point2f>& pointsLeft_t1,
                         cv::Mat& points3D_t0,
                         cv::Mat& rotation,
                         cv::Mat& translation,
                         bool mono_rotation = true);

// reference src/visualOdometry.cpp:195-224 (imshow visualisation): a no-op here.
void displayTracking(cv::Mat& imageLeft_t1, std::vector<cv::Point2f>& pointsLeft_t0, std::vector<cv::Point2f>& pointsLeft_t1);

// --- additions of this library (not in the reference) -------------------------------------------
// The call site reference src/main.cpp:170-171 (triangulatePoints + convertPointsFromHomogeneous):
// fills points3D (N x 1 CV_32FC3).
void triangulateStereo(cv::Mat& projMatrl, cv::Mat& projMatrr, std::vector<cv::Point2f>& pointsLeft,
                       std::vector<cv::Point2f>& pointsRight, cv::Mat& points3D);
// Inlier indices of the last trackingFrame2Frame call (the reference only prints their count).
const std::vector<int>& lastPnPInliers();
// Selects the CUDA device of the process-wide context the facade creates lazily (default 0).
void voCompatSetDevice(int device);

#endif
