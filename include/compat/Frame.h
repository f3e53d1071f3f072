// compat/Frame.h -- drop-in for the reference's Frame holder (reference src/Frame.h:12-36,
// src/Frame.cpp:4-28): same members, same method names; triangulation runs on the GPU.
#ifndef FRAME_H
#define FRAME_H
#include "vo_cv.h"
#include <vector>

class Frame {
public:
    // projection matrices (3x4 CV_32F), world pose, stereo features of this frame
    cv::Mat m_projMatL, m_projMatR, m_worldRotation, m_worldTranslation;
    std::vector<cv::Point2f> m_pointsFeatureLeft, m_pointsFeatureRight;

    Frame();
    Frame(int frameId, const cv::Mat projMatL, const cv::Mat projMatR, cv::Mat worldRotation, cv::Mat worldTranslation);

    void setFeatures(std::vector<cv::Point2f> left, std::vector<cv::Point2f> right);
    // cv::triangulatePoints(m_projMatL, m_projMatR, left, right, points4D): points4D = 4 x N CV_32F, column i the
    // unit-norm homogeneous point of feature i -- the same bits OpenCV's per-point SVD produces (reference src/Frame.cpp:25-28).
    void triangulateFeaturePoints(cv::Mat& points4D);
};
#endif
