// compat/Frame.h -- the reference's Frame holder (reference src/Frame.h:12-36, src/Frame.cpp).
#ifndef FRAME_H
#define FRAME_H
#include "vo_cv.h"
#include <vector>

class Frame {
public:
    Frame();
    Frame(int frameId, const cv::Mat projMatL, const cv::Mat projMatR, cv::Mat worldRotation, cv::Mat worldTranslation);
    void setFeatures(std::vector<cv::Point2f> pointsFeatureLeft, std::vector<cv::Point2f> pointsFeatureRight);
    // cv::triangulatePoints(m_projMatL, m_projMatR, left, right, points4D): 4 x N CV_32F homogeneous points.
    // The GPU path returns the de-homogenised point (x, y, z, 1); see INTEGRATION.md.
    void triangulateFeaturePoints(cv::Mat& points4D);

    cv::Mat m_projMatL, m_projMatR;
    cv::Mat m_worldRotation, m_worldTranslation;
    std::vector<cv::Point2f> m_pointsFeatureLeft, m_pointsFeatureRight;
};
#endif
