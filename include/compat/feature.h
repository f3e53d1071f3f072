// compat/feature.h -- the reference's feature.h interface (reference src/feature.h:27-79), same
// names, argument order and in-place vector semantics, implemented over the B200 C-ABI
// (include/vo_b200.h).  A translation unit of the reference that includes "feature.h" and links
// libvo_facade.so instead of libfeature.so / libbucket.so keeps compiling and behaving the same.
#ifndef FEATURE_H
#define FEATURE_H

#include "vo_cv.h"
#include <vector>

struct FeaturePoint {          // reference src/feature.h:27-31 (unused there too)
    cv::Point2f point;
    int id;
    int age;
};

struct FeatureSet {            // reference src/feature.h:33-43
    std::vector<cv::Point2f> points;
    std::vector<int> ages;
    int size() { return (int)points.size(); }
    void clear() { points.clear(); ages.clear(); }
};

// reference src/feature.cpp:20-37
void deleteUnmatchFeatures(std::vector<cv::Point2f>& points0, std::vector<cv::Point2f>& points1, std::vector<uchar>& status);
// reference src/feature.cpp:39-47  -> vo_fast_detect
void featureDetectionFast(cv::Mat image, std::vector<cv::Point2f>& points);
// reference src/feature.cpp:49-62  (never called by the reference's main loop; not built on the GPU)
void featureDetectionGoodFeaturesToTrack(cv::Mat image, std::vector<cv::Point2f>& points);
// reference src/feature.cpp:64-74  -> vo_lk_track
void featureTracking(cv::Mat img_1, cv::Mat img_2, std::vector<cv::Point2f>& points1, std::vector<cv::Point2f>& points2, std::vector<uchar>& status);
// reference src/feature.cpp:76-116
void deleteUnmatchFeaturesCircle(std::vector<cv::Point2f>& points0, std::vector<cv::Point2f>& points1,
                                 std::vector<cv::Point2f>& points2, std::vector<cv::Point2f>& points3,
                                 std::vector<cv::Point2f>& points0_return,
                                 std::vector<uchar>& status0, std::vector<uchar>& status1,
                                 std::vector<uchar>& status2, std::vector<uchar>& status3,
                                 std::vector<int>& ages);
// reference src/feature.cpp:118-148 -> vo_circular_match
void circularMatching(cv::Mat img_l_0, cv::Mat img_r_0, cv::Mat img_l_1, cv::Mat img_r_1,
                      std::vector<cv::Point2f>& points_l_0, std::vector<cv::Point2f>& points_r_0,
                      std::vector<cv::Point2f>& points_l_1, std::vector<cv::Point2f>& points_r_1,
                      std::vector<cv::Point2f>& points_l_0_return,
                      FeatureSet& current_features);
// The reference's USE_CUDA variant (src/feature.cpp:150-204) has the same contract; here it is the same call.
inline void circularMatching_gpu(cv::Mat img_l_0, cv::Mat img_r_0, cv::Mat img_l_1, cv::Mat img_r_1,
                                 std::vector<cv::Point2f>& points_l_0, std::vector<cv::Point2f>& points_r_0,
                                 std::vector<cv::Point2f>& points_l_1, std::vector<cv::Point2f>& points_r_1,
                                 std::vector<cv::Point2f>& points_l_0_return, FeatureSet& current_features)
{
    circularMatching(img_l_0, img_r_0, img_l_1, img_r_1, points_l_0, points_r_0, points_l_1, points_r_1, points_l_0_return, current_features);
}
// reference src/feature.cpp:206-253
void bucketingFeatures(cv::Mat& image, FeatureSet& current_features, int bucket_size, int features_per_bucket);
// reference src/feature.cpp:255-269
void appendNewFeatures(cv::Mat& image, FeatureSet& current_features);
void appendNewFeatures(std::vector<cv::Point2f> points_new, FeatureSet& current_features);

#endif
