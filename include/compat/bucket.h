// compat/bucket.h -- drop-in for the reference's Bucket (reference src/bucket.h:19-36; behaviour of
// src/bucket.cpp:5-51 restated in visual_odom_b200/csrc/facade.cpp, including its quirks:
// features older than 9 frames are refused, a full bucket always overwrites its slot 0).
#ifndef BUCKET_H
#define BUCKET_H
#include "feature.h"

class Bucket {
public:
    explicit Bucket(int capacity);
    ~Bucket();

    int size();                                   // features currently held
    void add_feature(cv::Point2f p, int age);     // admit / overwrite rule of the reference
    void get_features(FeatureSet& out);           // appends the held features to `out`

    FeatureSet features;
    int max_size;
    int id;
};
#endif
