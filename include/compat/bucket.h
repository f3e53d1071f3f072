// compat/bucket.h -- the reference's Bucket (reference src/bucket.h:19-36, src/bucket.cpp:5-51).
#ifndef BUCKET_H
#define BUCKET_H
#include "feature.h"

class Bucket {
public:
    int id;
    int max_size;
    FeatureSet features;

    Bucket(int);
    ~Bucket();
    void add_feature(cv::Point2f, int);
    void get_features(FeatureSet&);
    int size();
};
#endif
