// cvshim.h -- the handful of OpenCV types the reference's feature.h / visualOdometry.h / Frame.h
// signatures mention, for builds WITHOUT OpenCV (this container has no OpenCV C++ headers).
// With real OpenCV present the facade compiles against it instead: see compat/vo_cv.h.
// Only what the drop-in needs: POD points, Size, and a ref-counted dense cv::Mat (header copy on
// pass-by-value, pixels never copied -- the semantics circularMatching(cv::Mat ...) relies on).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <vector>

typedef unsigned char uchar;      // OpenCV defines it at global scope too

#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

namespace cv {

using ::uchar;

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
typedef Point_<int> Point;
template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point3_<float> Point3f;
struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};

class Mat {
public:
    int rows, cols;
    size_t step;      // bytes per row
    uchar* data;

    Mat() : rows(0), cols(0), step(0), data(nullptr), type_(0) {}
    Mat(int r, int c, int type) { create(r, c, type); }
    // wraps caller memory (no ownership), like cv::Mat(rows, cols, type, data, step)
    Mat(int r, int c, int type, void* d, size_t s = 0)
        : rows(r), cols(c), step(s ? s : (size_t)c * elem_size(type)), data((uchar*)d), type_(type) {}

    void create(int r, int c, int type)
    {
        rows = r; cols = c; type_ = type; step = (size_t)c * elem_size(type);
        store_.reset(new uchar[step * (size_t)(r > 0 ? r : 1)], std::default_delete<uchar[]>());
        data = store_.get();
        std::memset(data, 0, step * (size_t)r);
    }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    static Mat eye(int r, int c, int type)
    {
        Mat m(r, c, type);
        for (int i = 0; i < r && i < c; i++) {
            if (depth_of(type) == CV_64F) m.at<double>(i, i) = 1.0;
            else if (depth_of(type) == CV_32F) m.at<float>(i, i) = 1.f;
            else throw std::runtime_error("cvshim: eye() depth");
        }
        return m;
    }
    int type() const { return type_; }
    int depth() const { return depth_of(type_); }
    int channels() const { return (type_ >> 3) + 1; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t elemSize() const { return elem_size(type_); }
    Mat clone() const
    {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; r++) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * elemSize());
        return m;
    }
    template <typename T> T& at(int r, int c = 0) { return *reinterpret_cast<T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> const T& at(int r, int c = 0) const { return *reinterpret_cast<const T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }

private:
    static int depth_of(int type) { return type & 7; }
    static size_t elem_size(int type)
    {
        static const size_t dsz[8] = {1, 1, 2, 2, 4, 4, 8, 2};
        return dsz[type & 7] * (size_t)((type >> 3) + 1);
    }
    int type_;
    std::shared_ptr<uchar> store_;
};

} // namespace cv
