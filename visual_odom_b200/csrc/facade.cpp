// facade.cpp -- the reference's C++ entry points (include/compat/feature.h, bucket.h,
// visualOdometry.h, Frame.h) implemented over the C-ABI of libvo_b200.so.
//
// What runs where:
//   GPU (through include/vo_b200.h): FAST detection, the four chained LK calls, triangulation,
//        PnP/RANSAC + LM + Rodrigues -- every OpenCV call of the reference's hot path.
//   host (here): the reference's own O(N) vector glue, restated with its quirks (SURVEY.md
//        Appendix A): status / negative-coordinate erase loops, the age counter, the bucket grid
//        with its index-stride aliasing, the integer-truncated circular check.
// No OpenCV is needed; with OpenCV headers present the same file compiles against real cv::Mat.
#include "../../include/compat/visualOdometry.h"
#include "../../include/vo_b200.h"

#include <cmath>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace {

int g_device = 0;
vo_ctx* g_ctx = nullptr;
std::vector<int> g_last_inliers;

vo_ctx* context()
{
    if (!g_ctx) {
        vo_params p;
        vo_default_params(&p);
        p.max_features = 16384;
        vo_ctx* c = nullptr;
        const int rc = vo_create(g_device, &p, &c);
        if (rc != VO_OK) {
            std::string msg = c ? vo_last_error(c) : "vo_create failed";
            if (c) vo_destroy(c);
            throw std::runtime_error("vo_b200: " + msg);
        }
        g_ctx = c;
    }
    return g_ctx;
}

void check(int rc, const char* what)
{
    if (rc != VO_OK) throw std::runtime_error(std::string("vo_b200 ") + what + ": " + vo_last_error(g_ctx));
}

void require_gray8(const cv::Mat& m, const char* what)
{
    if (m.empty() || m.type() != CV_8UC1) throw std::runtime_error(std::string(what) + ": expected a non-empty CV_8UC1 image");
}

inline const vo_point2f* as_vo(const std::vector<cv::Point2f>& v) { return reinterpret_cast<const vo_point2f*>(v.data()); }
inline vo_point2f* as_vo(std::vector<cv::Point2f>& v) { return reinterpret_cast<vo_point2f*>(v.data()); }

// keep[i] != 0 -> element i survives; order preserved (what the reference's erase loops compute)
template <typename T> void compact(std::vector<T>& v, const std::vector<uchar>& keep, size_t n)
{
    size_t o = 0;
    for (size_t i = 0; i < n && i < v.size(); i++)
        if (keep[i]) v[o++] = v[i];
    // entries beyond n (possible for `ages`, Appendix A item 8) keep their relative position
    for (size_t i = n; i < v.size(); i++) v[o++] = v[i];
    v.resize(o);
}

// visualOdometry.cpp:44-61 -- `int offset` truncates the float maximum
void checkValidMatch(const std::vector<cv::Point2f>& points, const std::vector<cv::Point2f>& points_return,
                     std::vector<uchar>& status, int threshold)
{
    status.assign(points.size(), 1);
    for (size_t i = 0; i < points.size(); i++) {
        const float m = std::max(std::abs(points[i].x - points_return[i].x), std::abs(points[i].y - points_return[i].y));
        const int offset = (int)m;
        if (offset > threshold) status[i] = 0;
    }
}

} // namespace

void voCompatSetDevice(int device) { g_device = device; }
const std::vector<int>& lastPnPInliers() { return g_last_inliers; }

// ------------------------------------------------------------------------------------------------ feature.h
void deleteUnmatchFeatures(std::vector<cv::Point2f>& points0, std::vector<cv::Point2f>& points1, std::vector<uchar>& status)
{
    const size_t n = status.size();
    std::vector<uchar> keep(n);
    for (size_t i = 0; i < n; i++) {
        const cv::Point2f pt = points1.at(i);
        const bool neg = (pt.x < 0) || (pt.y < 0);
        if (neg) status[i] = 0;
        keep[i] = (status[i] != 0);
    }
    compact(points0, keep, n);
    compact(points1, keep, n);
}

void featureDetectionFast(cv::Mat image, std::vector<cv::Point2f>& points)
{
    require_gray8(image, "featureDetectionFast");
    vo_ctx* c = context();
    int cap = 1 << 16, n = 0;
    for (;;) {
        points.resize((size_t)cap);
        const int rc = vo_fast_detect(c, image.data, image.cols, image.rows, image.step, as_vo(points), nullptr, cap, &n);
        if (rc == VO_E_CAPACITY && n > cap && cap < (1 << 24)) { cap = n; continue; }
        if (rc != VO_E_CAPACITY) check(rc, "vo_fast_detect");
        break;
    }
    points.resize((size_t)std::min(n, cap));
}

void featureDetectionGoodFeaturesToTrack(cv::Mat, std::vector<cv::Point2f>&)
{
    throw std::runtime_error("featureDetectionGoodFeaturesToTrack: not part of the hot path (the reference never calls it); not built");
}

void featureTracking(cv::Mat img_1, cv::Mat img_2, std::vector<cv::Point2f>& points1, std::vector<cv::Point2f>& points2, std::vector<uchar>& status)
{
    require_gray8(img_1, "featureTracking"); require_gray8(img_2, "featureTracking");
    const int n = (int)points1.size();
    points2.resize((size_t)n); status.resize((size_t)n);
    std::vector<float> err((size_t)n);
    if (n) check(vo_lk_track(context(), img_1.data, img_2.data, img_1.cols, img_1.rows, img_1.step, as_vo(points1), n,
                             as_vo(points2), status.data(), err.data()), "vo_lk_track");
    deleteUnmatchFeatures(points1, points2, status);
}

void deleteUnmatchFeaturesCircle(std::vector<cv::Point2f>& points0, std::vector<cv::Point2f>& points1,
                                 std::vector<cv::Point2f>& points2, std::vector<cv::Point2f>& points3,
                                 std::vector<cv::Point2f>& points0_return,
                                 std::vector<uchar>& status0, std::vector<uchar>& status1,
                                 std::vector<uchar>& status2, std::vector<uchar>& status3,
                                 std::vector<int>& ages)
{
    for (size_t i = 0; i < ages.size(); i++) ages[i] += 1;          // every feature ages, survivors or not
    const size_t n = status3.size();
    std::vector<uchar> keep(n);
    for (size_t i = 0; i < n; i++) {
        const cv::Point2f &p0 = points0.at(i), &p1 = points1.at(i), &p2 = points2.at(i), &p3 = points3.at(i);
        const bool neg = (p0.x < 0) || (p0.y < 0) || (p1.x < 0) || (p1.y < 0) || (p2.x < 0) || (p2.y < 0) || (p3.x < 0) || (p3.y < 0);
        const bool lost = (status3[i] == 0) || (status2.at(i) == 0) || (status1.at(i) == 0) || (status0.at(i) == 0);
        if (neg) status3[i] = 0;                                    // points0_return is not part of the test
        keep[i] = !(neg || lost);
    }
    compact(points0, keep, n); compact(points1, keep, n); compact(points2, keep, n); compact(points3, keep, n);
    compact(points0_return, keep, n);
    compact(ages, keep, n);
}

void circularMatching(cv::Mat img_l_0, cv::Mat img_r_0, cv::Mat img_l_1, cv::Mat img_r_1,
                      std::vector<cv::Point2f>& points_l_0, std::vector<cv::Point2f>& points_r_0,
                      std::vector<cv::Point2f>& points_l_1, std::vector<cv::Point2f>& points_r_1,
                      std::vector<cv::Point2f>& points_l_0_return,
                      FeatureSet& current_features)
{
    require_gray8(img_l_0, "circularMatching"); require_gray8(img_r_0, "circularMatching");
    require_gray8(img_l_1, "circularMatching"); require_gray8(img_r_1, "circularMatching");
    const int n = (int)points_l_0.size();
    std::vector<uchar> st((size_t)4 * n);
    std::vector<cv::Point2f> raw((size_t)4 * n);
    if (n) {
        int kept = 0;
        // one launch chain on the GPU: L0->R0, R0->R1, R1->L1, L1->L0 (raw outputs, original indexing)
        check(vo_circular_match(context(), img_l_0.data, img_r_0.data, img_l_1.data, img_r_1.data, img_l_0.cols, img_l_0.rows,
                                img_l_0.step, as_vo(points_l_0), n, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                st.data(), as_vo(raw), nullptr, &kept), "vo_circular_match");
    }
    points_r_0.assign(raw.begin(), raw.begin() + n);
    points_r_1.assign(raw.begin() + n, raw.begin() + 2 * n);
    points_l_1.assign(raw.begin() + 2 * n, raw.begin() + 3 * n);
    points_l_0_return.assign(raw.begin() + 3 * n, raw.begin() + 4 * n);
    std::vector<uchar> s0(st.begin(), st.begin() + n), s1(st.begin() + n, st.begin() + 2 * n),
        s2(st.begin() + 2 * n, st.begin() + 3 * n), s3(st.begin() + 3 * n, st.begin() + 4 * n);
    deleteUnmatchFeaturesCircle(points_l_0, points_r_0, points_r_1, points_l_1, points_l_0_return, s0, s1, s2, s3,
                                current_features.ages);
}

// ------------------------------------------------------------------------------------------------ bucket.h
Bucket::Bucket(int capacity) : max_size(capacity), id(0) {}
Bucket::~Bucket() {}
int Bucket::size() { return (int)features.points.size(); }

void Bucket::add_feature(cv::Point2f point, int age)
{
    if (age >= 10) return;                      // features older than 10 frames are never re-admitted
    if (size() < max_size) {
        features.points.push_back(point);
        features.ages.push_back(age);
        return;
    }
    // The reference's "replace the youngest" scan compares the incoming age with itself, so its
    // chosen slot is always 0: a full bucket keeps the LAST admitted feature in slot 0.
    features.points[0] = point;
    features.ages[0] = age;
}

void Bucket::get_features(FeatureSet& out)
{
    out.points.insert(out.points.end(), features.points.begin(), features.points.end());
    out.ages.insert(out.ages.end(), features.ages.begin(), features.ages.end());
}

void bucketingFeatures(cv::Mat& image, FeatureSet& current_features, int bucket_size, int features_per_bucket)
{
    const int nh = image.rows / bucket_size, nw = image.cols / bucket_size;
    // (nh+1) x (nw+1) buckets are allocated but addressed with row stride nw (not nw+1): cell
    // (h, nw) aliases cell (h+1, 0) and the read-back visits nh aliased cells twice.
    std::vector<Bucket> cells((size_t)(nh + 1) * (nw + 1), Bucket(features_per_bucket));
    for (size_t i = 0; i < current_features.points.size(); i++) {
        const int bh = (int)(current_features.points[i].y / bucket_size);
        const int bw = (int)(current_features.points[i].x / bucket_size);
        const long idx = (long)bh * nw + bw;
        if (idx < 0 || idx >= (long)cells.size())
            throw std::out_of_range("bucketingFeatures: feature outside the image (undefined behaviour in the reference)");
        cells[(size_t)idx].add_feature(current_features.points[i], current_features.ages[i]);
    }
    current_features.clear();
    for (int h = 0; h <= nh; h++)
        for (int w = 0; w <= nw; w++) cells[(size_t)h * nw + w].get_features(current_features);
}

void appendNewFeatures(cv::Mat& image, FeatureSet& current_features)
{
    std::vector<cv::Point2f> fresh;
    featureDetectionFast(image, fresh);
    appendNewFeatures(fresh, current_features);
}

void appendNewFeatures(std::vector<cv::Point2f> points_new, FeatureSet& current_features)
{
    current_features.points.insert(current_features.points.end(), points_new.begin(), points_new.end());
    current_features.ages.insert(current_features.ages.end(), points_new.size(), 0);
}

// ------------------------------------------------------------------------------------------------ visualOdometry.h
void matchingFeatures(cv::Mat& imageLeft_t0, cv::Mat& imageRight_t0, cv::Mat& imageLeft_t1, cv::Mat& imageRight_t1,
                      FeatureSet& currentVOFeatures, std::vector<cv::Point2f>& pointsLeft_t0,
                      std::vector<cv::Point2f>& pointsRight_t0, std::vector<cv::Point2f>& pointsLeft_t1,
                      std::vector<cv::Point2f>& pointsRight_t1)
{
    std::vector<cv::Point2f> pointsLeftReturn_t0;
    if (currentVOFeatures.size() < 2000) appendNewFeatures(imageLeft_t0, currentVOFeatures);
    const int bucket_size = imageLeft_t0.rows / 10;
    bucketingFeatures(imageLeft_t0, currentVOFeatures, bucket_size, /*features_per_bucket=*/1);
    pointsLeft_t0 = currentVOFeatures.points;
    circularMatching(imageLeft_t0, imageRight_t0, imageLeft_t1, imageRight_t1, pointsLeft_t0, pointsRight_t0, pointsLeft_t1,
                     pointsRight_t1, pointsLeftReturn_t0, currentVOFeatures);
    std::vector<uchar> valid;
    checkValidMatch(pointsLeft_t0, pointsLeftReturn_t0, valid, 0);
    const size_t n = valid.size();
    compact(pointsLeft_t0, valid, n); compact(pointsLeft_t1, valid, n);
    compact(pointsRight_t0, valid, n); compact(pointsRight_t1, valid, n);
    currentVOFeatures.points = pointsLeft_t1;          // ages keep their pre-check length (Appendix A item 8)
}

void triangulateStereo(cv::Mat& projMatrl, cv::Mat& projMatrr, std::vector<cv::Point2f>& pointsLeft,
                       std::vector<cv::Point2f>& pointsRight, cv::Mat& points3D)
{
    if (projMatrl.type() != CV_32FC1 || projMatrr.type() != CV_32FC1 || projMatrl.rows != 3 || projMatrl.cols != 4)
        throw std::runtime_error("triangulateStereo: projection matrices must be 3x4 CV_32F (main.cpp:73-74)");
    if (pointsLeft.size() != pointsRight.size()) throw std::runtime_error("triangulateStereo: point count mismatch");
    float Pl[12], Pr[12];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) { Pl[r * 4 + c] = projMatrl.at<float>(r, c); Pr[r * 4 + c] = projMatrr.at<float>(r, c); }
    const int n = (int)pointsLeft.size();
    points3D = cv::Mat(n, 1, CV_32FC3);
    if (n) check(vo_triangulate(context(), Pl, Pr, as_vo(pointsLeft), as_vo(pointsRight), n,
                                reinterpret_cast<vo_point3f*>(points3D.data)), "vo_triangulate");
}

void trackingFrame2Frame(cv::Mat& projMatrl, cv::Mat& /*projMatrr: unused by the reference too*/,
                         std::vector<cv::Point2f>& pointsLeft_t0,
                         std::vector<cv::Point2f>& pointsLeft_t1, cv::Mat& points3D_t0, cv::Mat& rotation,
                         cv::Mat& translation, bool mono_rotation)
{
    if (mono_rotation) {
        // findEssentialMat(RANSAC, 0.999, 1.0) + recoverPose -> `rotation` (src/visualOdometry.cpp:146-157); the PnP below then
        // only provides `translation` (the reference skips its cv::Rodrigues in this mode, :186-189)
        if (pointsLeft_t0.size() != pointsLeft_t1.size()) throw std::runtime_error("trackingFrame2Frame: point count mismatch");
        const double focal = projMatrl.at<float>(0, 0);
        const double ppx = projMatrl.at<float>(0, 2), ppy = projMatrl.at<float>(1, 2);
        double Rm[9];
        check(vo_mono_rotation(context(), as_vo(pointsLeft_t0), as_vo(pointsLeft_t1), (int)pointsLeft_t0.size(), focal, ppx, ppy, Rm,
                               nullptr, nullptr, nullptr), "vo_mono_rotation");
        rotation = cv::Mat(3, 3, CV_64FC1);
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) rotation.at<double>(r, c) = Rm[r * 3 + c];
    }
    float K[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) K[r * 3 + c] = projMatrl.at<float>(r, c);
    const int n = (int)pointsLeft_t1.size();
    if (points3D_t0.rows * points3D_t0.cols * points3D_t0.channels() != 3 * n || points3D_t0.depth() != CV_32F)
        throw std::runtime_error("trackingFrame2Frame: points3D_t0 must hold N float triples (N x 1 CV_32FC3)");
    double rvec[3] = {0, 0, 0};                                     // reset every call (visualOdometry.cpp:162)
    double tvec[3] = {translation.at<double>(0), translation.at<double>(1), translation.at<double>(2)};
    double R[9];
    std::vector<int32_t> inl((size_t)std::max(n, 1));
    int n_in = 0, iters = 0;
    check(vo_pnp_ransac(context(), reinterpret_cast<const vo_point3f*>(points3D_t0.data), as_vo(pointsLeft_t1), n, K, rvec, tvec,
                        inl.data(), &n_in, R, &iters), "vo_pnp_ransac");
    for (int k = 0; k < 3; k++) translation.at<double>(k) = tvec[k];
    if (!mono_rotation) {                                           // `if (!mono_rotation) cv::Rodrigues(rvec, rotation);`
        if (rotation.empty() || rotation.type() != CV_64FC1 || rotation.rows != 3 || rotation.cols != 3) rotation = cv::Mat(3, 3, CV_64FC1);
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) rotation.at<double>(r, c) = R[r * 3 + c];
    }
    g_last_inliers.assign(inl.begin(), inl.begin() + n_in);
    std::printf("[trackingFrame2Frame] inliers size: %d\n", n_in);
}

void displayTracking(cv::Mat&, std::vector<cv::Point2f>&, std::vector<cv::Point2f>&) {}

// ------------------------------------------------------------------------------------------------ Frame.h
Frame::Frame() {}
Frame::Frame(int, const cv::Mat projMatL, const cv::Mat projMatR, cv::Mat worldRotation, cv::Mat worldTranslation)
    : m_projMatL(projMatL), m_projMatR(projMatR), m_worldRotation(worldRotation), m_worldTranslation(worldTranslation) {}

void Frame::setFeatures(std::vector<cv::Point2f> l, std::vector<cv::Point2f> r)
{
    m_pointsFeatureLeft = l;
    m_pointsFeatureRight = r;
}

void Frame::triangulateFeaturePoints(cv::Mat& points4D)
{
    // cv::triangulatePoints(m_projMatL, m_projMatR, left, right, points4D): 4 x N CV_32F, unit-norm homogeneous columns
    const int n = (int)m_pointsFeatureLeft.size();
    if ((int)m_pointsFeatureRight.size() != n) throw std::runtime_error("Frame::triangulateFeaturePoints: left / right sizes differ");
    points4D = cv::Mat(4, n, CV_32FC1);
    if (n == 0) return;
    if (m_projMatL.type() != CV_32FC1 || m_projMatR.type() != CV_32FC1 || m_projMatL.rows != 3 || m_projMatL.cols != 4)
        throw std::runtime_error("Frame::triangulateFeaturePoints: projection matrices must be 3x4 CV_32F");
    float Pl[12], Pr[12];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) { Pl[r * 4 + c] = m_projMatL.at<float>(r, c); Pr[r * 4 + c] = m_projMatR.at<float>(r, c); }
    std::vector<float> x4((size_t)n * 4);
    check(vo_triangulate_homogeneous(context(), Pl, Pr, as_vo(m_pointsFeatureLeft), as_vo(m_pointsFeatureRight), n, x4.data()),
          "vo_triangulate_homogeneous");
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 4; k++) points4D.at<float>(k, i) = x4[(size_t)i * 4 + k];
}

// ------------------------------------------------------------------------------------------------ utils.h
#include "../../include/compat/utils.h"

static void mat_to_rt(const cv::Mat& rotation, const cv::Mat& translation, double R[9], double t[3])
{
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) R[3 * r + c] = rotation.at<double>(r, c);
        t[r] = translation.at<double>(r);
    }
}

void integrateOdometryStereo(int /*frame_id*/, cv::Mat& rigid_body_transformation, cv::Mat& frame_pose,
                             const cv::Mat& rotation, const cv::Mat& translation_stereo)
{
    double R[9], t[3], pose[16], inv[16];
    mat_to_rt(rotation, translation_stereo, R, t);
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) pose[4 * r + c] = frame_pose.at<double>(r, c);
    const int rc = vo_pose_integrate(pose, R, t, inv);
    if (rc < 0) throw std::runtime_error("integrateOdometryStereo: singular transformation");
    rigid_body_transformation = cv::Mat(4, 4, CV_64FC1);
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) rigid_body_transformation.at<double>(r, c) = inv[4 * r + c];
    if (rc == 1) {
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) frame_pose.at<double>(r, c) = pose[4 * r + c];
    } else {
        std::printf("[WARNING] scale below 0.1, or incorrect translation\n");
    }
}

bool isRotationMatrix(cv::Mat& R)
{
    double r9[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) r9[3 * r + c] = R.at<double>(r, c);
    return vo_pose_is_rotation(r9) != 0;
}

cv::Vec3f rotationMatrixToEulerAngles(cv::Mat& R)
{
    double r9[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) r9[3 * r + c] = R.at<double>(r, c);
    cv::Vec3f v;
    vo_pose_euler(r9, v.val);
    return v;
}

static void load_png_pair(cv::Mat& image_color, cv::Mat& image_gray, int cam, int frame_id, const std::string& filepath)
{
    char name[64];
    std::snprintf(name, sizeof(name), "image_%d/%06d.png", cam, frame_id);
    const std::string path = filepath + name;               // the reference concatenates, no separator added
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("loadImage: cannot open " + path);
    std::vector<unsigned char> bytes;
    unsigned char buf[1 << 16];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) bytes.insert(bytes.end(), buf, buf + n);
    std::fclose(f);
    int w = 0, h = 0;
    if (vo_png_info(bytes.data(), bytes.size(), &w, &h, nullptr, nullptr) != VO_OK)
        throw std::runtime_error("loadImage: " + path + ": " + vo_png_last_error());
    image_color = cv::Mat(h, w, CV_8UC3);
    image_gray = cv::Mat(h, w, CV_8UC1);
    if (vo_png_decode(bytes.data(), bytes.size(), image_color.data, image_color.step, image_gray.data, image_gray.step) != VO_OK)
        throw std::runtime_error("loadImage: " + path + ": " + vo_png_last_error());
}

void loadImageLeft(cv::Mat& image_color, cv::Mat& image_gray, int frame_id, std::string filepath) { load_png_pair(image_color, image_gray, 0, frame_id, filepath); }
void loadImageRight(cv::Mat& image_color, cv::Mat& image_gray, int frame_id, std::string filepath) { load_png_pair(image_color, image_gray, 1, frame_id, filepath); }
