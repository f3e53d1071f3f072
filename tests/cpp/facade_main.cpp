// facade_main.cpp -- a miniature of the reference's main loop (reference src/main.cpp:123-224)
// written against the compat headers, used by tests/test_gpu_facade.py.
//   usage: facade_main <in.bin> <out.bin>
//   in : int32 w, h, n_frames ; float P_l[12], P_r[12] ; then n_frames x (left, right) u8 images
//   out: per processed frame pair: int32 n ; n x (pL0, pR0, pL1, pR1) float2 ; n x float3 X ;
//        int32 n_inl ; n_inl x int32 ; double R[9] ; double t[3] ; int32 n_features_after (currentVOFeatures) ;
//        double frame_pose[16] after the Euler gate + integrateOdometryStereo (reference src/main.cpp:196-208) ;
//        4 x n float: Frame::triangulateFeaturePoints of (pL0, pR0) (reference src/Frame.cpp:25-28) ;
//        double R_mono[9], t_mono[3]: the same trackingFrame2Frame call with the flag's header default (mono_rotation = true)
#include "visualOdometry.h"
#include "utils.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

static void rd(FILE* f, void* p, size_t n) { if (fread(p, 1, n, f) != n) { std::fprintf(stderr, "short read\n"); std::exit(2); } }

int main(int argc, char** argv)
{
    if (argc < 3) return 1;
    FILE* fi = std::fopen(argv[1], "rb"); FILE* fo = std::fopen(argv[2], "wb");
    if (!fi || !fo) return 1;
    int32_t w, h, nf;
    rd(fi, &w, 4); rd(fi, &h, 4); rd(fi, &nf, 4);
    cv::Mat projMatrl(3, 4, CV_32FC1), projMatrr(3, 4, CV_32FC1);
    rd(fi, projMatrl.data, 48); rd(fi, projMatrr.data, 48);
    std::vector<cv::Mat> L(nf), R(nf);
    for (int i = 0; i < nf; i++) {
        L[i] = cv::Mat(h, w, CV_8UC1); R[i] = cv::Mat(h, w, CV_8UC1);
        rd(fi, L[i].data, (size_t)w * h); rd(fi, R[i].data, (size_t)w * h);
    }
    cv::Mat rotation = cv::Mat::eye(3, 3, CV_64FC1);
    cv::Mat translation = cv::Mat::zeros(3, 1, CV_64FC1);
    FeatureSet currentVOFeatures;
    cv::Mat frame_pose = cv::Mat::eye(4, 4, CV_64FC1);
    cv::Mat imageLeft_t0 = L[0], imageRight_t0 = R[0];
    for (int frame_id = 1; frame_id < nf; frame_id++) {
        cv::Mat imageLeft_t1 = L[frame_id], imageRight_t1 = R[frame_id];
        std::vector<cv::Point2f> pointsLeft_t0, pointsRight_t0, pointsLeft_t1, pointsRight_t1;
        matchingFeatures(imageLeft_t0, imageRight_t0, imageLeft_t1, imageRight_t1, currentVOFeatures,
                         pointsLeft_t0, pointsRight_t0, pointsLeft_t1, pointsRight_t1);
        imageLeft_t0 = imageLeft_t1; imageRight_t0 = imageRight_t1;
        cv::Mat points3D_t0;
        triangulateStereo(projMatrl, projMatrr, pointsLeft_t0, pointsRight_t0, points3D_t0);
        cv::Mat rot_mono = cv::Mat::eye(3, 3, CV_64FC1), tr_mono = translation.clone();
        trackingFrame2Frame(projMatrl, projMatrr, pointsLeft_t0, pointsLeft_t1, points3D_t0, rotation, translation, false);
        const int32_t n = (int32_t)pointsLeft_t0.size();
        std::fwrite(&n, 4, 1, fo);
        std::fwrite(pointsLeft_t0.data(), 8, n, fo); std::fwrite(pointsRight_t0.data(), 8, n, fo);
        std::fwrite(pointsLeft_t1.data(), 8, n, fo); std::fwrite(pointsRight_t1.data(), 8, n, fo);
        std::fwrite(points3D_t0.data, 12, n, fo);
        const std::vector<int>& inl = lastPnPInliers();
        const int32_t ni = (int32_t)inl.size();
        std::fwrite(&ni, 4, 1, fo); std::fwrite(inl.data(), 4, ni, fo);
        std::fwrite(rotation.data, 8, 9, fo); std::fwrite(translation.data, 8, 3, fo);
        const int32_t na = (int32_t)currentVOFeatures.points.size();
        std::fwrite(&na, 4, 1, fo);
        cv::Vec3f e = rotationMatrixToEulerAngles(rotation);
        cv::Mat rigid_body_transformation;
        if (std::fabs(e[1]) < 0.1 && std::fabs(e[0]) < 0.1 && std::fabs(e[2]) < 0.1)
            integrateOdometryStereo(frame_id, rigid_body_transformation, frame_pose, rotation, translation);
        std::fwrite(frame_pose.data, 8, 16, fo);
        Frame fr(frame_id, projMatrl, projMatrr, rotation, translation);       // the reference's (otherwise unused) holder class
        fr.setFeatures(pointsLeft_t0, pointsRight_t0);
        cv::Mat points4D;
        fr.triangulateFeaturePoints(points4D);
        for (int r = 0; r < 4; r++) std::fwrite(points4D.ptr<float>(r), 4, n, fo);
        // a caller that omits the flag gets the reference's default, mono_rotation = true (src/visualOdometry.h:42)
        trackingFrame2Frame(projMatrl, projMatrr, pointsLeft_t0, pointsLeft_t1, points3D_t0, rot_mono, tr_mono);
        std::fwrite(rot_mono.data, 8, 9, fo); std::fwrite(tr_mono.data, 8, 3, fo);
    }
    std::fclose(fi); std::fclose(fo);
    return 0;
}
