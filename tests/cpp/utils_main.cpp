// Host-only driver for include/compat/utils.h (no GPU work): reads n, then n x (double R[9], double t[3]) from argv[1],
// applies the reference main loop's Euler gate + integrateOdometryStereo (src/main.cpp:196-208) and writes per step:
// float euler[3] ; int32 is_rotation ; double frame_pose[16].
#include "utils.h"
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
// "load <dir/> <frame_id> <out>" mode: loadImageLeft + loadImageRight, writes int32 w,h then color(3wh) gray(wh) of
// the left image and the same for the right one.
#include <cstring>
#include <string>
static int load_mode(char** argv)
{
    cv::Mat c, g;
    FILE* fo = std::fopen(argv[4], "wb");
    if (!fo) return 3;
    for (int cam = 0; cam < 2; cam++) {
        try {
            if (cam == 0) loadImageLeft(c, g, std::atoi(argv[3]), argv[2]);
            else loadImageRight(c, g, std::atoi(argv[3]), argv[2]);
        } catch (const std::exception& e) { std::fprintf(stderr, "%s\n", e.what()); return 6; }
        const int32_t wh[2] = {c.cols, c.rows};
        std::fwrite(wh, 4, 2, fo);
        for (int r = 0; r < c.rows; r++) std::fwrite(c.data + r * c.step, 1, (size_t)3 * c.cols, fo);
        for (int r = 0; r < g.rows; r++) std::fwrite(g.data + r * g.step, 1, (size_t)g.cols, fo);
    }
    std::fclose(fo);
    return 0;
}

int main(int argc, char** argv)
{
    if (argc == 5 && !std::strcmp(argv[1], "load")) return load_mode(argv);
    if (argc < 3) return 2;
    FILE* fi = std::fopen(argv[1], "rb");
    FILE* fo = std::fopen(argv[2], "wb");
    if (!fi || !fo) return 3;
    int32_t n = 0;
    if (std::fread(&n, 4, 1, fi) != 1) return 4;
    cv::Mat frame_pose = cv::Mat::eye(4, 4, CV_64FC1);
    for (int i = 0; i < n; i++) {
        cv::Mat rotation(3, 3, CV_64FC1), translation(3, 1, CV_64FC1);
        if (std::fread(rotation.data, 8, 9, fi) != 9 || std::fread(translation.data, 8, 3, fi) != 3) return 5;
        cv::Vec3f e = rotationMatrixToEulerAngles(rotation);
        const int32_t isrot = isRotationMatrix(rotation) ? 1 : 0;
        cv::Mat rigid_body_transformation;
        if (std::fabs(e[1]) < 0.1 && std::fabs(e[0]) < 0.1 && std::fabs(e[2]) < 0.1)
            integrateOdometryStereo(i, rigid_body_transformation, frame_pose, rotation, translation);
        std::fwrite(&e[0], 4, 3, fo);
        std::fwrite(&isrot, 4, 1, fo);
        std::fwrite(frame_pose.data, 8, 16, fo);
    }
    std::fclose(fi); std::fclose(fo);
    return 0;
}
