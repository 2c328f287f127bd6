// stand-in for opencv2/imgproc.hpp: the three primitives tvl1flow.cpp calls.  Implemented in
// oracle/ref_glue.cpp on top of the C restatements in oracle/tvl1_cpu.c, which tests/test_oracle_cpu.py
// pins against the live cv2 functions (remap bit-exact, resize <= 2 ulp, medianBlur bit-exact).
#pragma once
#include "opencv2/core.hpp"
namespace cv {
enum { INTER_LINEAR = 1, INTER_CUBIC = 2 };
void resize(const Mat_<float> &src, Mat_<float> &dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void remap(const Mat_<float> &src, Mat_<float> &dst, const Mat_<float> &map1, const Mat_<float> &map2, int interpolation);
void medianBlur(const Mat_<float> &src, Mat_<float> &dst, int ksize);
}  // namespace cv
