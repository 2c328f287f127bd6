// Compile-only check of the REAL-OpenCV branch of cudaoptflow_compat.hpp (every class instantiated, every calc called)
// against the declaration-only core stand-ins of tests/cpp/opencv_stub/ -- nothing is linked or run.
#include <b200flow/cudaoptflow_compat.hpp>
namespace cvcuda = b200flow::cuda;
void use(cv::cuda::GpuMat &a, cv::cuda::GpuMat &b, cv::cuda::GpuMat &f, cv::cuda::GpuMat &pts, cv::cuda::GpuMat &st, cv::cuda::GpuMat &err) {
    cvcuda::OpticalFlowDual_TVL1::create()->calc(a, b, f);
    cvcuda::FarnebackOpticalFlow::create()->calc(a, b, f, cv::cuda::Stream::Null());
    cvcuda::BroxOpticalFlow::create()->calc(a, b, f);
    cvcuda::DensePyrLKOpticalFlow::create(cv::Size(13, 13))->calc(a, b, f);
    auto s = cvcuda::SparsePyrLKOpticalFlow::create();
    s->calc(a, b, pts, f, st);             // err = cv::noArray()
    s->calc(a, b, pts, f, st, err);
    cvcuda::interpolateFrames(a, b, f, f, f, f, 0.5f, st, err);
}
