// stand-in for opencv2/optflow.hpp: only the abstract interface the reference's tvl1flow.cpp implements
// (modules/optflow/include/opencv2/optflow.hpp, class DualTVL1OpticalFlow and createOptFlow_DualTVL1).
#pragma once
#include "opencv2/core.hpp"
#include "opencv2/video.hpp"
namespace cv {
namespace optflow {
class DualTVL1OpticalFlow : public DenseOpticalFlow {
public:
#define TVL1_PROP(T, Name) virtual T get##Name() const = 0; virtual void set##Name(T val) = 0;
    TVL1_PROP(double, Tau) TVL1_PROP(double, Lambda) TVL1_PROP(double, Theta) TVL1_PROP(double, Gamma)
    TVL1_PROP(int, ScalesNumber) TVL1_PROP(int, WarpingsNumber) TVL1_PROP(double, Epsilon)
    TVL1_PROP(int, InnerIterations) TVL1_PROP(int, OuterIterations) TVL1_PROP(bool, UseInitialFlow)
    TVL1_PROP(double, ScaleStep) TVL1_PROP(int, MedianFiltering)
#undef TVL1_PROP
    static Ptr<DualTVL1OpticalFlow> create(double tau = 0.25, double lambda = 0.15, double theta = 0.3, int nscales = 5,
                                           int warps = 5, double epsilon = 0.01, int innnerIterations = 30,
                                           int outerIterations = 10, double scaleStep = 0.8, double gamma = 0.0,
                                           int medianFiltering = 5, bool useInitialFlow = false);
};
Ptr<DualTVL1OpticalFlow> createOptFlow_DualTVL1();
}  // namespace optflow
}  // namespace cv
