// stand-in for opencv2/video.hpp: the DenseOpticalFlow base interface (video/tracking.hpp)
#pragma once
#include "opencv2/core.hpp"
namespace cv {
class DenseOpticalFlow : public Algorithm {
public:
    virtual void calc(InputArray I0, InputArray I1, InputOutputArray flow) = 0;
    virtual void collectGarbage() = 0;
};
}  // namespace cv
