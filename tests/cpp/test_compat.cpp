// Exercises the C++ adapter (cudaoptflow_compat.hpp) exactly the way a cudaoptflow call site does
// (cf. modules/cudaoptflow/samples/optical_flow.cpp:170-238): upload two frames, create(), calc(),
// download.  Reads raw u8 frames, writes the raw float flow; tests/test_compat_gpu.py compares it
// bit-for-bit with the ctypes path.   usage: test_compat <tvl1|farneback> rows cols in0 in1 out
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "b200flow/cudaoptflow_compat.hpp"

namespace cvcuda = b200flow::cuda;

static std::vector<unsigned char> slurp(const char *p, size_t n) {
    std::vector<unsigned char> v(n);
    FILE *f = fopen(p, "rb");
    if (!f || fread(v.data(), 1, n, f) != n) {
        fprintf(stderr, "cannot read %s\n", p);
        exit(2);
    }
    fclose(f);
    return v;
}

int main(int argc, char **argv) {
    if (argc < 7) return 2;
    const int rows = atoi(argv[2]), cols = atoi(argv[3]);
    auto h0 = slurp(argv[4], (size_t)rows * cols), h1 = slurp(argv[5], (size_t)rows * cols);
    cudaStream_t s;
    cudaStreamCreate(&s);
    cvcuda::Stream stream(s);
    cvcuda::GpuMat d0(rows, cols, 0 /*CV_8UC1*/), d1(rows, cols, 0), flow;
    d0.upload(h0.data(), cols, stream);
    d1.upload(h1.data(), cols, stream);
    cvcuda::Ptr<cvcuda::DenseOpticalFlow> alg;
    try {
        if (!strcmp(argv[1], "tvl1")) {
            auto a = cvcuda::OpticalFlowDual_TVL1::create(0.25, 0.15, 0.3, 3, 3, 0.0, 20);
            if (a->getNumIterations() != 20 || a->getDefaultName() != "DenseOpticalFlow.OpticalFlowDual_TVL1") return 3;
            a->setNumWarps(2);
            alg = a;
        } else {
            auto a = cvcuda::FarnebackOpticalFlow::create();
            if (a->getWinSize() != 13 || a->getDefaultName() != "DenseOpticalFlow.FarnebackOpticalFlow") return 3;
            alg = a;
        }
        alg->calc(d0, d1, flow, stream);  // flow is allocated by calc, as in the reference
        // precondition failures surface as exceptions with the reference's codes
        cvcuda::GpuMat bad(rows, cols + 1, 0);
        bool threw = false;
        try {
            alg->calc(d0, bad, flow, stream);
        } catch (const std::exception &) {
            threw = true;
        }
        if (!threw) return 4;
    } catch (const std::exception &e) {
        fprintf(stderr, "exception: %s\n", e.what());
        return 5;
    }
    std::vector<float> out((size_t)rows * cols * 2);
    flow.download(out.data(), (size_t)cols * 8, stream);
    stream.waitForCompletion();
    FILE *f = fopen(argv[6], "wb");
    fwrite(out.data(), 4, out.size(), f);
    fclose(f);
    printf("ok %d %d type=%d step=%zu\n", flow.rows, flow.cols, flow.type(), flow.step);
    return 0;
}
