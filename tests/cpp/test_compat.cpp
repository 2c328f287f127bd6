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
    if (argc >= 8) {
        // the consumer right after calc (cudalegacy interpolateFrames): forward flow as computed, backward flow
        // approximated by its negation, frames as CV_32FC1 in [0, 1]; the Python side repeats the same call
        std::vector<float> f0((size_t)rows * cols), f1(f0.size()), pu(f0.size()), pv(f0.size()), nu(f0.size()), nv(f0.size());
        for (size_t i = 0; i < f0.size(); ++i) {
            f0[i] = h0[i] / 255.f;
            f1[i] = h1[i] / 255.f;
            pu[i] = out[2 * i];
            pv[i] = out[2 * i + 1];
            nu[i] = -pu[i];
            nv[i] = -pv[i];
        }
        cvcuda::GpuMat g0(rows, cols, 5 /*CV_32FC1*/), g1(rows, cols, 5), gu(rows, cols, 5), gv(rows, cols, 5),
            hu(rows, cols, 5), hv(rows, cols, 5), mid, buf;
        const size_t hs = (size_t)cols * 4;
        g0.upload(f0.data(), hs, stream);
        g1.upload(f1.data(), hs, stream);
        gu.upload(pu.data(), hs, stream);
        gv.upload(pv.data(), hs, stream);
        hu.upload(nu.data(), hs, stream);
        hv.upload(nv.data(), hs, stream);
        try {
            // corrected mode: the reference mode's coverage-clear defect depends on the GpuMat pitch (cudaMallocPitch pads
            // 168 floats to 256), which the Python call site (contiguous tensors) does not share
            cvcuda::interpolateFrames(g0, g1, gu, gv, hu, hv, 0.5f, mid, buf, stream, true);
        } catch (const std::exception &e) {
            fprintf(stderr, "exception: %s\n", e.what());
            return 6;
        }
        if (buf.rows != 6 * rows || mid.rows != rows || mid.type() != 5) return 7;
        std::vector<float> m((size_t)rows * cols);
        mid.download(m.data(), hs, stream);
        stream.waitForCompletion();
        FILE *g = fopen(argv[7], "wb");
        fwrite(m.data(), 4, m.size(), g);
        fclose(g);
    }
    if (argc >= 8) {
        // the sparse sibling through the same adapter: three interior points must be tracked
        const float pts[6] = {cols * 0.5f, rows * 0.5f, cols * 0.25f, rows * 0.6f, cols * 0.7f, rows * 0.3f};
        cvcuda::GpuMat dp(1, 3, 13 /*CV_32FC2*/), dn, ds, de;
        dp.upload(pts, sizeof(pts), stream);
        try {
            auto sp = cvcuda::SparsePyrLKOpticalFlow::create();
            if (sp->getWinSize().width != 21 || sp->getMaxLevel() != 3 || sp->getNumIters() != 30) return 8;
            sp->calc(d0, d1, dp, dn, ds, &de, stream);
        } catch (const std::exception &e) {
            fprintf(stderr, "exception: %s\n", e.what());
            return 9;
        }
        unsigned char st[3] = {0, 0, 0};
        float np[6];
        ds.download(st, 3, stream);
        dn.download(np, sizeof(np), stream);
        stream.waitForCompletion();
        if (dn.cols != 3 || ds.cols != 3 || de.cols != 3 || !(st[0] && st[1] && st[2])) return 10;
        printf("sparse %.2f %.2f\n", np[0] - pts[0], np[1] - pts[1]);
    }
    printf("ok %d %d type=%d step=%zu\n", flow.rows, flow.cols, flow.type(), flow.step);
    return 0;
}
