// opencv2/core.hpp -- ORACLE BUILD SHIM (test infrastructure, not product code).
// OpenCV C++ is not installed in this image; the reference hook (vtm-mlt-cpp/source/Lib/EncoderLib/EncCu.cpp:810-867)
// needs exactly: cv::Size, cv::Mat over a caller-owned CV_16UC1 buffer, cv::absdiff on two such Mats, Mat::convertTo to
// CV_32FC1 with a scale, rows / cols / data / at<float>() / release().  This header provides those and nothing else so
// that the reference's own translation unit compiles unmodified (oracle/vtm/Makefile).  Arithmetic:
//   absdiff   : exact |a - b| on uint16                                         (OpenCV: absdiff16u)
//   convertTo : dst = (float)src * (float)alpha + (float)beta, one fp32 fma     (OpenCV: cvt_32f / cvtScale16u32f);
//               pinned against cv2 4.13 on all 1024 codes by tests/golden/stage_kat.npz (tools/gen_golden.py)
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>

#define CV_16UC1 2
#define CV_32FC1 5

namespace cv {

struct Size {
    int width, height;
    Size(int w, int h) : width(w), height(h) {}
};

class Mat {
public:
    int rows = 0, cols = 0, type_ = 0;
    unsigned char *data = nullptr;

    Mat() {}
    Mat(Size s, int type, void *ext) : rows(s.height), cols(s.width), type_(type), data((unsigned char *)ext) {}

    void create(int r, int c, int type)
    {
        rows = r, cols = c, type_ = type;
        own_.reset(new unsigned char[(size_t)r * c * (type == CV_32FC1 ? 4 : 2)]);
        data = own_.get();
    }
    void release()
    {
        own_.reset();
        data = nullptr;
        rows = cols = 0;
    }
    template <typename T> T &at(int i, int j) { return ((T *)data)[(size_t)i * cols + j]; }

    void convertTo(Mat &dst, int rtype, double alpha = 1, double beta = 0) const
    {
        if (type_ != CV_16UC1 || rtype != CV_32FC1) std::abort(); // only the conversion the hook performs
        dst.create(rows, cols, rtype);
        const float a = (float)alpha, b = (float)beta;
        const uint16_t *s = (const uint16_t *)data;
        float *d = (float *)dst.data;
        for (size_t k = 0, n = (size_t)rows * cols; k < n; k++) d[k] = std::fmaf((float)s[k], a, b);
    }

private:
    std::shared_ptr<unsigned char[]> own_;
};

inline void absdiff(const Mat &a, const Mat &b, Mat &dst)
{
    if (a.type_ != CV_16UC1 || b.type_ != CV_16UC1 || a.rows != b.rows || a.cols != b.cols) std::abort();
    dst.create(a.rows, a.cols, CV_16UC1);
    const uint16_t *x = (const uint16_t *)a.data, *y = (const uint16_t *)b.data;
    uint16_t *d = (uint16_t *)dst.data;
    for (size_t k = 0, n = (size_t)a.rows * a.cols; k < n; k++) d[k] = x[k] > y[k] ? x[k] - y[k] : y[k] - x[k];
}

} // namespace cv
