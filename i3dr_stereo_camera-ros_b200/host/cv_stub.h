// Minimal stand-in for the handful of OpenCV core types the matcher plugin layer touches.
// ONLY used for the ROS/OpenCV-free build of the adapter and its harness in this repository (OpenCV C++ and
// ROS are not available in the build image).  In the reference tree the adapter is compiled against the real
// <opencv2/opencv.hpp>; nothing here is part of the product's data path.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <vector>

#define CV_8UC1 0
#define CV_16S 3
#define CV_16SC1 3
#define CV_32F 5
#define CV_32FC1 5

namespace cv {

struct Size {
    int width = 0, height = 0;
    Size() = default;
    Size(int w, int h) : width(w), height(h) {}
    bool operator==(const Size& o) const { return width == o.width && height == o.height; }
};

struct Scalar {
    double v = 0;
    Scalar(double x = 0) : v(x) {}
};

enum { INTER_CUBIC = 2 };

class Mat {
public:
    int rows = 0, cols = 0;
    uint8_t* data = nullptr;
    size_t step = 0;

    Mat() = default;
    Mat(Size s, int type) { create(s.height, s.width, type); }
    Mat(Size s, int type, Scalar fill) { create(s.height, s.width, type); setTo(fill.v); }
    Mat(int r, int c, int type) { create(r, c, type); }
    // header over caller-owned data, like cv::Mat(rows, cols, type, data, step)
    Mat(int r, int c, int type, void* ext, size_t stp = 0) : rows(r), cols(c), data(static_cast<uint8_t*>(ext)), type_(type)
    {
        step = stp ? stp : size_t(c) * elem_size(type);
    }

    static size_t elem_size(int type) { return type == CV_8UC1 ? 1 : (type == CV_16S ? 2 : 4); }
    void create(int r, int c, int type)
    {
        rows = r; cols = c; type_ = type; step = size_t(c) * elem_size(type);
        buf_ = std::make_shared<std::vector<uint8_t>>(step * size_t(r));
        data = buf_->data();
    }
    static Mat zeros(Size s, int type) { Mat m(s, type); std::memset(m.data, 0, m.step * m.rows); return m; }
    void setTo(double v)
    {
        for (int y = 0; y < rows; y++)
            for (int x = 0; x < cols; x++) {
                if (type_ == CV_8UC1) data[y * step + x] = uint8_t(v);
                else if (type_ == CV_16S) reinterpret_cast<int16_t*>(data + y * step)[x] = int16_t(v);
                else reinterpret_cast<float*>(data + y * step)[x] = float(v);
            }
    }
    int type() const { return type_; }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return step == size_t(cols) * elem_size(type_); }
    template <typename T> T& at(int y, int x) { return reinterpret_cast<T*>(data + y * step)[x]; }
    template <typename T> const T& at(int y, int x) const { return reinterpret_cast<const T*>(data + y * step)[x]; }

    void copyTo(Mat& dst) const
    {
        if (dst.rows != rows || dst.cols != cols || dst.type_ != type_ || dst.data == data) dst.create(rows, cols, type_);
        for (int y = 0; y < rows; y++) std::memcpy(dst.data + y * dst.step, data + y * step, size_t(cols) * elem_size(type_));
    }
    void convertTo(Mat& dst, int rtype, double alpha = 1.0) const
    {
        if (rtype == type_ && alpha == 1.0) {   // OpenCV: same depth, no scale -> copyTo, which returns at once when dst is this matrix
            if (dst.data != data) copyTo(dst);
            return;
        }
        Mat out(rows, cols, rtype);
        for (int y = 0; y < rows; y++)
            for (int x = 0; x < cols; x++) {
                double v = type_ == CV_8UC1 ? at<uint8_t>(y, x) : (type_ == CV_16S ? at<int16_t>(y, x) : at<float>(y, x));
                v *= alpha;
                if (rtype == CV_32F) out.at<float>(y, x) = float(v);
                else if (rtype == CV_16S) out.at<int16_t>(y, x) = int16_t(v);
                else out.at<uint8_t>(y, x) = uint8_t(v);
            }
        dst = out;
    }

private:
    int type_ = CV_8UC1;
    std::shared_ptr<std::vector<uint8_t>> buf_;
};

// Only the identity case the node uses (setDownsampleScale(1), generate_disparity.cpp:349) is supported here.
inline void resize(const Mat& src, Mat& dst, Size, double fx, double fy, int)
{
    if (fx != 1.0 || fy != 1.0) throw std::runtime_error("cv_stub: resize only supports scale 1");
    src.copyTo(dst);
}

}  // namespace cv
