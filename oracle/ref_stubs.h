// TEST INFRASTRUCTURE.  Minimal stand-ins for the OpenCV / PCL / ROS types that the reference's reprojection code touches, so
// that the reference's OWN statements -- extracted at build time from /root/reference/src/disparity_to_depth.cpp and
// /root/reference/src/generate_disparity.cpp by oracle/build_ref.py, never copied into this repository -- can be compiled and
// run here (OpenCV C++, PCL and ROS are not installed).  Only the members those statements use exist; semantics follow the
// real libraries for exactly those uses (element access, zeros, convertTo with a scale, compare-with-scalar, masked setTo).
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

typedef unsigned char uchar;
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32F 5
#define CV_32FC1 5
#define CV_64F 6

namespace cv {

struct Vec3b {
    uchar v[3];
    uchar& operator[](int i) { return v[i]; }
    const uchar& operator[](int i) const { return v[i]; }
};

class Mat {
public:
    int rows = 0, cols = 0;
    uchar* data = nullptr;
    size_t step = 0;
    Mat() = default;
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* ext, size_t stp = 0) : rows(r), cols(c), data(static_cast<uchar*>(ext)), type_(type)
    {
        step = stp ? stp : size_t(c) * esz(type);
    }
    static size_t esz(int type) { return type == CV_8UC1 ? 1 : type == CV_8UC3 ? 3 : type == CV_32F ? 4 : 8; }
    void create(int r, int c, int type)
    {
        rows = r; cols = c; type_ = type; step = size_t(c) * esz(type);
        buf_ = std::make_shared<std::vector<uchar>>(step * size_t(r), uchar(0));
        data = buf_->data();
    }
    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); return m; }
    int type() const { return type_; }
    template <typename T> T& at(int y, int x) { return reinterpret_cast<T*>(data + size_t(y) * step)[x]; }
    template <typename T> const T& at(int y, int x) const { return reinterpret_cast<const T*>(data + size_t(y) * step)[x]; }
    // cv::Mat::convertTo(dst, CV_32F, alpha) on CV_32F data: OpenCV's cvt32f path computes saturate_cast<float>(src * (float)alpha)
    // (a power-of-two alpha makes the float / double distinction immaterial)
    void convertTo(Mat& dst, int rtype, double alpha = 1.0) const
    {
        if (dst.data == nullptr || dst.rows != rows || dst.cols != cols) dst.create(rows, cols, rtype);
        for (int y = 0; y < rows; y++)
            for (int x = 0; x < cols; x++) dst.at<float>(y, x) = float(at<float>(y, x) * float(alpha));
    }
    // dmat.setTo(value, mask): masked assignment
    void setTo(double v, const Mat& mask)
    {
        for (int y = 0; y < rows; y++)
            for (int x = 0; x < cols; x++)
                if (mask.at<uchar>(y, x)) at<float>(y, x) = float(v);
    }

protected:
    int type_ = CV_8UC1;
    std::shared_ptr<std::vector<uchar>> buf_;
};

template <typename T> struct MatType;
template <> struct MatType<float> { static const int value = CV_32F; };
template <> struct MatType<uint8_t> { static const int value = CV_8UC1; };
template <> struct MatType<Vec3b> { static const int value = CV_8UC3; };

template <typename T> class Mat_ : public Mat {
public:
    Mat_(int r, int c, T* ext, size_t stp = 0) : Mat(r, c, MatType<T>::value, ext, stp) {}
};

// `dmat < s` / `dmat > s` with a scalar: cv::compare converts the scalar to double and compares element-wise in the element type's
// exact value domain (float -> double is exact), giving a 0 / 255 mask
inline Mat cmp_(const Mat& a, double s, bool less)
{
    Mat m(a.rows, a.cols, CV_8UC1);
    for (int y = 0; y < a.rows; y++)
        for (int x = 0; x < a.cols; x++) {
            const double v = double(a.at<float>(y, x));
            m.at<uchar>(y, x) = (less ? v < s : v > s) ? 255 : 0;
        }
    return m;
}
inline Mat operator<(const Mat& a, double s) { return cmp_(a, s, true); }
inline Mat operator>(const Mat& a, double s) { return cmp_(a, s, false); }

}  // namespace cv

namespace pcl {
struct PointXYZRGB {
    float x = 0, y = 0, z = 0;
    uchar r = 0, g = 0, b = 0;
};
template <typename P> struct PointCloud {
    typedef std::shared_ptr<PointCloud<P>> Ptr;
    std::vector<P> points;
    void push_back(const P& p) { points.push_back(p); }
    size_t size() const { return points.size(); }
};
}  // namespace pcl
