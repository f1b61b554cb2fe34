// svo_ref_stubs.h -- TEST INFRASTRUCTURE.  Stand-in declarations for the third-party headers that
// /root/reference/libstereo-odometry/include/libstereo-odometry.h includes (OpenCV, MRPT 1.x, Eigen, stlplus), just
// wide enough to compile the reference's OWN, UNMODIFIED sources (common.cpp, process_new_image_pair.cpp,
// stage1_rectify.cpp .. stage5_optimization.cpp) into oracle/_ref/libsvo_ref.so (recipe: oracle/Makefile.ref).
// None of those libraries is installed in this image.  What the stand-ins do:
//   * containers and PODs (cv::Mat, cv::KeyPoint, cv::DMatch, CImage, smart pointers, CMatrixB ...): plain C++ with the
//     semantics the reference relies on (Mat headers share their buffer, Mat::row() is a view, push_back appends rows);
//   * the OpenCV algorithms (FAST, ORB::compute, ORB::detectAndCompute, BFMatcher::match, findFundamentalMat, pyrDown)
//     are forwarded through a table of C callbacks (svo_ref_backend) that oracle/ref.py fills with the real OpenCV of
//     this image (cv2) -- so the compiled reference text runs on the real third-party arithmetic;
//   * Eigen::MatrixXd / JacobiSVD: a small dense matrix and a restatement of Eigen's two-sided Jacobi SVD and its
//     solve() (Eigen's text is not available here);
//   * MRPT pose classes, KeypointResponseSorter, buildPyramidFast: restated from MRPT 1.x semantics (UNPINNED, same
//     assumptions as DESIGN.md section 4);
//   * everything off the hot path (GUI, KLT, FASTER, SAD, rectification maps, optical flow) aborts when reached.
// Nothing under stereo_vo_b200/ may include this file.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <sys/stat.h>
#include <vector>

#if __cplusplus >= 201703L
namespace std {  // removed from C++17; the reference derives its comparators from it (libstereo-odometry.h:111,120)
template <class A, class B, class R> struct binary_function { typedef A first_argument_type; typedef B second_argument_type; typedef R result_type; };
}
#endif

#define CV_MAJOR_VERSION 4
#define MRPT_VERSION 0x130

[[noreturn]] inline void svo_ref_off_path(const char* what) {
    std::fprintf(stderr, "[oracle/_ref] '%s' is off the hot path and not provided by the stand-in headers\n", what);
    std::abort();
}

// ---------------------------------------------------------------------------------------------------------------
// callbacks into the real OpenCV (filled by oracle/ref.py through svo_ref_set_backend)
// ---------------------------------------------------------------------------------------------------------------
extern "C" {
struct svo_ref_kp { float x, y, size, angle, response; int32_t octave, class_id; };      // = cv::KeyPoint
struct svo_ref_dm { int32_t queryIdx, trainIdx, imgIdx; float distance; };              // = cv::DMatch
struct svo_ref_backend {
    int (*fast)(const uint8_t* img, int w, int h, int stride, int threshold, svo_ref_kp* out, int cap);
    int (*orb_compute)(const uint8_t* img, int w, int h, int stride, svo_ref_kp* kps_inout, int n_in, uint8_t* desc_out);
    int (*orb_detect_and_compute)(const uint8_t* img, int w, int h, int stride, int nfeatures, float scale, int nlevels,
                                  int edge, int first_level, int wta_k, int score_type, int patch, int fast_th,
                                  svo_ref_kp* out, int cap, uint8_t* desc_out);
    int (*bf_match)(const uint8_t* q, int nq, const uint8_t* t, int nt, svo_ref_dm* out);
    // returns the mask length (0: no mask), < 0: OpenCV raised an exception
    int (*find_fundamental)(const float* p1, const float* p2, int n, int method, double thr, double conf, uint8_t* mask);
    int (*pyr_down)(const uint8_t* img, int w, int h, int stride, uint8_t* out, int ow, int oh);
};
extern svo_ref_backend g_svo_ref_backend;
}

// ---------------------------------------------------------------------------------------------------------------
// OpenCV
// ---------------------------------------------------------------------------------------------------------------
struct IplImage { int width, height, widthStep, nChannels; char* imageData; };
typedef unsigned char uchar;                                       // OpenCV declares it at global scope (cvdef.h)

namespace cv {
using ::uchar;
using std::max;
using std::min;

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T a, T b) : x(a), y(b) {}
    template <class U> Point_(const Point_<U>& o) : x(T(o.x)), y(T(o.y)) {}
};
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
typedef Point_<int> Point;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
    bool operator==(const Size& o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size& o) const { return !(*this == o); }
};

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};
struct DMatch {
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(FLT_MAX) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
};
static_assert(sizeof(KeyPoint) == 28 && sizeof(DMatch) == 16, "cv::KeyPoint / cv::DMatch layouts");

enum { CV_8U_ = 0, CV_32F_ = 5, CV_64F_ = 6 };
#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 0
#define CV_32FC1 5
#define CV_32FC2 13
#define CV_64FC1 6
template <typename T> struct DataType;
template <> struct DataType<float> { enum { type = CV_32F }; };
template <> struct DataType<double> { enum { type = CV_64F }; };
template <> struct DataType<uchar> { enum { type = CV_8U }; };

template <typename T> using MatIterator_ = T*;
template <typename T> using MatConstIterator_ = const T*;

class Mat {
public:
    int rows, cols;
    Mat() : rows(0), cols(0), type_(0), step_(0), data_(nullptr) {}
    Mat(int r, int c, int type) : rows(0), cols(0), type_(0), step_(0), data_(nullptr) { create(r, c, type); }
    Mat(int r, int c, int type, void* ext, size_t step) : rows(r), cols(c), type_(type), step_(step), data_((uchar*)ext) {}
    template <class T> explicit Mat(const std::vector<Point_<T>>& v) : rows(0), cols(0), type_(0), step_(0), data_(nullptr) {
        create((int)v.size(), 1, DataType<T>::type + 8);          // two channels
        if (!v.empty()) std::memcpy(data_, v.data(), v.size() * sizeof(Point_<T>));
    }
    void create(int r, int c, int type) {
        if (r == rows && c == cols && type == type_ && (data_ || r * c == 0)) return;
        rows = r; cols = c; type_ = type;
        step_ = (size_t)c * elemSize();
        buf_ = std::make_shared<std::vector<uchar>>((size_t)r * step_);
        data_ = buf_->data();
    }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize() const { static const int sz[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return (size_t)sz[depth()] * channels(); }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return rows == 0 || cols == 0; }
    size_t step() const { return step_; }
    uchar* ptr(int r = 0) { return data_ + (size_t)r * step_; }
    const uchar* ptr(int r = 0) const { return data_ + (size_t)r * step_; }
    template <class T> T* ptr(int r = 0) { return (T*)(data_ + (size_t)r * step_); }
    template <class T> const T* ptr(int r = 0) const { return (const T*)(data_ + (size_t)r * step_); }
    template <class T> T& at(int r, int c) { return ((T*)(data_ + (size_t)r * step_))[c]; }
    template <class T> const T& at(int r, int c) const { return ((const T*)(data_ + (size_t)r * step_))[c]; }
    template <class T> T* begin() { return (T*)data_; }
    template <class T> T* end() { return (T*)(data_ + (size_t)rows * step_); }
    template <class T> const T* begin() const { return (const T*)data_; }
    template <class T> const T* end() const { return (const T*)(data_ + (size_t)rows * step_); }
    Mat row(int i) const {                                         // a view: shares the buffer
        Mat m;
        m.rows = 1; m.cols = cols; m.type_ = type_; m.step_ = step_; m.buf_ = buf_;
        m.data_ = data_ + (size_t)i * step_;
        return m;
    }
    void copyTo(const Mat& dst_) const {                           // cv::OutputArray also binds temporaries (row views)
        Mat& dst = const_cast<Mat&>(dst_);
        dst.create(rows, cols, type_);                             // no-op when size and type already match
        const size_t line = (size_t)cols * elemSize();
        for (int r = 0; r < rows; ++r) std::memcpy(dst.ptr(r), ptr(r), line);
    }
    void reserve(size_t) {}
    void push_back(const Mat& m) {                                 // append the rows of m
        if (m.rows == 0) return;
        Mat out(rows + m.rows, empty() ? m.cols : cols, empty() ? m.type_ : type_);
        const size_t line = (size_t)out.cols * out.elemSize();
        for (int r = 0; r < rows; ++r) std::memcpy(out.ptr(r), ptr(r), line);
        for (int r = 0; r < m.rows; ++r) std::memcpy(out.ptr(rows + r), m.ptr(r), line);
        *this = out;
    }
    void release() { *this = Mat(); }
private:
    int type_;
    size_t step_;
    uchar* data_;
    std::shared_ptr<std::vector<uchar>> buf_;
};

template <typename T> using Ptr = std::shared_ptr<T>;

inline Mat cvarrToMat(const IplImage* im) { return Mat(im->height, im->width, CV_8UC1, im->imageData, (size_t)im->widthStep); }

class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage(const std::string& name, int) : f_(name.c_str()) {}
    FileStorage& operator<<(const char* s) { key_ = s; return *this; }
    FileStorage& operator<<(const std::string& s) { key_ = s; return *this; }
    FileStorage& operator<<(const Mat& m) {                        // a readable dump, not cv::FileStorage's YAML dialect
        f_ << key_ << ": rows " << m.rows << " cols " << m.cols << "\n";
        for (int r = 0; r < m.rows; ++r) { for (int c = 0; c < m.cols; ++c) f_ << (int)m.at<uchar>(r, c) << (c + 1 < m.cols ? "," : ""); f_ << "\n"; }
        return *this;
    }
    void release() { f_.close(); }
private:
    std::ofstream f_;
    std::string key_;
};

class FastFeatureDetector {
public:
    static Ptr<FastFeatureDetector> create(int threshold = 10) { auto p = std::make_shared<FastFeatureDetector>(); p->th_ = threshold; return p; }
    void detect(const Mat& img, std::vector<KeyPoint>& kps) const {
        if (!g_svo_ref_backend.fast) svo_ref_off_path("cv::FastFeatureDetector without a backend");
        const int cap = std::max(1024, img.rows * img.cols / 2);
        kps.resize(cap);
        const int n = g_svo_ref_backend.fast(img.ptr(), img.cols, img.rows, (int)img.step(), th_, (svo_ref_kp*)kps.data(), cap);
        if (n < 0 || n > cap) throw std::runtime_error("cv::FastFeatureDetector::detect failed");
        kps.resize(n);
    }
    int th_;
};

class ORB {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };
    static Ptr<ORB> create(int nfeatures = 500, float scaleFactor = 1.2f, int nlevels = 8, int edgeThreshold = 31, int firstLevel = 0,
                           int WTA_K = 2, int scoreType = HARRIS_SCORE, int patchSize = 31, int fastThreshold = 20) {
        auto p = std::make_shared<ORB>();
        p->nfeatures = nfeatures; p->scale = scaleFactor; p->nlevels = nlevels; p->edge = edgeThreshold; p->first = firstLevel;
        p->wta = WTA_K; p->score = scoreType; p->patch = patchSize; p->fast_th = fastThreshold;
        return p;
    }
    void compute(const Mat& img, std::vector<KeyPoint>& kps, Mat& desc) const {
        if (!g_svo_ref_backend.orb_compute) svo_ref_off_path("cv::ORB::compute without a backend");
        std::vector<uchar> d(kps.size() * 32 + 32);
        const int n = g_svo_ref_backend.orb_compute(img.ptr(), img.cols, img.rows, (int)img.step(), (svo_ref_kp*)kps.data(), (int)kps.size(), d.data());
        if (n < 0 || n > (int)kps.size()) throw std::runtime_error("cv::ORB::compute failed");
        kps.resize(n);
        desc = Mat();
        if (n > 0) { desc.create(n, 32, CV_8UC1); std::memcpy(desc.ptr(), d.data(), (size_t)n * 32); }
    }
    void detectAndCompute(const Mat& img, const Mat&, std::vector<KeyPoint>& kps, Mat& desc) const {
        if (!g_svo_ref_backend.orb_detect_and_compute) svo_ref_off_path("cv::ORB::detectAndCompute without a backend");
        const int cap = nfeatures * 2 + 1024;
        kps.resize(cap);
        std::vector<uchar> d((size_t)cap * 32);
        const int n = g_svo_ref_backend.orb_detect_and_compute(img.ptr(), img.cols, img.rows, (int)img.step(), nfeatures, scale, nlevels, edge,
                                                               first, wta, score, patch, fast_th, (svo_ref_kp*)kps.data(), cap, d.data());
        if (n < 0 || n > cap) throw std::runtime_error("cv::ORB::detectAndCompute failed");
        kps.resize(n);
        desc = Mat();
        if (n > 0) { desc.create(n, 32, CV_8UC1); std::memcpy(desc.ptr(), d.data(), (size_t)n * 32); }
    }
    int nfeatures, nlevels, edge, first, wta, score, patch, fast_th;
    float scale;
};

enum { NORM_HAMMING = 6 };
class BFMatcher {
public:
    BFMatcher(int normType = 4, bool crossCheck = false) { if (normType != NORM_HAMMING || crossCheck) svo_ref_off_path("BFMatcher other than (NORM_HAMMING,false)"); }
    void match(const Mat& q, const Mat& t, std::vector<DMatch>& out) const {
        if (!g_svo_ref_backend.bf_match) svo_ref_off_path("cv::BFMatcher without a backend");
        out.resize(q.rows);
        // contiguous copies: the operands may be views
        std::vector<uchar> qb((size_t)q.rows * 32 + 32), tb((size_t)t.rows * 32 + 32);
        for (int r = 0; r < q.rows; ++r) std::memcpy(&qb[(size_t)r * 32], q.ptr(r), 32);
        for (int r = 0; r < t.rows; ++r) std::memcpy(&tb[(size_t)r * 32], t.ptr(r), 32);
        const int n = g_svo_ref_backend.bf_match(qb.data(), q.rows, tb.data(), t.rows, (svo_ref_dm*)out.data());
        if (n < 0 || n > q.rows) throw std::runtime_error("cv::BFMatcher::match failed");
        out.resize(n);
    }
};

enum { FM_7POINT = 1, FM_8POINT = 2, FM_LMEDS = 4, FM_RANSAC = 8 };
inline Mat findFundamentalMat(const Mat& p1, const Mat& p2, int method, double thr, double conf, std::vector<uchar>& mask) {
    if (!g_svo_ref_backend.find_fundamental) svo_ref_off_path("cv::findFundamentalMat without a backend");
    // p1 / p2: N x 2 CV_32F (stage4_match_consecutive.cpp:171) or N x 1 CV_32FC2 (:55, :683)
    const int n = p1.rows;
    std::vector<float> a((size_t)n * 2 + 2), b((size_t)n * 2 + 2);
    for (int r = 0; r < n; ++r) { std::memcpy(&a[(size_t)r * 2], p1.ptr(r), 8); std::memcpy(&b[(size_t)r * 2], p2.ptr(r), 8); }
    mask.assign((size_t)n + 1, 0);
    const int k = g_svo_ref_backend.find_fundamental(a.data(), b.data(), n, method, thr, conf, mask.data());
    if (k < 0) throw std::runtime_error("cv::findFundamentalMat raised an OpenCV exception");
    mask.resize(k);
    return Mat();
}
inline int countNonZero(const std::vector<uchar>& v) { int c = 0; for (uchar x : v) c += x != 0; return c; }
template <class A, class B> inline void goodFeaturesToTrack(const A&, B&, int, double, double) { svo_ref_off_path("cv::goodFeaturesToTrack (dmKLT)"); }
template <class A, class B, class C, class D> inline void calcOpticalFlowPyrLK(const Mat&, const Mat&, A&, B&, C&, D&) { svo_ref_off_path("cv::calcOpticalFlowPyrLK (ifmOpticalFlow)"); }
inline std::string format(const char* fmt, ...) { char b[1024]; va_list ap; va_start(ap, fmt); vsnprintf(b, sizeof b, fmt, ap); va_end(ap); return b; }
}  // namespace cv

// ---------------------------------------------------------------------------------------------------------------
// Eigen (the subset stage5_optimization.cpp uses)
// ---------------------------------------------------------------------------------------------------------------
namespace Eigen {
enum { ComputeThinU = 4, ComputeThinV = 16 };
class MatrixXd {
public:
    MatrixXd() : r_(0), c_(0) {}
    MatrixXd(int r, int c) : r_(r), c_(c), d_((size_t)r * c, 0.0) {}
    explicit MatrixXd(int n) : r_(n), c_(1), d_((size_t)n, 0.0) {}
    static MatrixXd Zero(int r, int c) { return MatrixXd(r, c); }
    static MatrixXd Identity(int r, int c) { MatrixXd m(r, c); for (int i = 0; i < std::min(r, c); ++i) m(i, i) = 1; return m; }
    int rows() const { return r_; }
    int cols() const { return c_; }
    double& operator()(int i, int j) { return d_[(size_t)i * c_ + j]; }
    double operator()(int i, int j) const { return d_[(size_t)i * c_ + j]; }
    double& operator()(int i) { return d_[i]; }                    // vectors
    double operator()(int i) const { return d_[i]; }
    MatrixXd transpose() const { MatrixXd t(c_, r_); for (int i = 0; i < r_; ++i) for (int j = 0; j < c_; ++j) t(j, i) = (*this)(i, j); return t; }
    MatrixXd operator*(const MatrixXd& o) const {
        MatrixXd m(r_, o.c_);
        for (int i = 0; i < r_; ++i) for (int j = 0; j < o.c_; ++j) { double s = 0; for (int k = 0; k < c_; ++k) s += (*this)(i, k) * o(k, j); m(i, j) = s; }
        return m;
    }
    MatrixXd& operator+=(const MatrixXd& o) { for (size_t i = 0; i < d_.size(); ++i) d_[i] += o.d_[i]; return *this; }
    friend MatrixXd operator*(double s, const MatrixXd& a) { MatrixXd m = a; for (double& v : m.d_) v *= s; return m; }
    friend std::ostream& operator<<(std::ostream& os, const MatrixXd& a) {
        for (int i = 0; i < a.r_; ++i) { for (int j = 0; j < a.c_; ++j) os << a(i, j) << (j + 1 < a.c_ ? " " : ""); if (i + 1 < a.r_) os << "\n"; }
        return os;
    }
    struct CommaInit {
        MatrixXd& m; int k;
        CommaInit& operator,(double v) { m.d_[k++] = v; return *this; }
    };
    CommaInit operator<<(double v) { d_[0] = v; return CommaInit{*this, 1}; }
protected:
    int r_, c_;
    std::vector<double> d_;
};
class VectorXd : public MatrixXd {
public:
    VectorXd() {}
    explicit VectorXd(int n) : MatrixXd(n, 1) {}
    VectorXd(const MatrixXd& m) : MatrixXd(m) {}
};
class Vector4d : public MatrixXd {
public:
    Vector4d() : MatrixXd(4, 1) {}
};

// Eigen::JacobiSVD for a real square matrix (no QR preconditioner is involved when rows == cols): two-sided Jacobi
// sweeps with real_2x2_jacobi_svd, singular values sorted descending, solve() = V * diag(1/sigma) * U^T * rhs over the
// rank Eigen determines with its default threshold (diagSize * epsilon relative to the largest singular value).
template <class M> class JacobiSVD {
public:
    JacobiSVD(const M& A, unsigned = 0) { compute(A); }
    const VectorXd& singularValues() const { return sv_; }
    M solve(const M& rhs) const {
        const int n = sv_.rows();
        int rank = 0;
        if (n > 0) {
            const double thr = std::max(sv_(0) * (n * std::numeric_limits<double>::epsilon()), std::numeric_limits<double>::min());
            int i = nonzero_ - 1;
            while (i >= 0 && sv_(i) < thr) --i;
            rank = i + 1;
        }
        M out(n, rhs.cols());
        for (int c = 0; c < rhs.cols(); ++c) {
            std::vector<double> tmp(rank, 0.0);
            for (int k = 0; k < rank; ++k) { double s = 0; for (int i = 0; i < n; ++i) s += U_(i, k) * rhs(i, c); tmp[k] = s / sv_(k); }
            for (int i = 0; i < n; ++i) { double s = 0; for (int k = 0; k < rank; ++k) s += V_(i, k) * tmp[k]; out(i, c) = s; }
        }
        return out;
    }
private:
    struct Rot { double c, s; };
    static bool make_jacobi(double x, double y, double z, Rot& j) {
        const double deno = 2 * std::fabs(y);
        if (deno < std::numeric_limits<double>::min()) { j.c = 1; j.s = 0; return false; }
        const double tau = (x - z) / deno, w = std::sqrt(tau * tau + 1);
        const double t = tau > 0 ? 1 / (tau + w) : 1 / (tau - w);
        const double sign_t = t > 0 ? 1 : -1, n = 1 / std::sqrt(t * t + 1);
        j.s = -sign_t * (y / std::fabs(y)) * std::fabs(t) * n;
        j.c = n;
        return true;
    }
    static void rot_rows(M& A, int p, int q, Rot j) {               // applyOnTheLeft(p, q, j)
        for (int k = 0; k < A.cols(); ++k) { const double x = A(p, k), y = A(q, k); A(p, k) = j.c * x + j.s * y; A(q, k) = -j.s * x + j.c * y; }
    }
    static void rot_cols(M& A, int p, int q, Rot j) {               // applyOnTheRight(p, q, j): the transposed rotation on the columns
        for (int k = 0; k < A.rows(); ++k) { const double x = A(k, p), y = A(k, q); A(k, p) = j.c * x - j.s * y; A(k, q) = j.s * x + j.c * y; }
    }
    void compute(const M& A) {
        const int n = A.rows();
        double scale = 0;
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) scale = std::max(scale, std::fabs(A(i, j)));
        if (!(scale > 0) && !(scale != scale)) scale = 1;            // a NaN stays a NaN, as in Eigen (maxCoeff propagates it)
        for (int i = 0; i < n && scale == scale; ++i) for (int j = 0; j < n; ++j) if (A(i, j) != A(i, j)) scale = A(i, j);
        M W(n, n);
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) W(i, j) = A(i, j) / scale;
        U_ = M::Identity(n, n);
        V_ = M::Identity(n, n);
        const double precision = 2 * std::numeric_limits<double>::epsilon();
        const double considerAsZero = std::numeric_limits<double>::min();
        double maxDiag = 0;
        for (int i = 0; i < n; ++i) maxDiag = std::max(maxDiag, std::fabs(W(i, i)));
        bool finished = false;
        int guard = 0;
        while (!finished && guard++ < 1000) {
            finished = true;
            for (int p = 1; p < n; ++p)
                for (int q = 0; q < p; ++q) {
                    const double thr = std::max(considerAsZero, precision * maxDiag);
                    if (std::fabs(W(p, q)) > thr || std::fabs(W(q, p)) > thr) {
                        finished = false;
                        // real_2x2_jacobi_svd
                        double m00 = W(p, p), m01 = W(p, q), m10 = W(q, p), m11 = W(q, q);
                        Rot rot1;
                        const double t = m00 + m11, d = m10 - m01;
                        if (std::fabs(d) < std::numeric_limits<double>::min()) { rot1.s = 0; rot1.c = 1; }
                        else { const double u = t / d, tmp = std::sqrt(1 + u * u); rot1.s = 1 / tmp; rot1.c = u / tmp; }
                        const double a00 = rot1.c * m00 + rot1.s * m10, a01 = rot1.c * m01 + rot1.s * m11;
                        const double a11 = -rot1.s * m01 + rot1.c * m11;
                        Rot jr;
                        make_jacobi(a00, a01, a11, jr);
                        Rot jl;                                      // rot1 * jr^T
                        jl.c = rot1.c * jr.c + rot1.s * jr.s;
                        jl.s = rot1.s * jr.c - rot1.c * jr.s;
                        rot_rows(W, p, q, jl);
                        rot_cols(U_, p, q, Rot{jl.c, -jl.s});
                        rot_cols(W, p, q, jr);
                        rot_cols(V_, p, q, jr);
                        maxDiag = std::max(maxDiag, std::max(std::fabs(W(p, p)), std::fabs(W(q, q))));
                    }
                }
        }
        sv_ = VectorXd(n);
        for (int i = 0; i < n; ++i) {
            const double a = std::fabs(W(i, i));
            sv_(i) = a;
            if (W(i, i) < 0) for (int k = 0; k < n; ++k) U_(k, i) = -U_(k, i);
        }
        for (int i = 0; i < n; ++i) sv_(i) *= scale;
        nonzero_ = n;
        for (int i = 0; i < n; ++i) {
            int pos = i;
            double best = sv_(i);
            for (int k = i + 1; k < n; ++k) if (sv_(k) > best) { best = sv_(k); pos = k; }
            if (best == 0) { nonzero_ = i; break; }
            if (pos != i) {
                std::swap(sv_(i), sv_(pos));
                for (int k = 0; k < n; ++k) { std::swap(U_(k, i), U_(k, pos)); std::swap(V_(k, i), V_(k, pos)); }
            }
        }
    }
    M U_, V_;
    VectorXd sv_;
    int nonzero_;
};
}  // namespace Eigen

// ---------------------------------------------------------------------------------------------------------------
// stlplus, MRPT 1.x
// ---------------------------------------------------------------------------------------------------------------
namespace stlplus {
template <class T> class smart_ptr {
public:
    smart_ptr() {}
    explicit smart_ptr(T* p) : p_(p) {}
    bool present() const { return (bool)p_; }
    explicit operator bool() const { return (bool)p_; }
    T* operator->() const { return p_.get(); }
    T& operator*() const { return *p_; }
    T* pointer() const { return p_.get(); }
private:
    std::shared_ptr<T> p_;
};
}  // namespace stlplus

#define MRPT_TODO(x)
#define ASSERT_(c) { if (!(c)) throw std::logic_error(std::string("Assert condition failed: ") + #c); }
#define ASSERTMSG_(c, msg) { if (!(c)) throw std::logic_error(std::string(msg)); }
#define ASSERTDEB_(c)
#define ASSERTDEBMSG_(c, msg)
#define THROW_EXCEPTION(msg) { throw std::logic_error(std::string(msg)); }
#define RAD2DEG(x) ((x) * 180.0 / M_PI)
#define INVALID_TIMESTAMP (0)

namespace mrpt {
typedef std::vector<size_t> vector_size_t;
inline std::string format(const char* fmt, ...) { char b[2048]; va_list ap; va_start(ap, fmt); vsnprintf(b, sizeof b, fmt, ap); va_end(ap); return b; }

namespace system {
typedef uint64_t TTimeStamp;
inline bool directoryExists(const std::string& p) { struct stat s; return stat(p.c_str(), &s) == 0 && S_ISDIR(s.st_mode); }
inline bool fileExists(const std::string& p) { struct stat s; return stat(p.c_str(), &s) == 0; }
inline bool createDirectory(const std::string& p) { return mkdir(p.c_str(), 0777) == 0; }
inline void pause() {}
namespace os {
inline FILE* fopen(const char* n, const char* m) { return ::fopen(n, m); }
inline FILE* fopen(const std::string& n, const char* m) { return ::fopen(n.c_str(), m); }
inline int fprintf(FILE* f, const char* fmt, ...) { if (!f) return 0; va_list ap; va_start(ap, fmt); const int r = vfprintf(f, fmt, ap); va_end(ap); return r; }
inline void fclose(FILE* f) { if (f) ::fclose(f); }
}  // namespace os
struct TThreadHandle { bool isClear() const { return true; } };
template <class C> inline TThreadHandle createThreadFromObjectMethod(C*, void (C::*)()) { svo_ref_off_path("GUI thread"); }
inline void joinThread(const TThreadHandle&) {}
}  // namespace system

namespace math {
inline bool isNaN(double v) { return v != v; }
inline bool isFinite(double v) { return std::isfinite(v); }
struct TPoint3D { double x, y, z; TPoint3D() : x(0), y(0), z(0) {} TPoint3D(double a, double b, double c) : x(a), y(b), z(c) {} };
class CMatrixB {                                                   // CMatrixBool: rows x cols of bool
public:
    CMatrixB(size_t r = 0, size_t c = 0) : r_(r), c_(c), d_(r * c, 0) {}
    void fillAll(bool v) { std::fill(d_.begin(), d_.end(), (char)v); }
    bool operator()(size_t r, size_t c) const { return d_[r * c_ + c] != 0; }
    void set_unsafe(size_t r, size_t c, bool v) { d_[r * c_ + c] = (char)v; }
private:
    size_t r_, c_;
    std::vector<char> d_;
};
}  // namespace math

namespace poses {
using mrpt::math::TPoint3D;
struct CPose3DQuat { double m[7]; CPose3DQuat() { for (double& v : m) v = 0; m[3] = 1; } double& operator[](int i) { return m[i]; } double operator[](int i) const { return m[i]; } };
class CPose3D;
// rotation vector + translation.  getInverse(): (-w, -R(w)^T t)  [MRPT 1.x CPose3DRotVec::inverse(), restated]
class CPose3DRotVec {
public:
    double m_coords[3], m_rotvec[3];
    CPose3DRotVec() { for (int i = 0; i < 3; ++i) m_coords[i] = m_rotvec[i] = 0; }
    CPose3DRotVec(double w1, double w2, double w3, double x, double y, double z) { m_rotvec[0] = w1; m_rotvec[1] = w2; m_rotvec[2] = w3; m_coords[0] = x; m_coords[1] = y; m_coords[2] = z; }
    explicit CPose3DRotVec(const CPose3D& p);
    static void rodrigues(const double* w, double R[9]) {
        const double th = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
        double a, b;                                               // R = I + a [w]x + b [w]x^2
        if (th < 1e-12) { a = 1; b = 0.5; } else { a = std::sin(th) / th; b = (1 - std::cos(th)) / (th * th); }
        const double x = w[0], y = w[1], z = w[2];
        R[0] = 1 - b * (y * y + z * z); R[1] = -a * z + b * x * y;    R[2] = a * y + b * x * z;
        R[3] = a * z + b * x * y;      R[4] = 1 - b * (x * x + z * z); R[5] = -a * x + b * y * z;
        R[6] = -a * y + b * x * z;     R[7] = a * x + b * y * z;      R[8] = 1 - b * (x * x + y * y);
    }
    CPose3DRotVec getInverse() const {
        double R[9];
        rodrigues(m_rotvec, R);
        CPose3DRotVec o;
        for (int i = 0; i < 3; ++i) {
            o.m_rotvec[i] = -m_rotvec[i];
            o.m_coords[i] = -(R[0 + i] * m_coords[0] + R[3 + i] * m_coords[1] + R[6 + i] * m_coords[2]);
        }
        return o;
    }
};
class CPose3D {
public:
    double R[9], t[3];
    CPose3D() { for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0); t[0] = t[1] = t[2] = 0; }
    explicit CPose3D(const CPose3DRotVec& rv) { CPose3DRotVec::rodrigues(rv.m_rotvec, R); for (int i = 0; i < 3; ++i) t[i] = rv.m_coords[i]; }
    double x() const { return t[0]; }
    double y() const { return t[1]; }
    double z() const { return t[2]; }
    // MRPT: R = Rz(yaw) Ry(pitch) Rx(roll)
    double pitch() const { return std::atan2(-R[6], std::hypot(R[0], R[3])); }
    double yaw() const { return std::atan2(R[3], R[0]); }
    double roll() const { return std::atan2(R[7], R[8]); }
    void inverse() {
        double Rt[9], nt[3];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rt[3 * i + j] = R[3 * j + i];
        for (int i = 0; i < 3; ++i) nt[i] = -(Rt[3 * i] * t[0] + Rt[3 * i + 1] * t[1] + Rt[3 * i + 2] * t[2]);
        std::memcpy(R, Rt, sizeof R);
        std::memcpy(t, nt, sizeof t);
    }
    friend std::ostream& operator<<(std::ostream& os, const CPose3D& p) {
        return os << "(" << p.x() << "," << p.y() << "," << p.z() << "," << RAD2DEG(p.yaw()) << "deg," << RAD2DEG(p.pitch()) << "deg," << RAD2DEG(p.roll()) << "deg)";
    }
};
inline CPose3DRotVec::CPose3DRotVec(const CPose3D& p) {           // log map of R
    const double c = std::max(-1.0, std::min(1.0, (p.R[0] + p.R[4] + p.R[8] - 1) * 0.5)), th = std::acos(c);
    double k = th < 1e-12 ? 0.5 : th / (2 * std::sin(th));
    m_rotvec[0] = k * (p.R[7] - p.R[5]); m_rotvec[1] = k * (p.R[2] - p.R[6]); m_rotvec[2] = k * (p.R[3] - p.R[1]);
    for (int i = 0; i < 3; ++i) m_coords[i] = p.t[i];
}
}  // namespace poses

namespace utils {
struct TPixelCoord { int x, y; TPixelCoord() : x(0), y(0) {} TPixelCoord(int a, int b) : x(a), y(b) {} };
struct TPixelCoordf { float x, y; TPixelCoordf() : x(0), y(0) {} TPixelCoordf(float a, float b) : x(a), y(b) {} };
inline std::ostream& operator<<(std::ostream& os, const TPixelCoordf& p) { return os << "(" << p.x << "," << p.y << ")"; }
typedef TPixelCoord TImageSize;
enum TConstructorFlags_CImage { UNINITIALIZED_IMAGE = 0, FAST_REF_OR_CONVERT_TO_GRAY = 1 };

class CImage {                                                      // 8-bit gray only (stage 1's colour conversion is off this path)
public:
    CImage() : w_(0), h_(0) {}
    CImage(TConstructorFlags_CImage) : w_(0), h_(0) {}
    CImage(const CImage& o, TConstructorFlags_CImage) : w_(o.w_), h_(o.h_), d_(o.d_) {}   // already gray: a fast reference
    CImage(const CImage& o) : w_(o.w_), h_(o.h_), d_(o.d_ ? std::make_shared<std::vector<uint8_t>>(*o.d_) : nullptr) {}
    CImage& operator=(const CImage& o) { w_ = o.w_; h_ = o.h_; d_ = o.d_ ? std::make_shared<std::vector<uint8_t>>(*o.d_) : nullptr; return *this; }
    void loadFromMemoryBuffer(int w, int h, const uint8_t* p, int stride) {
        w_ = w; h_ = h;
        d_ = std::make_shared<std::vector<uint8_t>>((size_t)w * h);
        for (int r = 0; r < h; ++r) std::memcpy(d_->data() + (size_t)r * w, p + (size_t)r * stride, w);
    }
    size_t getWidth() const { return w_; }
    size_t getHeight() const { return h_; }
    TImageSize getSize() const { return TImageSize(w_, h_); }
    size_t getRowStride() const { return w_; }
    bool isColor() const { return false; }
    unsigned char* get_unsafe(unsigned col, unsigned row, unsigned = 0) const { return d_->data() + (size_t)row * w_ + col; }
    template <class T> const T* getAs() const { ipl_.width = w_; ipl_.height = h_; ipl_.widthStep = w_; ipl_.nChannels = 1; ipl_.imageData = d_ ? (char*)d_->data() : nullptr; return &ipl_; }
    void swap(CImage& o) { std::swap(w_, o.w_); std::swap(h_, o.h_); d_.swap(o.d_); }
    void copyFastFrom(CImage& o) { w_ = o.w_; h_ = o.h_; d_ = o.d_; o.d_.reset(); o.w_ = o.h_ = 0; }   // move semantics (P:110)
    void copyFromForceLoad(const CImage& o) { *this = o; }
    bool saveToFile(const std::string& name, int = 95) const {
        FILE* f = ::fopen(name.c_str(), "wb");                    // PGM under whatever name the caller chose
        if (!f) return false;
        std::fprintf(f, "P5\n%d %d\n255\n", w_, h_);
        if (d_) fwrite(d_->data(), 1, d_->size(), f);
        ::fclose(f);
        return true;
    }
    float KLT_response(unsigned, unsigned, unsigned) const { svo_ref_off_path("CImage::KLT_response (dmFASTER)"); }
private:
    int w_, h_;
    std::shared_ptr<std::vector<uint8_t>> d_;
    mutable IplImage ipl_;
};

class CConfigFileBase {
public:
    virtual ~CConfigFileBase() {}
    int read_int(const std::string& s, const std::string& k, int def, bool = false) const { std::string v; return get(s, k, v) ? (int)std::strtol(v.c_str(), nullptr, 0) : def; }
    double read_double(const std::string& s, const std::string& k, double def, bool = false) const { std::string v; return get(s, k, v) ? std::strtod(v.c_str(), nullptr) : def; }
    bool read_bool(const std::string& s, const std::string& k, bool def, bool = false) const {
        std::string v;
        if (!get(s, k, v)) return def;
        std::string l;
        for (char c : v) l += (char)std::tolower(c);
        return l == "1" || l == "true" || l == "yes";
    }
    std::string read_string(const std::string& s, const std::string& k, const std::string& def, bool = false) const { std::string v; return get(s, k, v) ? v : def; }
protected:
    bool get(const std::string& s, const std::string& k, std::string& v) const {
        auto i = kv_.find(lower(s) + "\n" + lower(k));
        if (i == kv_.end()) return false;
        v = i->second;
        return true;
    }
    static std::string lower(const std::string& a) { std::string l; for (char c : a) l += (char)std::tolower(c); return l; }
    static std::string trim(const std::string& a) { size_t b = a.find_first_not_of(" \t\r\n"), e = a.find_last_not_of(" \t\r\n"); return b == std::string::npos ? "" : a.substr(b, e - b + 1); }
    std::map<std::string, std::string> kv_;
};
class CConfigFile : public CConfigFileBase {
public:
    explicit CConfigFile(const std::string& file) {
        std::ifstream f(file.c_str());
        std::string line, sec;
        while (std::getline(f, line)) {
            const size_t c = line.find_first_of(";#");
            if (c != std::string::npos && (c == 0 || line[c] == ';' || line.find("//") == c)) line = line.substr(0, c);
            const size_t cc = line.find("//");
            if (cc != std::string::npos) line = line.substr(0, cc);
            line = trim(line);
            if (line.empty()) continue;
            if (line[0] == '[') { sec = trim(line.substr(1, line.find(']') - 1)); continue; }
            const size_t eq = line.find('=');
            if (eq == std::string::npos) continue;
            kv_[lower(sec) + "\n" + lower(trim(line.substr(0, eq)))] = trim(line.substr(eq + 1));
        }
    }
};

class CTimeLogger {
public:
    CTimeLogger(bool = true) {}
    void enter(const char*) {}
    double leave(const char*) { return 0; }
    void enable(bool = true) {}
};

struct TCamera {
    double fx_, fy_, cx_, cy_;
    uint32_t ncols, nrows;
    TCamera() : fx_(0), fy_(0), cx_(0), cy_(0), ncols(640), nrows(480) {}
    double fx() const { return fx_; }
    double fy() const { return fy_; }
    double cx() const { return cx_; }
    double cy() const { return cy_; }
    void setIntrinsicParamsFromValues(double fx, double fy, double cx, double cy) { fx_ = fx; fy_ = fy; cx_ = cx; cy_ = cy; }
};
struct TStereoCamera { TCamera leftCamera, rightCamera; mrpt::poses::CPose3DQuat rightCameraPose; };
}  // namespace utils

namespace synch {
struct CCriticalSection { void enter() {} void leave() {} };
template <class T> struct CThreadSafeVariable {};
}  // namespace synch
namespace gui { struct CDisplayWindow3D {}; struct CDisplayWindow {}; }
namespace opengl { struct CSetOfObjects {}; }

namespace obs {
class CObservationStereoImages {
public:
    mrpt::system::TTimeStamp timestamp;
    mrpt::utils::CImage imageLeft, imageRight;
    bool rectified;
    CObservationStereoImages() : timestamp(0), rectified(true) {}
    bool areImagesRectified() const { return rectified; }
};
class CObservationStereoImagesPtr {
public:
    CObservationStereoImagesPtr() {}
    explicit CObservationStereoImagesPtr(CObservationStereoImages* p) : p_(p) {}
    bool present() const { return (bool)p_; }
    CObservationStereoImages* pointer() const { return p_.get(); }
    CObservationStereoImages* operator->() const { return p_.get(); }
private:
    std::shared_ptr<CObservationStereoImages> p_;
};
}  // namespace obs

namespace vision {
struct TSimpleFeature { mrpt::utils::TPixelCoord pt; float response; uint8_t octave; uint64_t ID; TSimpleFeature() : response(0), octave(0), ID(0) {} };
class TSimpleFeatureList {
public:
    typedef std::vector<TSimpleFeature>::iterator iterator;
    typedef std::vector<TSimpleFeature>::const_iterator const_iterator;
    size_t size() const { return v_.size(); }
    TSimpleFeature& operator[](size_t i) { return v_[i]; }
    const TSimpleFeature& operator[](size_t i) const { return v_[i]; }
    iterator begin() { return v_.begin(); }
    iterator end() { return v_.end(); }
    const_iterator begin() const { return v_.begin(); }
    const_iterator end() const { return v_.end(); }
private:
    std::vector<TSimpleFeature> v_;
};
// MRPT 1.x mrpt/vision/TSimpleFeature.h: sorts indices by descending response
template <typename FEATURE_LIST> struct KeypointResponseSorter : public std::binary_function<size_t, size_t, bool> {
    const FEATURE_LIST& m_data;
    KeypointResponseSorter(const FEATURE_LIST& data) : m_data(data) {}
    bool operator()(size_t k1, size_t k2) const { return (m_data[k1].response > m_data[k2].response); }
};
struct CFeatureExtraction {
    template <class... A> static void detectFeatures_SSE2_FASTER12(A&&...) { svo_ref_off_path("CFeatureExtraction::detectFeatures_SSE2_FASTER12 (dmFASTER)"); }
};
class CStereoRectifyMap {
public:
    bool isSet() const { return false; }
    template <class T> void setFromCamParams(const T&) { svo_ref_off_path("CStereoRectifyMap (unrectified input)"); }
    template <class... A> void rectify(A&&...) { svo_ref_off_path("CStereoRectifyMap::rectify"); }
};
// CImagePyramid::buildPyramidFast(img, nOctaves, smooth_halves = true): octave 0 takes the image over (move), octave k is
// CImage::scaleHalfSmooth of octave k-1 -- MRPT 1.x: the SSE2 2x2 average (rows with _mm_avg_epu8, then column pairs with
// _mm_avg_epu16, both rounding half up) for 8-bit gray images whose width is a multiple of 16, cvPyrDown otherwise.
// Restated, UNPINNED (MRPT's source is not in this image); the same statement as oracle/oracle.py pyr_half.
class CImagePyramid {
public:
    std::vector<mrpt::utils::CImage> images;
    void buildPyramidFast(mrpt::utils::CImage& img, size_t nOctaves, bool = true, bool = false) {
        images.resize(nOctaves);
        images[0].copyFastFrom(img);
        for (size_t o = 1; o < nOctaves; ++o) {
            const mrpt::utils::CImage& s = images[o - 1];
            const int w = (int)s.getWidth(), h = (int)s.getHeight();
            std::vector<uint8_t> out;
            int ow, oh;
            if (w % 16 == 0) {
                ow = w >> 1; oh = h >> 1;
                out.resize((size_t)ow * oh);
                for (int y = 0; y < oh; ++y) {
                    const uint8_t *r0 = s.get_unsafe(0, 2 * y), *r1 = s.get_unsafe(0, 2 * y + 1);
                    for (int x = 0; x < ow; ++x) {
                        const int a = (r0[2 * x] + r1[2 * x] + 1) >> 1, b = (r0[2 * x + 1] + r1[2 * x + 1] + 1) >> 1;
                        out[(size_t)y * ow + x] = (uint8_t)((a + b + 1) >> 1);
                    }
                }
            } else {
                if (!g_svo_ref_backend.pyr_down) svo_ref_off_path("cvPyrDown without a backend");
                ow = w >> 1; oh = h >> 1;                  // scaleHalfSmooth allocates (w>>1) x (h>>1) and lets cvPyrDown fill it
                out.resize((size_t)ow * oh);
                g_svo_ref_backend.pyr_down(s.get_unsafe(0, 0), w, h, w, out.data(), ow, oh);
            }
            images[o].loadFromMemoryBuffer(ow, oh, out.data(), ow);
        }
    }
};
}  // namespace vision
namespace slam { using mrpt::obs::CObservationStereoImages; using mrpt::obs::CObservationStereoImagesPtr; }
}  // namespace mrpt
