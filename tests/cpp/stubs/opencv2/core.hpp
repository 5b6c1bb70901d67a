// Minimal stand-in for the OpenCV core declarations the reference's headers and the adapters use.  Declarations only (syntax check).
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <iostream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#define CV_8U 0
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_8UC4 24
#define CV_32F 5
#define CV_32FC1 5
#define CV_64F 6
#define CV_64FC1 6
#define CV_VERSION_MAJOR 4
#define CV_MAJOR_VERSION 4
typedef unsigned char uchar;
int cvRound(double); int cvFloor(double); int cvCeil(double);
namespace cv {
template <class T> struct Point_ { T x, y; Point_(); Point_(T, T); template <class U> Point_(const Point_<U>&); };
template <class T> Point_<T> operator+(const Point_<T>&, const Point_<T>&);
template <class T> Point_<T> operator-(const Point_<T>&, const Point_<T>&);
template <class T> Point_<T> operator*(const Point_<T>&, double);
template <class T> Point_<T>& operator*=(Point_<T>&, double);
template <class T> bool operator==(const Point_<T>&, const Point_<T>&);
using Point2f = Point_<float>; using Point2i = Point_<int>; using Point = Point2i; using Point2d = Point_<double>;
template <class T> struct Point3_ { T x, y, z; Point3_(); Point3_(T, T, T); };
using Point3f = Point3_<float>; using Point3d = Point3_<double>;
template <class T> struct Size_ { T width, height; Size_(); Size_(T, T); };
using Size = Size_<int>;
template <class T> struct Rect_ { T x, y, width, height; Rect_(); Rect_(T, T, T, T); };
using Rect = Rect_<int>;
template <class T, int N> struct Vec { T val[N]; T& operator[](int); const T& operator[](int) const; };
using Vec3b = Vec<uchar, 3>; using Vec3f = Vec<float, 3>;
struct Scalar { Scalar(); Scalar(double); Scalar(double, double, double, double = 0); };
struct Range { int start, end; Range(); Range(int, int); static Range all(); };
struct KeyPoint {
    Point2f pt; float size; float angle; float response; int octave; int class_id;
    KeyPoint(); KeyPoint(Point2f, float, float = -1, float = 0, int = 0, int = -1); KeyPoint(float, float, float, float = -1, float = 0, int = 0, int = -1);
};
struct MatSize { int operator[](int) const; };
class _InputArray; class _OutputArray;
class Mat {
public:
    Mat(); Mat(int, int, int); Mat(int, int, int, const Scalar&); Mat(int, int, int, void*, size_t = 0); Mat(Size, int); Mat(const Mat&);
    template <class T> explicit Mat(const std::vector<T>&);
    Mat& operator=(const Mat&);
    int rows, cols; uchar* data; MatSize size;
    struct Step { operator size_t() const; size_t operator[](int) const; } step;
    bool empty() const; int type() const; int channels() const; int depth() const; size_t total() const; size_t elemSize() const; bool isContinuous() const;
    Mat clone() const; void copyTo(const _OutputArray&) const; void copyTo(Mat&) const; void convertTo(Mat&, int, double = 1, double = 0) const;
    void create(int, int, int); void release(); Mat row(int) const; Mat col(int) const; Mat rowRange(int, int) const; Mat colRange(int, int) const;
    Mat operator()(const Rect&) const; Mat operator()(Range, Range) const; Mat t() const; Mat inv() const; Mat reshape(int, int = 0) const;
    template <class T> T& at(int, int); template <class T> const T& at(int, int) const; template <class T> T& at(int); template <class T> const T& at(int) const;
    template <class T> T* ptr(int = 0); template <class T> const T* ptr(int = 0) const; uchar* ptr(int = 0); const uchar* ptr(int = 0) const;
    static Mat zeros(int, int, int); static Mat ones(int, int, int); static Mat eye(int, int, int);
    void push_back(const Mat&); Mat& setTo(const Scalar&);
};
template <class T> class Mat_ : public Mat { public: Mat_(); Mat_(int, int); T& operator()(int, int); struct Init { template <class U> Init& operator,(const U&); operator Mat_() const; }; template <class U> Init operator<<(const U&); };
Mat operator*(const Mat&, const Mat&); Mat operator+(const Mat&, const Mat&); Mat operator-(const Mat&, const Mat&); Mat operator*(const Mat&, double); Mat operator*(double, const Mat&);
class _InputArray {
public:
    _InputArray(); _InputArray(const Mat&); template <class T> _InputArray(const std::vector<T>&);
    Mat getMat(int = -1) const; bool empty() const; int type() const;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray(); _OutputArray(Mat&); template <class T> _OutputArray(std::vector<T>&);
    void create(int, int, int) const; void create(Size, int) const; void release() const; Mat& getMatRef() const;
};
using InputArray = const _InputArray&; using OutputArray = const _OutputArray&; using InputOutputArray = const _OutputArray&;
InputArray noArray();
double norm(InputArray, int = 4); double norm(InputArray, InputArray, int = 4);
enum { NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6 };
void vconcat(InputArray, InputArray, OutputArray); void hconcat(InputArray, InputArray, OutputArray);
std::string format(const char*, ...);
template <class T> T saturate_cast(double);
}  // namespace cv
