/*
 * tests/host/opencv_shim/opencv2/opencv.hpp -- TEST INFRASTRUCTURE ONLY.
 * Declarations (no definitions) of the part of the OpenCV 4 C++ API that the reference's renderer uses
 * (/root/reference/src/visualize.cpp:73-461), with OpenCV's own signatures, so that the UNCHANGED renderer can be
 * type-checked (g++ -fsyntax-only) against this repo's include/visualize.hpp in an image without OpenCV:
 * tests/test_cpp_surface.py::test_reference_renderer_type_checks_against_our_header.  Nothing links against this.
 */
#pragma once
#include <array>
#include <string>
#include <utility>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_8UC3 16

namespace cv {

typedef std::string String;

template <typename T> class Scalar_ {
public:
    Scalar_();
    Scalar_(T v0, T v1, T v2 = 0, T v3 = 0);
    Scalar_(T v0);
    static Scalar_<T> all(T v0);
    T val[4];
};
typedef Scalar_<double> Scalar;

template <typename T> class Point_ {
public:
    Point_();
    Point_(T x, T y);
    T x, y;
};
typedef Point_<int> Point;

template <typename T> class Size_ {
public:
    Size_();
    Size_(T width, T height);
    T width, height;
};
typedef Size_<int> Size;

class Mat;
class _InputArray {
public:
    _InputArray(const Mat& m);
    _InputArray(const std::vector<Mat>& v);
};
class _OutputArray : public _InputArray {
public:
    _OutputArray(Mat& m);
};
typedef const _InputArray& InputArray;
typedef InputArray InputArrayOfArrays;
typedef const _OutputArray& OutputArray;
typedef const _OutputArray& InputOutputArray;

class Mat {
public:
    Mat();
    Mat(int rows, int cols, int type);
    Mat(int rows, int cols, int type, const Scalar& s);
    Mat(const Mat& m);
    ~Mat();
    Mat& operator=(const Mat& m);
    template <typename T> T& at(int row, int col);
    template <typename T> const T& at(int row, int col) const;
    void convertTo(OutputArray m, int rtype, double alpha = 1, double beta = 0) const;
    bool empty() const;
    Mat clone() const;
    int rows, cols;
};

enum ColormapTypes { COLORMAP_AUTUMN = 0, COLORMAP_BONE = 1, COLORMAP_JET = 2 };
enum HersheyFonts { FONT_HERSHEY_SIMPLEX = 0, FONT_HERSHEY_PLAIN = 1 };
enum LineTypes { FILLED = -1, LINE_4 = 4, LINE_8 = 8, LINE_AA = 16 };
enum BorderTypes { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1 };

void hconcat(InputArrayOfArrays src, OutputArray dst);
void hconcat(InputArray src1, InputArray src2, OutputArray dst);
void vconcat(InputArrayOfArrays src, OutputArray dst);
void vconcat(InputArray src1, InputArray src2, OutputArray dst);
void applyColorMap(InputArray src, OutputArray dst, int colormap);
void flip(InputArray src, OutputArray dst, int flipCode);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType, const Scalar& value = Scalar());
void putText(InputOutputArray img, const String& text, Point org, int fontFace, double fontScale, Scalar color, int thickness = 1,
             int lineType = LINE_8, bool bottomLeftOrigin = false);
void line(InputOutputArray img, Point pt1, Point pt2, const Scalar& color, int thickness = 1, int lineType = LINE_8, int shift = 0);
void rectangle(InputOutputArray img, Point pt1, Point pt2, const Scalar& color, int thickness = 1, int lineType = LINE_8, int shift = 0);
bool imwrite(const String& filename, InputArray img, const std::vector<int>& params = std::vector<int>());

class VideoWriter {
public:
    VideoWriter();
    VideoWriter(const String& filename, int fourcc, double fps, Size frameSize, bool isColor = true);
    ~VideoWriter();
    bool open(const String& filename, int fourcc, double fps, Size frameSize, bool isColor = true);
    bool isOpened() const;
    void release();
    void write(InputArray image);
    static int fourcc(char c1, char c2, char c3, char c4);
};

} // namespace cv
