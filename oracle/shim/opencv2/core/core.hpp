// No-op stand-in for the OpenCV 2.4 C++ API surface that fish_mod.cpp touches
// (fish_mod.cpp:14-18,86,89,110-111,144,147,212-222,527,585,634).  TEST INFRASTRUCTURE ONLY:
// it lets the unmodified reference translation unit compile headless so that its own
// functions can serve as the parity oracle.  Nothing here draws anything.
#pragma once
#include <sstream>   // the reference uses std::stringstream (fish_mod.cpp:614) without including it
#include <string>

#define CV_VERSION "shim-0"
#define CV_8UC3 16
#define CV_AA 16
#define CV_FOURCC(a, b, c, d) 0

namespace cv {

struct Vec3b {
    unsigned char v[3];
    Vec3b() : v{0, 0, 0} {}
    Vec3b(int a, int b, int c) : v{(unsigned char)a, (unsigned char)b, (unsigned char)c} {}
};

struct Scalar {
    double v[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : v{a, b, c, d} {}
};

struct Point {
    int x, y;
    Point() : x(0), y(0) {}
    Point(double x_, double y_) : x((int)x_), y((int)y_) {}
};

struct Size {
    int width, height;
    Size(int w = 0, int h = 0) : width(w), height(h) {}
};

struct Mat {
    static Mat zeros(int, int, int) { return Mat(); }
    template <class T> T& at(int, int) {
        static thread_local T sink;
        return sink;
    }
};

struct VideoWriter {
    bool open(const std::string&, int, double, Size, bool = true) { return false; }
    bool isOpened() const { return false; }
    void write(const Mat&) {}
};

inline void circle(Mat&, Point, int, const Scalar&, int = 1, int = 8, int = 0) {}
inline void arrowedLine(Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0, double = 0.1) {}
inline void rectangle(Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0) {}
inline void addWeighted(const Mat&, double, const Mat&, double, double, Mat&) {}

}  // namespace cv
