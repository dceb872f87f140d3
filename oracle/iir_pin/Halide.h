/*
 * oracle/iir_pin/Halide.h -- TEST INFRASTRUCTURE.  The few Halide names that
 * /root/reference/lib/iir_coeff.cpp touches, so that the reference's coefficient design
 * (lib/iir_coeff.cpp:38-263, plain float/double arithmetic) compiles here UNCHANGED without the
 * Halide fork (unbuildable: needs LLVM 3.4).  Only the three Expr overloads
 * (lib/iir_coeff.cpp:179-192) use these names; they are never called by the pin.
 */
#pragma once
#include <cassert>
#include <cmath>
#include <vector>
namespace Halide {
struct Type {};
template <typename T> inline Type type_of() { return Type(); }
struct Expr {
    Expr() {}
    Expr(float) {}
    Expr(double) {}
    Expr(int) {}
};
inline Expr operator+(Expr, Expr) { return Expr(); }
inline Expr operator-(Expr, Expr) { return Expr(); }
inline Expr operator*(Expr, Expr) { return Expr(); }
inline Expr operator/(Expr, Expr) { return Expr(); }
inline Expr operator-(Expr) { return Expr(); }
inline Expr fast_exp(Expr) { return Expr(); }
inline Expr erf(Expr) { return Expr(); }
namespace Internal { struct Cast { static Expr make(Type, Expr e) { return e; } }; }
template <typename T> struct Image {
    int w, h; std::vector<T> d;
    Image(int w_, int h_) : w(w_), h(h_), d((size_t)w_ * h_) {}
    int width() const { return w; } int height() const { return h; }
    T& operator()(int x, int y) { return d[(size_t)y * w + x]; }
};
}
