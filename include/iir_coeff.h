/*
 * iir_coeff.h -- IIR coefficient design helpers with the reference's names and argument
 * meaning (/root/reference/lib/iir_coeff.h): every function returns {b0, a1..ar} as passed to
 * RecFilter::add_filter, i.e. feedback terms are ADDED (lib/iir_coeff.cpp:172-174).
 */
#ifndef _IIR_FILTER_COEFFICIENTS_H_
#define _IIR_FILTER_COEFFICIENTS_H_

#include <cmath>
#include <complex>
#include <iostream>
#include <vector>
#include <Halide.h>

/** feedback of the cascade of two all-pole filters (polynomial product), lib/iir_coeff.cpp:236-263 */
std::vector<float> overlap_feedback_coeff(std::vector<float> a, std::vector<float> b);

/** n-fold integral image: {1, 1}, {1, 2, -1}, ... lib/iir_coeff.cpp:222-234 */
std::vector<float> integral_image_coeff(int iterations);

/** Gaussian, its derivative and integral at x (lib/iir_coeff.cpp:193-203) */
float gaussian(float x, float mu, float sigma);
float gaussDerivative(float x, float mu, float sigma);
float gaussIntegral(float x, float mu, float sigma);

/** order 1..3 recursive Gaussian of van Vliet, Young and Verbeek (lib/iir_coeff.cpp:162-177) */
std::vector<float> gaussian_weights(float sigma, int order);

/** box width so that `iterations` box filters approximate a Gaussian (lib/iir_coeff.cpp:205-220) */
int gaussian_box_filter(int iterations, float sigma);

/** brute-force O(N^2) normalised Gaussian blur, for checking small images (lib/iir_coeff.h:79-100) */
template <typename T>
Halide::Image<T> reference_gaussian(Halide::Image<T> in, T sigma)
{
    const int W = in.width(), H = in.height();
    Halide::Image<T> ref(W, H);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            float acc = 0.0f, wsum = 0.0f;
            for (int j = 0; j < H; ++j)
                for (int i = 0; i < W; ++i) {
                    const float d2 = float((x - i) * (x - i) + (y - j) * (y - j));
                    const float g = gaussian(std::sqrt(d2), 0.0f, sigma);
                    acc += g * in(i, j);
                    wsum += g;
                }
            ref(x, y) = acc / wsum;
        }
    return ref;
}

#endif // _IIR_FILTER_COEFFICIENTS_H_
