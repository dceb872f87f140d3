/*
 * oracle/iir_pin/wrap.cpp -- TEST INFRASTRUCTURE.  C entry points over the coefficient-design
 * functions declared by iir_coeff.h, so that Python can call two builds side by side:
 *   -DPFX=ref_  linked with /root/reference/lib/iir_coeff.cpp (the reference, compiled unchanged)
 *   -DPFX=own_  linked with recfilter_b200/host/iir_coeff.cpp (this repo)
 * oracle/pin_iir.py compares them and commits the reference's numbers as tests/golden/iir_coeff_reference.json.
 */
#include <iir_coeff.h>
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define EXPORT extern "C" __attribute__((visibility("default")))

EXPORT int CAT(PFX, gaussian_weights)(float sigma, int order, float* out)
{
    const std::vector<float> w = gaussian_weights(sigma, order);
    for (size_t i = 0; i < w.size(); ++i) out[i] = w[i];
    return (int)w.size();
}
EXPORT int CAT(PFX, integral_image_coeff)(int n, float* out)
{
    const std::vector<float> w = integral_image_coeff(n);
    for (size_t i = 0; i < w.size(); ++i) out[i] = w[i];
    return (int)w.size();
}
EXPORT int CAT(PFX, overlap_feedback_coeff)(const float* a, int na, const float* b, int nb, float* out)
{
    const std::vector<float> w = overlap_feedback_coeff(std::vector<float>(a, a + na), std::vector<float>(b, b + nb));
    for (size_t i = 0; i < w.size(); ++i) out[i] = w[i];
    return (int)w.size();
}
EXPORT int   CAT(PFX, gaussian_box_filter)(int k, float sigma) { return gaussian_box_filter(k, sigma); }
EXPORT float CAT(PFX, gaussian)(float x, float mu, float sigma) { return gaussian(x, mu, sigma); }
EXPORT float CAT(PFX, gaussDerivative)(float x, float mu, float sigma) { return gaussDerivative(x, mu, sigma); }
EXPORT float CAT(PFX, gaussIntegral)(float x, float mu, float sigma) { return gaussIntegral(x, mu, sigma); }
