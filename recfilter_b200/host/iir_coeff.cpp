/*
 * iir_coeff.cpp -- coefficient design (host).  Restates /root/reference/lib/iir_coeff.cpp:
 * the recursive Gaussian of van Vliet, Young and Verbeek ("Recursive Gaussian derivative
 * filters", 1998) with the pole rescaling published in Nehab et al. 2011, integral image and
 * overlap coefficients.  float/double rounding happens at the same points as in the reference
 * (float results of double expressions) so the coefficient sets agree to an ulp or two.
 */
#include <iir_coeff.h>

namespace {

// q(sigma): scale of the pole radius, double expression -> float (lib/iir_coeff.cpp:38-40)
float pole_scale(float sigma) { return float(0.00399341 + 0.4715161 * double(sigma)); }

// pole d of the unit-sigma filter moved to sigma: |d|^(1/q) at angle arg(d)/q (lib/iir_coeff.cpp:60-85)
std::complex<double> rescale_pole(std::complex<double> d, float sigma)
{
    const double q = double(pole_scale(sigma));
    return std::polar(std::pow(std::abs(d), 1.0 / q), std::arg(d) / q);
}
double rescale_pole(double d, float sigma) { return std::pow(d, 1.0 / double(pole_scale(sigma))); }

struct Weights { float b0; float a[3]; };

// real pole: H(z) = b0 / (1 + a1 z^-1)     (lib/iir_coeff.cpp:103-108)
Weights first_order(float sigma)
{
    const float d = float(rescale_pole(1.86543, sigma));
    Weights w = { float(-(1.0 - double(d)) / double(d)), { float(-1.0 / double(d)), 0.0f, 0.0f } };
    return w;
}

// complex pole pair: H(z) = b0 / (1 + a1 z^-1 + a2 z^-2)     (lib/iir_coeff.cpp:124-131)
Weights second_order(float sigma)
{
    const std::complex<double> d = rescale_pole(std::complex<double>(1.41650, 1.00829), sigma);
    const float mag = float(std::abs(d));
    const float n2 = mag * mag;
    const float re = float(d.real());
    Weights w = { float((1.0 - 2.0 * double(re) + double(n2)) / double(n2)),
                  { float(-2.0 * double(re) / double(n2)), float(1.0 / double(n2)), 0.0f } };
    return w;
}

// third order = first order * second order (lib/iir_coeff.cpp:150-159)
Weights third_order(float sigma)
{
    const Weights p = first_order(sigma), q = second_order(sigma);
    Weights w;
    w.a[0] = p.a[0] + q.a[0];
    w.a[1] = p.a[0] * q.a[0] + q.a[1];
    w.a[2] = p.a[0] * q.a[1];
    w.b0 = p.b0 * q.b0;
    return w;
}

unsigned long factorial(int n) { unsigned long f = 1; for (int i = 2; i <= n; ++i) f *= (unsigned long)i; return f; }

} // namespace

std::vector<float> gaussian_weights(float sigma, int order)
{
    const Weights w = order == 1 ? first_order(sigma) : (order == 2 ? second_order(sigma) : third_order(sigma));
    std::vector<float> c(order + 1, 0.0f);
    c[0] = w.b0;
    for (int k = 1; k <= order && k <= 3; ++k) c[k] = -w.a[k - 1];     // textbook feedback is subtracted, ours is added
    return c;
}

std::vector<float> integral_image_coeff(int n)
{
    // feedback = -(coefficients of (1 - x)^n without the leading 1)
    std::vector<float> c(n + 1, 0.0f);
    c[0] = 1.0f;
    for (int i = 1; i <= n; ++i) {
        const float binom = float(factorial(n) / (factorial(i) * factorial(n - i)));
        c[i] = -1.0f * float(std::pow(-1.0, i) * double(binom));
    }
    return c;
}

std::vector<float> overlap_feedback_coeff(std::vector<float> a, std::vector<float> b)
{
    // (1 - sum a_k z^-k)(1 - sum b_k z^-k) = 1 - sum c_k z^-k
    std::vector<float> pa(1, 1.0f), pb(1, 1.0f);
    for (float v : a) pa.push_back(-v);
    for (float v : b) pb.push_back(-v);
    std::vector<float> pc(pa.size() + pb.size() - 1, 0.0f);
    for (size_t i = 0; i < pa.size(); ++i)
        for (size_t j = 0; j < pb.size(); ++j) pc[i + j] += pa[i] * pb[j];
    std::vector<float> c;
    for (size_t i = 1; i < pc.size(); ++i) c.push_back(-pc[i]);
    return c;
}

float gaussian(float x, float mu, float sigma)
{
    const float y = (x - mu) / sigma;
    return float(std::exp(-0.5 * double(y) * double(y)) / (double(sigma) * 2.50662827463));
}

float gaussDerivative(float x, float mu, float sigma)
{
    const float y = (x - mu) / sigma;
    return float(double(mu - x) * std::exp(-0.5 * double(y) * double(y)) /
                 (double(sigma) * double(sigma) * double(sigma) * 2.50662827463));
}

float gaussIntegral(float x, float mu, float sigma)
{
    return float(0.5 * (1.0 + std::erf(double(x - mu) / (double(sigma) * 1.41421356237))));
}

int gaussian_box_filter(int k, float sigma)
{
    // width of the box whose k-fold convolution has the variance of the Gaussian
    float sum = 0.0f;
    const float alpha = 0.005f;
    const int limit = int(std::floor((float(k) - 1.0f) / 2.0f));
    for (int i = 0; i <= limit; ++i) {
        const float choose = float(factorial(k) / (factorial(i) * factorial(k - i)));
        const float sign_over = float(std::pow(-1.0, i) / double(float(factorial(k - 1))));
        sum += float(double(sign_over) * double(choose) * std::pow(double(float(k)) / 2.0 - i, k - 1));
    }
    sum = float(std::sqrt(2.0 * M_PI) * double(sum + alpha) * double(sigma));
    return int(std::ceil(sum));
}
