/*
 * cascade_check.cpp -- a cascaded filter (RecFilter::cascade_by_dimension / cascade_by_causality,
 * lib/reorder.cpp:28-229 in the reference) must give the result of the filter it was cut from, whether the
 * launch planner fuses the links back into one plan (default) or runs one plan per link
 * (RECFILTER_NO_CHAIN_FUSION=1).  Random input, clamped border, 3rd-order Gaussian in x and y.
 */
#include <Halide.h>
#include <recfilter.h>
#include <iir_coeff.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>

using namespace Halide;

static RecFilter make(const char* name, Image<float>& img, RecFilterDim x, RecFilterDim y, const std::vector<float>& w)
{
    RecFilter f(name);
    f.set_clamped_image_border();
    f(x, y) = img(x, y);
    f.add_filter(+x, w); f.add_filter(-x, w);
    f.add_filter(+y, w); f.add_filter(-y, w);
    return f;
}

static double max_rel(const Image<float>& a, const Image<float>& b, int n)
{
    double worst = 0.0, scale = 0.0;
    for (int r = 0; r < n; r++)
        for (int c = 0; c < n; c++) {
            worst = std::max(worst, (double)std::fabs(a(c, r) - b(c, r)));
            scale = std::max(scale, (double)std::fabs(a(c, r)));
        }
    return worst / (scale + 1e-12);
}

int main(int argc, char** argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 256;
    Image<float> img(n, n);
    srand(777);
    for (int r = 0; r < n; r++)
        for (int c = 0; c < n; c++) img(c, r) = float(rand() % 4096) / 4096.0f;
    const std::vector<float> w3 = gaussian_weights(4.0f, 3);
    RecFilterDim x("x", n), y("y", n);

    RecFilter whole = make("whole", img, x, y, w3);
    whole.split_all_dimensions(32);
    Image<float> want(whole.realize());

    RecFilter by_dim = make("by_dim", img, x, y, w3);
    std::vector<RecFilter> a = by_dim.cascade_by_dimension();          // {+x,-x} then {+y,-y}
    for (RecFilter& f : a) f.split_all_dimensions(32);
    Image<float> got_dim(a.back().realize());

    RecFilter by_caus = make("by_caus", img, x, y, w3);
    std::vector<RecFilter> b = by_caus.cascade_by_causality();         // causal scans then anticausal scans
    for (RecFilter& f : b) f.split_all_dimensions(32);
    Image<float> got_caus(b.back().realize());

    const double e1 = max_rel(want, got_dim, n), e2 = max_rel(want, got_caus, n);
    printf("cascade_by_dimension vs the uncut filter: Max  relative error = %g %%\n", 100.0 * e1);
    printf("cascade_by_causality vs the uncut filter: Max  relative error = %g %%\n", 100.0 * e2);
    return (e1 <= 1e-5 && e2 <= 1e-5) ? 0 : 1;
}
