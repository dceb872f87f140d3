/*
 * tuple_check.cpp -- Tuple (multi-output) filters of the C++ surface (lib/recfilter.cpp:68-74,197-203 in the
 * reference: every Tuple element is filtered independently with the same scans; the demos filter RGB images that
 * way, demo/demo_gaussian_filter.cpp:50-60).  A 3-plane image is filtered as one Tuple filter and, plane by
 * plane, as three ordinary filters; the results must be identical.  Also checks F(x,y)[i] as another filter's input.
 */
#include <Halide.h>
#include <recfilter.h>
#include <iir_coeff.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>

using namespace Halide;

int main(int argc, char** argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 128;
    Image<float> rgb(n, n, 3);
    srand(99);
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < n; r++)
            for (int x = 0; x < n; x++) rgb(x, r, c) = float(rand() % 256) / 255.0f;
    const std::vector<float> g2 = gaussian_weights(3.0f, 2);

    RecFilterDim x("x", n), y("y", n);
    RecFilter T("rgb_blur");
    T.set_clamped_image_border();
    T(x, y) = Tuple(rgb(x, y, 0), rgb(x, y, 1), rgb(x, y, 2));
    T.add_filter(+x, g2); T.add_filter(-x, g2);
    T.add_filter(+y, g2); T.add_filter(-y, g2);
    T.split_all_dimensions(32);
    Realization out = T.realize();
    if (out.size() != 3) { printf("expected 3 outputs, got %d\n", (int)out.size()); return 1; }

    double worst = 0.0;
    for (int c = 0; c < 3; c++) {
        Image<float> plane(n, n);
        for (int r = 0; r < n; r++)
            for (int i = 0; i < n; i++) plane(i, r) = rgb(i, r, c);
        RecFilter S("single");
        S.set_clamped_image_border();
        S(x, y) = plane(x, y);
        S.add_filter(+x, g2); S.add_filter(-x, g2);
        S.add_filter(+y, g2); S.add_filter(-y, g2);
        S.split_all_dimensions(32);
        Image<float> want(S.realize()), got(out[c]);
        for (int r = 0; r < n; r++)
            for (int i = 0; i < n; i++) worst = std::max(worst, (double)std::fabs(want(i, r) - got(i, r)));
    }
    // element 1 of the Tuple filter as the input of another filter (a plain prefix sum along x)
    RecFilter P("prefix_of_green");
    P(x, y) = T(x, y)[1];
    P.add_filter(+x, { 1.0f, 1.0f });
    Image<float> p(P.realize()), green(out[1]);
    double worst2 = 0.0;
    for (int r = 0; r < n; r++) {
        float acc = 0.0f;
        for (int i = 0; i < n; i++) { acc += green(i, r); worst2 = std::max(worst2, (double)std::fabs(acc - p(i, r)) / (std::fabs(acc) + 1e-9)); }
    }
    printf("Tuple filter vs per-plane filters: Max  relative error = %g %%\n", 100.0 * worst);
    printf("filter of Tuple element [1] vs serial prefix sum: Max  relative error = %g %%\n", 100.0 * worst2);
    return (worst == 0.0 && worst2 <= 1e-5) ? 0 : 1;
}
