/*
 * dog_check.cpp -- a definition that reads TWO filters: the last stage of a difference of Gaussians subtracts
 * the two elements of a Tuple filter (DoG(x,y) = diff(SAT2y)[0] - diff(SAT2y)[1], /root/reference/apps/DoG/diff_gauss.cpp:
 * 84-96, which the reference only times).  Here the two band-pass inputs are the elements of a Tuple filter fed from a
 * 16-bit image through a float conversion, then filtered again along y as a second Tuple filter; the difference is also
 * formed on the host from the two realized elements, and both must agree.
 */
#include <Halide.h>
#include <recfilter.h>
#include <iir_coeff.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

using namespace Halide;

int main(int argc, char** argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 128;
    Image<int16_t> raw(n, n);
    srand(2024);
    for (int r = 0; r < n; r++)
        for (int c = 0; c < n; c++) raw(c, r) = int16_t(rand() % 2048 - 1024);

    RecFilterDim u("u", n), v("v", n);
    RecFilter lum("lum");
    lum(u, v) = Internal::Cast::make(type_of<float>(), raw(u, v));

    const std::vector<float> narrow = gaussian_weights(1.5f, 2), wide = gaussian_weights(3.0f, 2);
    // two smoothings of the same picture along u, as ONE Tuple filter ...
    RecFilter pair_u("pair_u");
    pair_u(u, v) = Tuple(lum.as_func()(u, v), 0.5f * lum.as_func()(u, v));
    pair_u.add_filter(+u, narrow);
    pair_u.add_filter(-u, narrow);
    // ... continued along v by a second Tuple filter that reads the elements of the first
    RecFilter pair_v("pair_v");
    pair_v(u, v) = Tuple(pair_u.as_func()(u, v)[0], pair_u.as_func()(u, v)[1]);
    pair_v.add_filter(+v, wide);
    pair_v.add_filter(-v, wide);
    pair_u.split_all_dimensions(32);
    pair_v.split_all_dimensions(32);

    // the band-pass picture: element 0 minus twice element 1, one tap shifted by a clamped offset
    RecFilter band("band");
    band(u, v) = pair_v.as_func()(u, v)[0] - 2.0f * pair_v.as_func()(min(u + 3, n - 1), v)[1];

    Image<float> got(band.realize());
    Realization both = pair_v.realize();
    Image<float> e0(both[0]), e1(both[1]);
    double worst = 0.0, scale = 0.0;
    for (int r = 0; r < n; r++)
        for (int c = 0; c < n; c++) {
            const float want = e0(c, r) - 2.0f * e1(std::min(c + 3, n - 1), r);
            worst = std::max(worst, (double)std::fabs(want - got(c, r)));
            scale = std::max(scale, (double)std::fabs(want));
        }
    const double pct = 100.0 * worst / (scale + 1e-9);
    printf("difference of two Tuple elements vs host combination: Max  relative error = %g %%\n", pct);
    // timed like the reference's program: the second source chain is evaluated inside every iteration
    const float ms = band.profile(3);
    return (pct <= 1e-4 && ms > 0.0f) ? 0 : 1;
}
