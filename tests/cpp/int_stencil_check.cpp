/*
 * int_stencil_check.cpp -- integer definitions lowered to the pointwise stencil epilogue (rf_stencil_execute).
 * An int32 summed-area table followed by the 4-tap finite differencing of apps/box/box_filter.h:36-39 WITHOUT the
 * "/ area" scale must equal the direct (2B+1)^2 box sum bit for bit, and weights other than +-1 must be applied
 * as ring elements (2*f(x) + 3*f(x+1)).  With "--div" the program defines "(...) / area" on the integer filter:
 * Halide's integer division is not a linear scale, so the host layer must refuse it (message + assert) instead
 * of rounding the weights silently.
 */
#include "recfilter.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>

using namespace Halide;

int main(int argc, char** argv)
{
    const bool div = argc > 1 && !strcmp(argv[1], "--div");
    const int B = 3, w = 128, h = 128;
    Image<int32_t> I(w, h);
    srand(7);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) I(x, y) = (x < 8 || x >= w - 8 || y < 8 || y >= h - 8) ? 0 : (rand() % 31) - 15;

    RecFilter::set_max_threads_per_cuda_warp(128);
    RecFilterDim x("x", w), y("y", h);
    RecFilter S("Sat");
    S(x, y) = I(x, y);
    S.add_filter(+x, {1.0f, 1.0f});
    S.add_filter(+y, {1.0f, 1.0f});
    S.split_all_dimensions(32);

    RecFilter D("Diff");
    Func s = S.as_func();
    Expr xp = clamp(x + B, 0, w - 1), xm = clamp(x - B - 1, 0, w - 1);
    Expr yp = clamp(y + B, 0, h - 1), ym = clamp(y - B - 1, 0, h - 1);
    if (div) {
        D(x, y) = (s(xp, yp) - s(xp, ym) - s(xm, yp) + s(xm, ym)) / ((2 * B + 1) * (2 * B + 1));   // must die
        return 0;
    }
    D(x, y) = s(xp, yp) - s(xp, ym) - s(xm, yp) + s(xm, ym);
    Image<int32_t> out(D.realize());
    long bad = 0;
    for (int yy = B + 1; yy < h - B - 1; yy++)
        for (int xx = B + 1; xx < w - B - 1; xx++) {
            int32_t ref = 0;
            for (int dy = -B; dy <= B; dy++)
                for (int dx = -B; dx <= B; dx++) ref += I(xx + dx, yy + dy);
            if (ref != out(xx, yy)) bad++;
        }
    printf("integer box sum from the summed-area table: %ld mismatches\n", bad);

    RecFilter Wt("Weighted");
    Wt(x, y) = 2 * I(x, y) + 3 * I(clamp(x + 1, 0, w - 1), y);
    Image<int32_t> o2(Wt.realize());
    long bad2 = 0;
    for (int yy = 0; yy < h; yy++)
        for (int xx = 0; xx < w; xx++)
            if (o2(xx, yy) != 2 * I(xx, yy) + 3 * I(xx + 1 < w ? xx + 1 : w - 1, yy)) bad2++;
    printf("integer weights 2, 3: %ld mismatches\n", bad2);
    return (bad == 0 && bad2 == 0) ? 0 : 1;
}
