/*
 * cast_check.cpp -- type-converting definitions: the filter type is the type of the RHS (lib/recfilter.cpp:197).
 * F(x,y) = cast<float>(image16(x,y)) as in apps/DoG/diff_gauss.cpp:66-73, followed by a float summed-area table,
 * must equal the float table of the converted image; image16(x,y) * 0.5f likewise runs as a float filter.
 */
#include "recfilter.h"
#include <cstdio>
#include <cstdlib>

using namespace Halide;

int main()
{
    const int w = 96, h = 64;
    Image<int16_t> I(w, h);
    srand(3);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) I(x, y) = (int16_t)((rand() % 201) - 100);
    RecFilter::set_max_threads_per_cuda_warp(128);
    RecFilterDim x("x", w), y("y", h);

    RecFilter V("V");
    V(x, y) = Internal::Cast::make(type_of<float>(), I(x, y));
    RecFilter S("Sat");
    S(x, y) = V.as_func()(x, y);
    S.add_filter(+x, {1.0f, 1.0f});
    S.add_filter(+y, {1.0f, 1.0f});
    S.split_all_dimensions(32);
    Realization r = S.realize();
    Image<float> out(r);

    RecFilter H("Half");
    H(x, y) = I(x, y) * 0.5f;
    H.add_filter(+x, {1.0f, 0.5f});
    Image<float> half(H.realize());

    long bad = 0, bad2 = 0;
    for (int yy = 0; yy < h; yy++) {
        float prev = 0.0f;
        for (int xx = 0; xx < w; xx++) {
            double s = 0.0;
            for (int j = 0; j <= yy; j++) for (int i = 0; i <= xx; i++) s += I(i, j);
            if (out(xx, yy) != (float)s) bad++;                   // small integers: the float table is exact
            prev = 0.5f * (float)I(xx, yy) + 0.5f * prev;
            if (half(xx, yy) != prev) bad2++;
        }
    }
    printf("cast<float>(int16) summed-area table: %ld mismatches\n", bad);
    printf("int16 * 0.5f first-order filter: %ld mismatches\n", bad2);
    return (bad == 0 && bad2 == 0) ? 0 : 1;
}
