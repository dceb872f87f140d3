/*
 * usm_check.cpp -- the unsharp mask of the reference's apps/usm/unsharp_mask_naive.cpp:41-61, restated with a
 * check (the reference app only times it): USM = (1+w)*image - w*blur, a pointwise combination of an image and
 * a filter result.  The blur is also realized on its own and combined on the host; the two must agree.
 */
#include <Halide.h>
#include <recfilter.h>
#include <iir_coeff.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>

using namespace Halide;
using std::vector;

int main(int argc, char** argv)
{
    const int width = argc > 1 ? atoi(argv[1]) : 256, height = width, tile_width = 32;
    Image<float> image(width, height);
    srand(4321);
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++) image(x, y) = float(rand() % 1024) / 1024.0f;

    const float sigma = 5.0f, weight = 1.0f;
    vector<float> W3 = gaussian_weights(sigma, 3);

    RecFilter USM("USM"), B("Blur");
    RecFilterDim x("x", width), y("y", height);
    B.set_clamped_image_border();
    B(x, y) = image(x, y);
    B.add_filter(+x, W3);
    B.add_filter(-x, W3);
    B.add_filter(+y, W3);
    B.add_filter(-y, W3);
    vector<RecFilter> fc = B.cascade_by_dimension();
    fc[0].split_all_dimensions(tile_width);
    fc[1].split_all_dimensions(tile_width);
    USM(x, y) = (1.0f + weight) * image(x, y) - (weight) * fc[1](x, y);
    RecFilter::set_max_threads_per_cuda_warp(128);
    fc[0].gpu_auto_schedule();
    fc[1].gpu_auto_schedule();
    USM.gpu_auto_schedule(tile_width);

    Image<float> out(USM.realize());
    Image<float> blur(fc[1].realize());
    double worst = 0.0, scale = 0.0;
    for (int j = 0; j < height; j++)
        for (int i = 0; i < width; i++) {
            const float ref = (1.0f + weight) * image(i, j) - weight * blur(i, j);
            worst = std::max(worst, (double)std::fabs(ref - out(i, j)));
            scale = std::max(scale, (double)std::fabs(ref));
        }
    const double pct = 100.0 * worst / (scale + 1e-9);
    printf("unsharp mask vs host combination of image and blur: Max  relative error = %g %%\n", pct);
    return pct <= 1e-4 ? 0 : 1;
}
