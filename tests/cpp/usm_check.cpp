/*
 * usm_check.cpp -- checks the two-source pointwise stencil of the C++ surface: an unsharp mask
 * sharp = (1 + amount) * picture - amount * smooth, where `smooth` is a cascaded 3rd-order recursive
 * Gaussian of `picture` (the construction the reference times in apps/usm/unsharp_mask_naive.cpp:41-61
 * without checking it).  `smooth` is also realized on its own and combined on the host; both must agree.
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
    const int n = argc > 1 ? atoi(argv[1]) : 256;
    const float amount = 1.0f;
    Halide::Image<float> picture(n, n);
    srand(4321);
    for (int r = 0; r < n; r++)
        for (int c = 0; c < n; c++) picture(c, r) = float(rand() % 1024) / 1024.0f;

    RecFilterDim col("col", n), row("row", n);
    RecFilter smooth("smooth"), sharp("sharp");
    const std::vector<float> g3 = gaussian_weights(5.0f, 3);
    smooth.set_clamped_image_border();
    smooth(col, row) = picture(col, row);
    for (int pass = 0; pass < 2; pass++) {                 // causal + anticausal along each dimension
        smooth.add_filter(+(pass ? row : col), g3);
        smooth.add_filter(-(pass ? row : col), g3);
    }
    std::vector<RecFilter> stages = smooth.cascade_by_dimension();
    for (RecFilter& s : stages) s.split_all_dimensions(32);
    RecFilter& last = stages.back();

    sharp(col, row) = (1.0f + amount) * picture(col, row) - amount * last(col, row);
    RecFilter::set_max_threads_per_cuda_warp(128);
    for (RecFilter& s : stages) s.gpu_auto_schedule();
    sharp.gpu_auto_schedule(32);

    Halide::Image<float> got(sharp.realize());
    Halide::Image<float> blur(last.realize());
    double worst = 0.0, scale = 0.0;
    for (int r = 0; r < n; r++)
        for (int c = 0; c < n; c++) {
            const float want = (1.0f + amount) * picture(c, r) - amount * blur(c, r);
            worst = std::max(worst, (double)std::fabs(want - got(c, r)));
            scale = std::max(scale, (double)std::fabs(want));
        }
    const double pct = 100.0 * worst / (scale + 1e-9);
    printf("unsharp mask vs host combination of image and blur: Max  relative error = %g %%\n", pct);
    return pct <= 1e-4 ? 0 : 1;
}
