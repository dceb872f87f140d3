#!/bin/bash
# Times the reference's own in-scope apps (unchanged sources, built by oracle/Makefile against include/recfilter.h and
# librecfilter_b200.so) with their own profile() loop: -w <width> -iter <n>.  Development aid; run on the GPU box.
export LD_LIBRARY_PATH=recfilter_b200:$LD_LIBRARY_PATH
W=${1:-4096}; IT=${2:-100}
for app in summed_table gaussian_filter_3xy gaussian_filter_3x_3y gaussian_filter_1xy_2xy gaussian_filter_1xy_2x_2y \
           gaussian_filter_1xy_1xy_1xy box_filter_1 box_filter_3 box_filter_6 unsharp_mask_naive unsharp_mask_optimized \
           bicubic_filter biquintic_cascaded_filter biquintic_overlapped_filter; do
    printf "%-32s w=%-5s " $app $W
    timeout 120 oracle/_ref/gpu/$app -w $W -iter $IT 2>&1 | grep "ms per iteration" | tail -1 | sed 's/ over .*iteration(s)//'
done
for app in audio_filter_high_order audio_filter_biquads; do
    printf "%-32s          " $app
    timeout 300 oracle/_ref/gpu/$app -iter 20 2>&1 | grep "ms per iteration" | tail -1 | sed 's/ over .*iteration(s)//'
done
