#!/bin/bash
# builds tools/bin/kbench_p{0,1}h{0,1}
cd "$(dirname "$0")" && mkdir -p bin
for P in 0 1; do for H in 0 1; do
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -DGD_KSW_PREFETCH=$P -DGD_KSW_HOTMEM=$H $KB_EXTRA -o bin/kbench_p${P}h${H} ksw_kbench.cu &
done; done; wait; ls -la bin
