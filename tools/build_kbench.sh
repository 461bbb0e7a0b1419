#!/bin/bash
# builds tools/bin/kbench_<variant>: the DP kernel alone with compile-time variants (see gd_ksw.cuh GD_KSW_*)
cd "$(dirname "$0")" && mkdir -p bin
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr"
build() { # name, flags
	$NV $2 $KB_EXTRA -o bin/kbench_$1 ksw_kbench.cu &
}
build base  "-DGD_KSW_ST16=0 -DGD_KSW_NBSHFL=0 -DGD_KSW_P32=0 -DGD_KSW_DP4A=0"
build dp4a  "-DGD_KSW_DP4A=1"
build notimad "-DGD_KSW_DP4A=1 -DGD_KSW_NOTIMAD=1"
build st16  "-DGD_KSW_ST16=1 -DGD_KSW_NBSHFL=0 -DGD_KSW_P32=0"
build shfl  "-DGD_KSW_ST16=0 -DGD_KSW_NBSHFL=1 -DGD_KSW_P32=0"
build p32   "-DGD_KSW_ST16=0 -DGD_KSW_NBSHFL=0 -DGD_KSW_P32=1"
build all   "-DGD_KSW_ST16=1 -DGD_KSW_NBSHFL=1 -DGD_KSW_P32=1"
build allpf "-DGD_KSW_ST16=1 -DGD_KSW_NBSHFL=1 -DGD_KSW_P32=1 -DGD_KSW_PREFETCH=1"
wait; ls -la bin
