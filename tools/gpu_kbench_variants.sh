#!/bin/bash
# one GPU session: every tools/bin/kbench_* variant (DP kernel alone, 150x200 pairs), then the parity tests of the library build
N=${1:-262144}
for v in ${VARIANTS:-base st16 shfl p32 all allpf}; do
	echo "=== kbench_$v"; tools/bin/kbench_$v $N 3
done 2>&1 | tee gpurun_out/r2_kbench_variants.txt
