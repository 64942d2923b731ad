#!/bin/bash
# build kernel-variant libraries: tools/build_variants.sh name1 "-DRVE_X=1 -DRVE_Y=2" name2 "..." ...
mkdir -p deepbnd_b200/lib/variants
while [ $# -ge 2 ]; do
  ( DEEPBND_RVE_LIB=$PWD/deepbnd_b200/lib/variants/$1.so DEEPBND_NVCC_EXTRA="$2" python -m deepbnd_b200.build --force -v 2> deepbnd_b200/lib/variants/$1.ptxas.txt > /dev/null || echo "BUILD FAILED $1" ) &
  shift 2
done
wait
ls -la deepbnd_b200/lib/variants/*.so
