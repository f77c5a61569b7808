#!/bin/bash
# build_variant.sh NAME 'sed-expression' ['src1 src2 ...'] : builds minddet_b200/lib/variant_NAME.so from a patched copy
# of csrc/ (the sed expression is applied to every .cu/.cuh of the copy; the listed sources -- default: proposal -- are
# recompiled, the other objects come from the regular build).  Experiments only:
#   MD_REGION_LIB=$PWD/minddet_b200/lib/variant_NAME.so python scripts/proposal_only_bench.py
set -e
cd "$(dirname "$0")/.."
name=$1; expr=$2; srcs=${3:-proposal}
tmp=minddet_b200/lib/_variant_$name
rm -rf $tmp; mkdir -p $tmp
cp minddet_b200/csrc/* $tmp/
sed -i "$expr" $tmp/*.cu $tmp/*.cuh
sed -i 's|"../../include/md_region_aot.h"|"../../../include/md_region_aot.h"|' $tmp/*.cu $tmp/*.h $tmp/*.cuh
FL="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-fvisibility=hidden -cudart static"
objs=""
for f in proposal assign roialign roialign_tma roialign_ch bev yolo rcnn_post aot_entry; do
  if [[ " $srcs " == *" $f "* ]]; then /usr/local/cuda/bin/nvcc $FL -c $tmp/$f.cu -o $tmp/$f.o; objs="$objs $tmp/$f.o"; else objs="$objs minddet_b200/lib/$f.o"; fi
done
/usr/local/cuda/bin/nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o minddet_b200/lib/variant_$name.so $objs
rm -rf $tmp
echo built minddet_b200/lib/variant_$name.so
