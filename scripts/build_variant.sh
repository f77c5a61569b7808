#!/bin/bash
# build_variant.sh NAME 'sed-expression' : builds minddet_b200/lib/variant_NAME.so from a patched copy of proposal.cu
# (experiments only: MD_REGION_LIB=minddet_b200/lib/variant_NAME.so python bench.py)
set -e
cd "$(dirname "$0")/.."
name=$1; expr=$2
tmp=minddet_b200/lib/_variant_$name
rm -rf $tmp; mkdir -p $tmp
cp minddet_b200/csrc/* $tmp/
sed -i "$expr" $tmp/*.cu $tmp/*.cuh
FL="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-fvisibility=hidden -cudart static -I include -I minddet_b200/csrc"
/usr/local/cuda/bin/nvcc $FL -c $tmp/proposal.cu -o $tmp/proposal.o
objs=""
for f in assign roialign roialign_tma bev yolo rcnn_post aot_entry; do objs="$objs minddet_b200/lib/$f.o"; done
/usr/local/cuda/bin/nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o minddet_b200/lib/variant_$name.so $tmp/proposal.o $objs
rm -rf $tmp
echo built minddet_b200/lib/variant_$name.so
