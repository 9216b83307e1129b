#!/bin/bash
# Kernel A/B builds of the lane-per-SNP solve (CPU side): recompiles only k3_inst.cu for P=4 with each flag set and links it
# against the objects of the current in-tree build.  Output: ab/libjxb200_v<i>.so (+ ab/variants.txt); select one at run time
# with JXB_LIB_PATH.  usage: tools/k3_ab_build.sh "<flags of v0>" "<flags of v1>" ...
set -eu
cd "$(dirname "$0")/.."
mkdir -p ab
: > ab/variants.txt
i=0
for v in "$@"; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3,-fno-fast-math \
      --expt-relaxed-constexpr -DJXB_P=4 -fmad=false $v -ccbin /usr/bin/g++ -Xptxas=-v -c janusx_b200/csrc/k3_inst.cu -o ab/k3_inst_p4_v$i.o \
      2> ab/ptxas_v$i.log
  objs=$(ls janusx_b200/build/*.o | grep -v k3_inst_p4.o)
  /usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -o ab/libjxb200_v$i.so $objs ab/k3_inst_p4_v$i.o -lcudart -lpthread -ldl
  echo "v$i: '$v' $(grep -A2 'solve_lane_kernelILi4ELb1' ab/ptxas_v$i.log | grep -o 'Used [0-9]* registers\|[0-9]* bytes spill stores' | tr '\n' ' ')" | tee -a ab/variants.txt
  i=$((i+1))
done
