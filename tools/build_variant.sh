#!/bin/bash
# usage: tools/build_variant.sh NAME [nvcc flags...]  -> scratch_ab/lib_NAME.so
# A/B builds of the whole library with extra nvcc flags; select one at run time with SURF_B200_LIB=/root/repo/scratch_ab/lib_NAME.so
set -e
name=$1; shift
cd /root/repo/surffeatures_b200
mkdir -p ../scratch_ab/$name
NV="/usr/local/cuda/bin/nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a"
for u in surf_kernels:-fmad=false surf_describe:-fmad=false surf_match: surf_match_tc: surf_engine: surf_db: surf_comm: surf_vote:-fmad=false surf_mha:-fmad=false surf_pose:-fmad=false; do
  src=${u%%:*}; fl=${u#*:}
  $NV -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I ../include -I csrc $fl "$@" -c csrc/$src.cu -o ../scratch_ab/$name/$src.o &
done
wait
$NV -shared -o ../scratch_ab/lib_$name.so ../scratch_ab/$name/*.o -lcudart -ldl
rm -rf ../scratch_ab/$name
echo built lib_$name.so
