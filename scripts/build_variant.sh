#!/bin/bash
# build_variant.sh NAME [extra nvcc flags...]  -> hmm_training_b200/libhmmb200_NAME.so
set -e
NAME=$1; shift
cd "$(dirname "$0")/../hmm_training_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fopenmp -shared "$@" -o ../libhmmb200_$NAME.so context.cu vq.cu bw.cu loader.cu mfcc.cu comm.cu -ldl
echo built $NAME
