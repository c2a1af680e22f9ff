#!/bin/bash
# build_variant.sh NAME [nvcc -D flags...]: a copy of the library with scan.cu compiled under extra defines, as
# regex_fpga_b200/lib/variants/librfb200_NAME.so; select it with RFB_LIB=<path>.  Dev tool for kernel experiments.
set -e
cd "$(dirname "$0")/../../regex_fpga_b200/csrc"
name=$1; shift
make -s >/dev/null 2>&1
mkdir -p ../lib/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC "$@" -c scan.cu -o ../lib/variants/scan_$name.o
objs=$(ls ../lib/*.o | grep -v '/scan.o')
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/variants/librfb200_$name.so $objs ../lib/variants/scan_$name.o -cudart static
rm ../lib/variants/scan_$name.o
echo ../lib/variants/librfb200_$name.so
