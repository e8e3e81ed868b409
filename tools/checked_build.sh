#!/bin/bash
# build/variants/lib_checked.so: the library with -DPCAMV_CHECKED (address-range checks on every reference-plane load of the
# evaluators and of the motion compensation; the kernel traps on a violation).  Run the GPU parity tests against it with
#   PCAMV_LIB=$PWD/build/variants/lib_checked.so python -m pytest tests/test_gpu_frame.py tests/test_gpu_kernels.py -m gpu -q
cd "$(dirname "$0")/.."
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DPCAMV_CHECKED \
     -o build/variants/lib_checked.so video-steganography-pcamv_b200/csrc/*.cu && echo built build/variants/lib_checked.so
