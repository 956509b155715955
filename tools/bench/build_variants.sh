#!/bin/bash
# builds psd_bench variants into tools/bench/bin (development only)
cd "$(dirname "$0")"
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17"
b() { name=$1; shift; $NV -DVARIANT_NAME="\"$name\"" "$@" -o bin/psd_$name psd_bench.cu 2>&1 | grep -E "error" ; }
b base_b128 &
b b64 -DBLOCK=64 &
b b256 -DBLOCK=256 &
b b128_min8 -DBLOCK=128 -DMINB=8 &
wait
b b128_min9 -DBLOCK=128 -DMINB=9 &
b b128_min10 -DBLOCK=128 -DMINB=10 &
b b256_min4 -DBLOCK=256 -DMINB=4 &
b lanes2 -DPSD_SBOX_LANES=2 &
wait
b lanes4 -DPSD_SBOX_LANES=4 &
b lanes6 -DPSD_SBOX_LANES=6 &
b lanes12 -DPSD_SBOX_LANES=12 &
b lanes4_min8 -DPSD_SBOX_LANES=4 -DMINB=8 &
wait
ls bin
