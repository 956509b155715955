#!/bin/bash
# compute-sanitizer is closed on this GPU pool.  Substitute for its memcheck on the kernels' INDEX MATH: the CPU replay of the
# kernel bodies (tests/emu: the same __host__ __device__ functions the GPU runs, driven block by block) built with
# AddressSanitizer, over the replay tests (commits at many shapes incl. row shards and the fused exchange, partial
# products, the split quotient kernels single and row-sharded).  Any out-of-bounds load / store aborts the run.
cd "$(dirname "$0")/.."
ASAN_LIB=$(/usr/bin/g++ -print-file-name=libasan.so)
EMU_ASAN=1 LD_PRELOAD=$ASAN_LIB ASAN_OPTIONS=detect_leaks=0:abort_on_error=1 \
  python -m pytest tests/test_replay.py tests/test_plonk_cpu.py -q -x -p no:cacheprovider "$@"
