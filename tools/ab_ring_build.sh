#!/bin/bash
# A/B build: the library with the padding-frame (ring) epilogue code compiled OUT of conv_gemm_tc / attn_fwd_tc
# -> tools/_bin/libpht_b200_noring.so (use with PHT_LIB_PATH=... PHT_FUSED_RING=0)
set -e
cd "$(dirname "$0")/.."
python -m pixel_heal_thyself_b200.build > /dev/null
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I include -I pixel_heal_thyself_b200/csrc -DPHT_NO_RING"
mkdir -p tools/_bin/noring
for f in igemm_tc attention_tc; do nvcc $FLAGS -c pixel_heal_thyself_b200/csrc/$f.cu -o tools/_bin/noring/$f.o & done; wait
OBJS=$(ls pixel_heal_thyself_b200/build/*.o | grep -v -e igemm_tc.o -e attention_tc.o)
nvcc -shared -Wno-deprecated-gpu-targets -o tools/_bin/libpht_b200_noring.so $OBJS tools/_bin/noring/igemm_tc.o tools/_bin/noring/attention_tc.o -Xlinker --no-undefined
ls -la tools/_bin/libpht_b200_noring.so
