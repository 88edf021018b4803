#!/bin/bash
# build_variant.sh NAME "-DFOO=1 ..." -> piano_a2s_b200/variants/libpa2s_NAME.so (kernel experiments; select with PA2S_LIB=...)
set -e
cd "$(dirname "$0")/../piano_a2s_b200"
mkdir -p variants build_$1
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
objs=""
for f in csrc/*.cu; do
  o=build_$1/$(basename ${f%.cu}).o
  if grep -q "$3" <<< "$f" || [ ! -f build/$(basename ${f%.cu}).o ]; then
    (cd csrc && nvcc $FLAGS $2 -c $(basename $f) -o ../$o) &
  else
    cp build/$(basename ${f%.cu}).o $o
  fi
  objs="$objs $o"
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o variants/libpa2s_$1.so $objs
rm -rf build_$1
echo variants/libpa2s_$1.so
