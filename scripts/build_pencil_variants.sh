# A/B builds of the LU-SGS pencil cross-section (lusgs_pencil.cuh): aither_b200/lib/variants/lib_<name>.so
# usage: scripts/build_pencil_variants.sh "name TJ TK CTAS" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants aither_b200/lib/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 177"
build() {  # name TJ TK CTAS
  for tu in MAIN 10 12 20 22 30 32; do
    if [ $tu = MAIN ]; then d="-DAITHER_MAIN_TU"; else d="-DAITHER_EQ_TU=$tu"; fi
    nvcc $FLAGS -DAITHER_PENCIL_TJ=$2 -DAITHER_PENCIL_TK=$3 -DAITHER_PENCIL_CTAS=$4 $d -c -o build/variants/$1_$tu.o aither_b200/csrc/aither_gpu.cu &
  done
  wait
  nvcc $FLAGS -shared -o aither_b200/lib/variants/lib_$1.so build/variants/$1_*.o -lcudart -ldl
}
for v in "$@"; do build $v; done
ls -la aither_b200/lib/variants/
