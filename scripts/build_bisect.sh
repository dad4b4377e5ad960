# Bisect builds of the library for scripts/diag_subsonic.py: one restructured operation switched back
# to the reference's form at a time (physics.cuh). Only the translation units the diagnosed cases
# need (main + one species, laminar / RANS). Output: build/bisect/lib_<variant>.so (git-ignored).
set -e
cd "$(dirname "$0")/.."
mkdir -p build/bisect
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 177"
for v in EXACT_RCP REF_ROE MUSCL_DIV; do
  (
  for tu in MAIN 10 12 20 22 30 32; do
    if [ $tu = MAIN ]; then d="-DAITHER_MAIN_TU"; else d="-DAITHER_EQ_TU=$tu"; fi
    nvcc $FLAGS -DAITHER_BISECT_$v $d -c -o build/bisect/${v}_$tu.o aither_b200/csrc/aither_gpu.cu &
  done
  wait
  nvcc $FLAGS -shared -o build/bisect/lib_$v.so build/bisect/${v}_*.o -lcudart -ldl
  ) 
done
ls -la build/bisect/*.so
