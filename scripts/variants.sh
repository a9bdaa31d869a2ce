#!/bin/bash
# Builds variants of libdmgs_raster.so that differ in -D flags of the blend kernels (experiments):
#   bash scripts/variants.sh name1 "-DFOO=1" name2 "-DBAR=2 -DBAZ" ...   -> build/variants/libdmgs_<name>.so
# Run one with DMGS_RASTER_LIB=build/variants/libdmgs_<name>.so python bench.py ...
set -e
cd "$(dirname "$0")/../dmgs_b200/csrc"
make -s -j8
mkdir -p ../../build/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
FILES=${VARIANT_FILES:-"blend blend_bwd"}
OTHERS=""
for o in *.o; do b=${o%.o}; case " $FILES " in *" $b "*) ;; *) OTHERS="$OTHERS $o";; esac; done
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  objs=""
  for f in $FILES; do
    nvcc -O3 -std=c++17 $ARCH -lineinfo -fmad=false -Xcompiler -fPIC,-O2 -Xptxas -v $flags -dc -o /tmp/v_${name}_$f.o $f.cu 2> /tmp/v_${name}_$f.log
    grep -h "Used" /tmp/v_${name}_$f.log | sed "s/^/$name $f: /"
    objs="$objs /tmp/v_${name}_$f.o"
  done
  nvcc $ARCH -shared -o ../../build/variants/libdmgs_$name.so $objs $OTHERS -lcudart
done
