# usage: build_variant.sh NAME -DFLAG... ; builds codlad_b200/_variants/lib_NAME.so
set -e
name=$1; shift
mkdir -p codlad_b200/_variants /tmp/var_$name
for f in codlad_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr "$@" -c $f -o /tmp/var_$name/$(basename $f .cu).o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o codlad_b200/_variants/lib_$name.so /tmp/var_$name/*.o
echo built $name
