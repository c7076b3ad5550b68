#!/bin/sh
# Builds the reference's UNMODIFIED mapper layer against the drop-in shim: every translation unit of
# /root/reference/src except the ones the shim replaces is compiled where it lies, against the stand-in
# headers in stubs/ (Eigen, OpenCV, Boost and glog are not in the image), and linked with the three shim sources and
# libmavmap_b200.so into build/mapper_harness/mapper.  Nothing of the reference is copied into the repo.
#   usage: build_mapper.sh [reference src dir] [output dir]
set -e
HERE=$(cd "$(dirname "$0")" && pwd); ROOT=$(cd "$HERE/../.." && pwd)
REF=${1:-/root/reference/src}; OUT=${2:-$ROOT/build/mapper_harness}
mkdir -p "$OUT"
INC="-I$HERE/stubs -I$ROOT/mavmap_b200/shim -I$ROOT/include -I$REF"     # shim first: base3d/bundle_adjustment.h is the drop-in header
CXX="g++ -std=c++11 -O1 -fopenmp -w"
OBJS=""
# reference sources, as they are.  Replaced by the shim and therefore NOT built: base3d/bundle_adjustment.cc, base3d/triangulation.cc
for f in mapper.cc sfm/sequential_mapper.cc fm/feature_management.cc base2d/feature.cc base2d/feature_cache.cc base2d/image.cc \
         base3d/camera_models.cc base3d/essential_matrix.cc base3d/p3p.cc base3d/projection.cc base3d/projective_transform.cc \
         base3d/similarity_transform.cc loop/detection.cc loop/voc_tree.cc loop/voc_tree_database.cc loop/voc_tree_inv_file.cc \
         util/estimation.cc util/io.cc util/math.cc util/opencv.cc util/path.cc util/timer.cc; do
  o="$OUT/ref_$(echo $f | tr '/' '_' | sed 's/\.cc$/.o/')"
  $CXX $INC -c "$REF/$f" -o "$o" &
  OBJS="$OBJS $o"
done
for f in base3d/bundle_adjustment.cc base3d/triangulation.cc base2d/feature_match.cc; do
  o="$OUT/shim_$(echo $f | tr '/' '_' | sed 's/\.cc$/.o/')"
  $CXX $INC -c "$ROOT/mavmap_b200/shim/$f" -o "$o" &
  OBJS="$OBJS $o"
done
$CXX $INC -c "$HERE/cv_impl.cc" -o "$OUT/cv_impl.o" &
$CXX $INC -c "$HERE/global_ba_driver.cc" -o "$OUT/global_ba_driver.o" &
wait
# The shim also replaces three functions that share a reference file with code that stays ("delete these
# bodies"): the reference objects keep them as WEAK symbols so that the shim's definitions are the ones linked.
weaken() { for sym in $(nm "$1" | awk '{print $3}' | grep -E "$2"); do objcopy --weaken-symbol="$sym" "$1"; done; }
weaken "$OUT/ref_base2d_feature.o" '^_Z17match_brute_force'
weaken "$OUT/ref_base3d_projection.o" '^_Z(18calc_reproj_errors|10calc_depth)'
$CXX -o "$OUT/mapper" $OBJS "$OUT/cv_impl.o" -L"$ROOT/mavmap_b200" -lmavmap_b200 -Wl,-rpath,"$ROOT/mavmap_b200"
# the same objects without mapper.cc's main(): the reference's SequentialMapper driven on a synthetic sequence (needs a GPU to run)
$CXX -o "$OUT/global_ba_driver" "$OUT/global_ba_driver.o" $(echo $OBJS | tr ' ' '\n' | grep -v ref_mapper.o) "$OUT/cv_impl.o" -L"$ROOT/mavmap_b200" -lmavmap_b200 -Wl,-rpath,"$ROOT/mavmap_b200"
echo "$OUT/mapper"
