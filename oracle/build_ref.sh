#!/bin/sh
# Compile the reference's own camera-model sources (where they lie under /root/reference) into
# oracle/_ref/libref_camera.so.  Runs only where the reference tree exists (the build container);
# the GPU box uses the prebuilt file.  The rest of the hot path cannot be built this way:
# bundle_adjustment.cc needs Ceres + Eigen, triangulation.cc needs Eigen::JacobiSVD, feature.cc
# needs OpenCV 2.4 C++ headers — none of which are in the image.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=/root/reference/src
[ -d "$REF" ] || { echo "no reference tree; keeping prebuilt oracle/_ref"; exit 0; }
mkdir -p "$HERE/_ref"
/usr/bin/g++ -O2 -fPIC -shared -std=c++11 -ffp-contract=off -I"$HERE/ref_wrap" -I"$REF" \
    "$HERE/ref_wrap/ref_camera_wrap.cc" "$REF/base3d/camera_models.cc" -o "$HERE/_ref/libref_camera.so"
echo "built $HERE/_ref/libref_camera.so"
