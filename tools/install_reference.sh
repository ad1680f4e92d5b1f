#!/bin/bash
# Reference install for the GPU-side baseline and the "runs unchanged" tests.
#
# The reference (yunminjin2/KD-PointCloud) has no installable package: its layer library, models and
# drivers are loose scripts and pointnet2/setup.py only builds the CUDA extension.  This recipe copies
# the UNMODIFIED Python sources the hot path touches into baseline/_ref/ (git-ignored, NOT
# gpurun-ignored: it travels to the GPU box exactly like oracle/_ref) so that
#   * tests/test_unchanged_reference_gpu.py can import models_bid_pointconv.py / models_bid_lighttoken_res.py /
#     loss_functions.py byte-for-byte as they are, once through compat/ (the kdpc kernels) and once through the
#     reference's own pointconv_util.py + pointnet2_utils.py + its own CUDA kernels (oracle/_ref/libpointnet2_ref.so
#     behind a `pointnet2_cuda` ctypes module, oracle/ref_gpu.py);
#   * tools/bench_reference_gpu.py can time that stock GPU path beside ours.
# Nothing under baseline/_ref is ever committed and nothing in kd_pointcloud_b200/ imports it.
set -e
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")/.." && pwd)"
DST="$HERE/baseline/_ref"
if [ ! -d "$REF" ]; then
  echo "reference checkout not present at $REF: keeping whatever is in $DST"; exit 0
fi
mkdir -p "$DST/pointnet2" "$DST/transforms" "$DST/utils" "$DST/datasets"
cp "$REF"/*.py "$DST"/
cp "$REF"/*.yaml "$DST"/ 2>/dev/null || true
cp "$REF"/pointnet2/*.py "$DST"/pointnet2/
cp "$REF"/transforms/*.py "$DST"/transforms/
cp "$REF"/utils/*.py "$DST"/utils/
cp "$REF"/datasets/*.py "$DST"/datasets/
( cd "$REF" && sha256sum *.py pointnet2/*.py transforms/*.py utils/*.py datasets/*.py ) > "$DST/SHA256SUMS"
echo "installed $(wc -l < "$DST/SHA256SUMS") unmodified reference files into $DST"
