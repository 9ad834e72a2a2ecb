#!/usr/bin/env bash
# Memory / race checking of the hand-written kernels: runs a subset of the GPU tests under
# compute-sanitizer.  usage: tools/run_sanitizer.sh [memcheck|racecheck|synccheck|initcheck] [pytest -k expr]
set -euo pipefail
TOOL=${1:-memcheck}
EXPR=${2:-"fmha_forward or fmha_qkvpacked or bias_dropout_add_layer_norm or bias_gelu or layer_norm or softmax_cross_entropy or adam or embedding"}
cd "$(dirname "$0")/.."
# small shapes only: the sanitizer slows kernels down by 10-100x
compute-sanitizer --tool "$TOOL" --error-exitcode 42 --print-limit 20 \
    python -m pytest tests/test_gpu_kernels.py tests/test_gpu_attention.py -x -q -m gpu -k "$EXPR" -p no:cacheprovider
