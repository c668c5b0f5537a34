#!/bin/bash
# compute-sanitizer passes over a small slice of the GPU suite (one tool per call: each is slow).
#   bash tools/sanitize.sh memcheck|racecheck|synccheck|initcheck
# The slice exercises every kernel family once at small sizes: products (dense, matrix-free,
# lattice table, sweep window), LU, QR, both Jacobi drivers, the direct solve, LSQR.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
TOOL=${1:-memcheck}
mkdir -p gpurun_out
SLICE='test_dense_apply[64-64-8] or test_kernelcov_apply[grid3-5-exponential] or test_grid_kernelcov_apply[grid3-spacing3-5-gaussian] or test_svd_small_fused_matches_per_round[60] or test_direct_solve_fused_matches_per_round[64] or test_device_lsqr_matches_oracle[20-5] or test_lu or test_qr'
timeout 900 compute-sanitizer --tool "$TOOL" --error-exitcode 86 --log-file gpurun_out/sanitizer_$TOOL.log \
    python -m pytest tests/test_gpu_blocks.py tests/test_gpu_pcga.py -m gpu -x -q -k "$SLICE" \
    > gpurun_out/sanitizer_${TOOL}_pytest.log 2>&1
echo "compute-sanitizer $TOOL rc=$?"
tail -5 gpurun_out/sanitizer_$TOOL.log
tail -3 gpurun_out/sanitizer_${TOOL}_pytest.log
