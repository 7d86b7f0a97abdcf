#!/bin/bash
# tools/sanitize.sh -- SURVEY.md section 5 tooling, run on the GPU box (gpurun -- 'bash tools/sanitize.sh'); logs -> gpurun_out/ (copy
# the summaries into profiles/).
#   1. compute-sanitizer --tool memcheck   over the small-configuration GPU tests (every kernel, out-of-bounds / misaligned accesses)
#   2. compute-sanitizer --tool racecheck  over the tests whose kernels share data through shared memory: the 3-lane team pairing
#      (__syncwarp between phases), the TMA-staged gather (mbarrier + cp.async.bulk), the vote scatter's shared bins, the tree kernel
#   3. compute-sanitizer --tool synccheck  on the same subset (illegal barrier use)
#   4. the host side of the C ABI under AddressSanitizer + UBSan (pos_evolution_b200/libb200pos_asan.so, built by
#      `python -m pos_evolution_b200.build --asan`) over the BLS / fork-choice / guard tests
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
SMALL="tests/test_gpu_bls.py tests/test_gpu_forkchoice.py tests/test_gpu_gather.py tests/test_gpu_participation.py tests/test_gpu_epoch.py"
RACE="tests/test_gpu_bls.py::test_fast_aggregate_verify_vs_oracle tests/test_gpu_bls.py::test_g1_aggregate_all_minimal_committees tests/test_gpu_forkchoice.py tests/test_gpu_gather.py"
run() {  # name, timeout, command...
    local name=$1 tmo=$2; shift 2
    echo "== $name"
    timeout "$tmo" "$@" > "gpurun_out/sanitize_$name.log" 2>&1
    echo "rc=$? ($name)"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" "gpurun_out/sanitize_$name.log" | tail -6
}
run memcheck 600 $SAN --tool memcheck --error-exitcode 1 --launch-timeout 600 python -m pytest $SMALL -x -q -m gpu -k "not test_epoch_pipeline or sync-3 or pipelined-3 or pipelined_rlc or pipelined_team"
run racecheck 420 $SAN --tool racecheck --racecheck-report all --error-exitcode 1 python -m pytest $RACE -x -q -m gpu
run synccheck 300 $SAN --tool synccheck --error-exitcode 1 python -m pytest $RACE -x -q -m gpu
if [ -f pos_evolution_b200/libb200pos_asan.so ]; then
    ASAN_LIB=$(gcc -print-file-name=libasan.so)
    B2_LIB=$PWD/pos_evolution_b200/libb200pos_asan.so LD_PRELOAD=$ASAN_LIB ASAN_OPTIONS=protect_shadow_gap=0:detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1 \
        run asan_host 300 python -m pytest tests/test_gpu_bls.py tests/test_gpu_forkchoice.py tests/test_gpu_gather.py tests/test_gpu_participation.py -x -q -m gpu
else
    echo "no ASan build (python -m pos_evolution_b200.build --asan)"
fi
