#!/bin/bash
# A/B run on one B200: GPU parity tests with the default (round-2) kernels, the same tests against a build with tiny staging
# buffers (multi-batch and cell-by-cell paths), then the bench with each round-2 kernel switched back to its round-1 version.
# Every command runs under its own timeout; outputs land in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/ab_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/ab_tests.log
if [ -f variants/libb200recon_cap60.so ]; then
  B3D_LIB=$PWD/variants/libb200recon_cap60.so timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/ab_tests_cap60.log 2>&1
  echo "tests(cap60) rc=$?"; tail -3 gpurun_out/ab_tests_cap60.log
fi
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --cpu-sample 1 --no-extra > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.load(open('gpurun_out/ab_$name.json'))
    k={x['name']:round(x['ms_per_step'],2) for x in d['kernels'][:9]}
    print(round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), k)
except Exception as e:
    print('no json', e)
PY
)"
}
for v in "$@"; do
  case $v in
    new) run new B3D_DUMMY=1 ;;
    stats) B3D_LIB=$PWD/variants/libb200recon_stats.so B3D_ICP_STATS=1 timeout 300 python tools/prof_step.py --pairs 16 --stats > gpurun_out/ab_stats.log 2>&1; tail -4 gpurun_out/ab_stats.log ;;
    icp_v1) run icp_v1 B3D_ICP_V1=1 ;;
    nrm_v1) run nrm_v1 B3D_NRM_V1=1 ;;
    all_v1) run all_v1 B3D_ICP_V1=1 B3D_NRM_V1=1 B3D_SORT_PAIRS=1 ;;
    lib:*) run ${v#lib:} B3D_LIB=$PWD/variants/libb200recon_${v#lib:}.so ;;
  esac
done
