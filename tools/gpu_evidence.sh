#!/bin/bash
# One-GPU evidence run for profiles/: TAG=<prefix> bash tools/gpu_evidence.sh [full]
#   1. the plain bench (all legs)                                   -> gpurun_out/${TAG}_bench_p64.json
#   2. ncu launch list of one 64-pair step (every kernel)           -> gpurun_out/${TAG}_launches_p64.csv
#   3. per-pass counters of the ICP pass (stats build)               -> gpurun_out/${TAG}_icp_pass_probe.txt
#   4. (full) ncu --set full of the ICP passes and the normals kernel of one 64-pair step -> gpurun_out/${TAG}_*.ncu-rep
# Every command runs under its own timeout; a number printed under ncu is never a bench value.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-r02x}
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_p64.json 2> gpurun_out/${TAG}_bench_p64.err
echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/${TAG}_bench_p64.json'))
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['pipeline_roofline']['frac'])
    print({x['name']: round(x['ms_per_step'], 2) for x in d['kernels'][:8]})
    print({k: (v.get('ms_per_pair') or v.get('ms') or list(v)[:4]) if isinstance(v, dict) else v for k, v in d.get('extra', {}).items()})
except Exception as e:
    print('no bench json', e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches_p64.csv \
    python tools/prof_step.py --pairs 64 --unique 8 > gpurun_out/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
python profiles/summarize_launches.py gpurun_out/${TAG}_launches_p64.csv > gpurun_out/${TAG}_launches_p64_summary.txt 2>&1; head -12 gpurun_out/${TAG}_launches_p64_summary.txt
if [ -f variants/libb200recon_stats.so ]; then
  B3D_LIB=$PWD/variants/libb200recon_stats.so timeout 300 python profiles/icp_pass_probe.py > gpurun_out/${TAG}_icp_pass_probe.txt 2>&1
  echo "probe rc=$?"; tail -14 gpurun_out/${TAG}_icp_pass_probe.txt
fi
if [ "$1" = full ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:icp_pass2_kernel -c 10 -o gpurun_out/${TAG}_icp_pass2_p64 -f \
      python tools/prof_step.py --pairs 64 --unique 8 > gpurun_out/${TAG}_ncu_icp.log 2>&1
  echo "ncu icp rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"normals_cov2_kernel|rs_onesweep_kernel" -c 4 -o gpurun_out/${TAG}_normals_sort_p64 -f \
      python tools/prof_step.py --pairs 64 --unique 8 > gpurun_out/${TAG}_ncu_nrm.log 2>&1
  echo "ncu normals rc=$?"
fi
