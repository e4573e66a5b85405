#!/bin/bash
# quick stall-reason comparison of library builds: ncu with a handful of metrics on the ICP pass and normals kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct
python tools/prof_step.py --pairs 16 > gpurun_out/q_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/q_plain.log; exit 1; }
for v in "$@"; do
  if [ "$v" = default ]; then unset B3D_LIB; else export B3D_LIB=$PWD/variants/libb200recon_$v.so; fi
  ncu --metrics $M --clock-control none -k regex:"icp_pass2_kernel|normals_cov2_kernel|icp_pass_kernel|normals_staged" -s 0 -c 4 --csv --log-file gpurun_out/q_$v.csv python tools/prof_step.py --pairs 16 > gpurun_out/q_$v.log 2>&1
  echo "== $v"; python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/q_$v.csv')) if len(r)>10]
hdr=rows[0]
ik,im,iv=hdr.index('Kernel Name'),hdr.index('Metric Name'),hdr.index('Metric Value')
iid=hdr.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[iid],r[ik][:40]),{})[r[im].split('__')[-1][:48]]=r[iv]
for k,v in d.items():
    print(k, ' '.join('%s=%s'%(a.replace('average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')[:22],b[:8]) for a,b in v.items()))
PY
done
