#!/bin/bash
# bench (no extra legs, 1-pair CPU sample) of the default library (`default`), of variants/libb200recon_<name>.so (`<name>`) and of the
# default library under an environment variable (`env:NAME=VALUE`), one line each
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --cpu-sample 1 --no-extra > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.load(open('gpurun_out/ab_$name.json'))
    k={x['name'].replace('_kernel',''):round(x['ms_per_step'],2) for x in d['kernels'][:6]}
    print(round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), k)
except Exception as e:
    print('no json', e)
PY
)"
}
for v in "$@"; do
  case $v in
    default) run default B3D_DUMMY=1 ;;
    env:*) kv=${v#env:}; run "$(echo $kv | tr -c 'A-Za-z0-9\n' '_')" $kv ;;   # env:NAME=VALUE -> the default library with that variable set
    *) run $v B3D_LIB=$PWD/variants/libb200recon_$v.so ;;
  esac
done
