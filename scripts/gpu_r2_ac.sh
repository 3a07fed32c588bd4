# round 2, call AC (1 GPU): exchange jobs longest-first (A/B), the opt-in kernel-form tests
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/ac_build.log 2>&1; tail -n 2 gpurun_out/ac_build.log
run() { # name, env..., -- bench args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 5 "$@" > gpurun_out/ac_bench_$name.json 2> gpurun_out/ac_bench_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/ac_bench_%s.json'%n).read().strip().splitlines()[-1])
    print(n, round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,2), 'Gc/s frac', round(d['roofline']['whole_step']['frac'],3), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()}, 'probe', d['probe']['values'][:2])
except Exception as e:
    print(n, 'FAILED', e); print(open('gpurun_out/ac_bench_%s.err'%n).read()[-1500:])
PY
}
run sorted -- --n 512
run unsorted MEEP_B200_HALO_SORT=0 -- --n 512
run c4_sorted -- --workload c4
run c4_unsorted MEEP_B200_HALO_SORT=0 -- --workload c4
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "opt_in or unfused" > gpurun_out/ac_pytest_forms.log 2>&1; tail -n 3 gpurun_out/ac_pytest_forms.log
