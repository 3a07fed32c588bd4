# round 2, call N (1 GPU): final-form ncu captures, exported to text on the box (reports over 64 MB do not travel)
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/n_build.log 2>&1; tail -n 2 gpurun_out/n_build.log
cap() { name=$1; regex=$2; shift 2
  pre=(); while [ "$1" = "-s" ] || [ "$1" = "-c" ]; do pre+=("$1" "$2"); shift 2; done
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$regex "${pre[@]}" -o /tmp/n_$name "$@" > gpurun_out/n_ncu_$name.log 2>&1
  ncu -i /tmp/n_$name.ncu-rep --page details > gpurun_out/n_${name}_ncu_details.txt 2>/dev/null
  ncu -i /tmp/n_$name.ncu-rep --page raw --csv > gpurun_out/n_${name}_ncu_raw.csv 2>/dev/null
  ls -la /tmp/n_$name.ncu-rep
}
cap step3c_512 step3c -s 8 -c 2 python bench.py --size 512 --steps 2 --warmup 4 --no-cpu-baseline
cap halo_512 halo_kernel -s 8 -c 2 python bench.py --size 512 --steps 2 --warmup 4 --no-cpu-baseline
cap dft_c3 dft_kernel -c 2 python bench.py --workload c3 --steps 30 --warmup 3 --no-cpu-baseline
cap flux flux_ -c 2 python -m pytest tests/test_kernels_gpu.py -q -k "flux"
cap plain_1024 step3_plain -s 10 -c 2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline
cp /tmp/n_plain_1024.ncu-rep gpurun_out/n_prof_plain_1024.ncu-rep
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/n_launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/n_ncu_launch_c3.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/n_launches_c3.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-60:]:
    if 'halo' in r[4] or 'zero' in r[4]: print(r[4][:40], r[-4:], )
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/n_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/n_ncu_launch_default.log 2>&1
run() { name=$1; shift
  env $ENVV timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/n_bench_$name.json 2> gpurun_out/n_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/n_bench_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,2), 'Gc/s frac', round(d['roofline']['whole_step']['frac'],3))
    print('   ', {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/n_bench_$name.err').read()[-1500:])
PY
}
ENVV="X=1" run 512_f32 --size 512 --prec f32
ENVV="X=1" run 512 --size 512
ENVV="X=1" run c3 --workload c3 --steps 40
timeout 600 python -m pytest tests -m gpu -q -x -k "c3_au or lorentz or gyro or noisy or polariton" > gpurun_out/n_pytest.log 2>&1
tail -n 3 gpurun_out/n_pytest.log
du -sh gpurun_out
