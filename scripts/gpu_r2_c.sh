# round 2, call C: new fast-path kernel form (A/B), deferred same-device D/B halos (A/B), flux tree,
# 1024^3 on ONE GPU (host RSS, set-up time), launch list of the default bench
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/c_build.log 2>&1; tail -n 2 gpurun_out/c_build.log
timeout 900 python -m pytest tests -m gpu -q -x -k "kernels or c2_3d_pml or 3d_metal or 3d_bloch or lorentz_3d or c3_au or 2d_bend or dft_fields or sync_magnetic or midrun or cyl_m1_flux or flux" > gpurun_out/c_pytest.log 2>&1
tail -n 5 gpurun_out/c_pytest.log
run() { name=$1; shift
  env $ENVV timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/c_bench_$name.json 2> gpurun_out/c_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/c_bench_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['ms_per_step'],3), 'ms', round(d['value']/1e9,2), 'Gc/s frac', round(d['roofline']['whole_step']['frac'],3), 'setup', round(d['config']['setup_s'],1), 'warm', round(d['config']['warmup_s'],1), 'rss', round(d['config']['host_max_rss_gb'],1), 'e2e', round(d['e2e']['value']/1e9,2))
    print('   ', {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/c_bench_$name.err').read()[-1500:])
PY
}
ENVV="X=1" run 512_default
ENVV="MEEP_B200_PLAIN_PER_JOB=0" run 512_tablejob
ENVV="MEEP_B200_DEFER_LOCAL=0" run 512_nodefer
ENVV="X=1" run 512_f32 --prec f32
ENVV="MEEP_B200_VERBOSE=1 MEEP_B200_BENCH_N1=1024" run 1024_1gpu --steps 10 --warmup 3
grep -v "no E/H fusion\|recorded\|phase " gpurun_out/c_bench_1024_1gpu.err | cut -c1-200 | tail -n 30
ENVV="X=1" run c3 --workload c3
ENVV="X=1" run c4 --workload c4
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c_launches_c2_512.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c_ncu_launch.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/c_launches_c2_512.csv')) if len(r)>10 and r[0].isdigit()]
agg={}
for r in rows[-160:]:
    k=r[4][:60]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[-1])
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(k, v[0], round(v[1]/1e3,1),'us')
PY
