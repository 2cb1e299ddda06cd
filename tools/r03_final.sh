# round-end evidence (1 GPU): parity suite, both bench arms, launch list + full ncu capture of the dominant kernel on C3
T=${1:-r03z}; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/${T}_pytest.txt; tail -2 $O/${T}_pytest.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
timeout 500 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
export SAILOR_PT_TRAVERSAL=local
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches_c3.csv python tools/profile_step.py c3 > $O/${T}_launches_c3.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_trace_fast_level" -s 0 -c 2 -f -o $O/${T}_c3 python tools/profile_step.py c3 > $O/${T}_full_c3.log 2>&1
python - <<PY
import json
d=json.loads(open('$O/${T}_bench.json').read().strip().splitlines()[-1])
print('c3 Mrays/s %.0f ms %.2f e2e %.0f frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']), {k[:9]:round(v*d['ms_per_step'],2) for k,v in d['share_of_step'].items()})
s=d['secondary']; print('c2 Mrays/s %.0f ms %.2f e2e %.0f' % (s['value'], s['ms_per_step'], s['e2e']['value']))
r=json.loads(open('$O/${T}_bench_reference.json').read().strip().splitlines()[-1]); print('reference', r.get('value'), r.get('cpu_baseline',{}).get('cores'))
PY
