#!/bin/bash
# Round-end evidence run (1 GPU): parity tests, both bench arms, ncu launch lists of one C2 and one C3 frame, and full captures of the
# dominant kernels (C3 k_trace_level, C2 FanOutKernel) for roofline.traffic.  Usage: gpurun -- 'bash tools/final_capture.sh TAG'
TAG=${1:-final}
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; tail -2 $O/${TAG}_pytest.log
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
for w in c2 c3; do
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_$w.csv python tools/profile_step.py $w > $O/${TAG}_launches_$w.log 2>&1
done
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_trace_level" -s 2 -c 4 -f -o $O/${TAG}_c3_trace python tools/profile_step.py c3 > $O/${TAG}_c3_trace.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_for_range|k_trace_rays_small|k_primary_pass" -c 8 -f -o $O/${TAG}_c2_frame python tools/profile_step.py c2 > $O/${TAG}_c2_frame.log 2>&1
python - <<PY
import json
d=json.load(open('$O/${TAG}_bench.json'))
def show(x,name):
    print(name, 'Mrays/s %.0f' % x['value'], 'ms %.2f' % x['ms_per_step'], 'e2e %.0f' % x['e2e']['value'], 'launches', x['gpu_launches'], 'roof %.3f' % x['roofline']['frac'], {k[:8]:round(v*x['ms_per_step'],2) for k,v in x['share_of_step'].items()})
show(d,'c2'); show(d['secondary'],'c3')
r=json.load(open('$O/${TAG}_bench_reference.json')); print('reference', r.get('value'), r.get('cpu_baseline'))
PY
