#!/bin/bash
# GPU round trip used during development: parity tests, then the bench line (C2 headline + C3 secondary). Usage: gpurun -- 'bash tools/gpu_check.sh TAG'
TAG=${1:-dev}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || tail -20 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
def show(x,name):
    print(name, 'Mrays/s %.0f' % x['value'], 'ms %.2f' % x['ms_per_step'], 'e2e %.0f' % x['e2e']['value'], 'launches', x['gpu_launches'], 'roof %.3f' % x['roofline']['frac'], {k[:8]:round(v*x['ms_per_step'],2) for k,v in x['share_of_step'].items()})
show(d,'c2'); show(d['secondary'],'c3')
PY
