# first GPU round trip of the origin-local kernel: parity, A/B against the exact kernel on C3, loop statistics
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${1}_pytest.txt
timeout 600 python tools/ab_trace.py c3 3 exact,fast > gpurun_out/${1}_ab_c3.txt 2>&1
timeout 600 python tools/fast_variants.py run c3 > gpurun_out/${1}_variants.txt 2>&1
tail -5 gpurun_out/${1}_pytest.txt; cat gpurun_out/${1}_ab_c3.txt; cat gpurun_out/${1}_variants.txt
