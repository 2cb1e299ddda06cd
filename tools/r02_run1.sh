set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02a_pytest.txt
python tools/ab_wide.py c3 3 > gpurun_out/r02a_ab_c3.txt 2>&1
python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -5 gpurun_out/r02a_pytest.txt; cat gpurun_out/r02a_ab_c3.txt; cat gpurun_out/r02a_bench.json; tail -5 gpurun_out/r02a_bench.err
