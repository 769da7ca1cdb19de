timeout 600 python -m pytest tests/test_gpu_generic_groups.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python tools/bench_generic_groups.py quick > gpurun_out/generic_groups.log 2>&1; grep -v '^\[' gpurun_out/generic_groups.log | cut -c1-200
