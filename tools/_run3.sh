timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python tools/bench_generic_groups.py quick > gpurun_out/generic_groups.log 2>&1; grep -v '^\[' gpurun_out/generic_groups.log | grep "auto" | cut -c1-200
timeout 600 python tools/bench_multivariate.py > gpurun_out/multivariate.log 2>&1; grep -v '^\[' gpurun_out/multivariate.log | cut -c1-220
timeout 600 python tools/bench_wide_state.py > gpurun_out/wide_state.log 2>&1; cat gpurun_out/wide_state.log | cut -c1-220
