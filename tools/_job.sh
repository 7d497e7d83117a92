mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/s7a_pytest.log 2>&1; tail -6 gpurun_out/s7a_pytest.log | cut -c1-300
