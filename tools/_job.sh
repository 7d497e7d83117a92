mkdir -p gpurun_out
timeout 60 python tools/tiny_chain.py bf16 128 8 > gpurun_out/s8g_tiny.log 2>&1 || { echo "tiny chain FAILED"; tail -5 gpurun_out/s8g_tiny.log; exit 1; }
timeout 200 python tools/trace_conv.py lidc > gpurun_out/s8g_trace_lidc.txt 2>&1; grep -A1 "^conv" gpurun_out/s8g_trace_lidc.txt | grep -v "^--" | cut -c1-230 | head -90
