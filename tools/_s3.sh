timeout 300 python tools/latency_parts.py > gpurun_out/s3_parts.log 2>&1
cat gpurun_out/s3_parts.log
