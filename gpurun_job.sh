timeout 900 python -X faulthandler -m pytest tests/test_backward_gpu.py -m gpu -q --no-header -p no:cacheprovider -s -k "lora or stepper" 2>&1 | grep -v "^  File" | tail -50 > gpurun_out/bwd4.log
tail -50 gpurun_out/bwd4.log
timeout 1500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sft2.json 2> gpurun_out/bench_sft2.err
tail -c 1500 gpurun_out/bench_sft2.json; tail -5 gpurun_out/bench_sft2.err
