timeout 900 python -X faulthandler -m pytest tests/test_backward_gpu.py tests/test_modules_gpu.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -6 > gpurun_out/bwd5.log
tail -6 gpurun_out/bwd5.log
timeout 1500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sft4.json 2> gpurun_out/bench_sft4.err
python -c "
import json; d=json.load(open('gpurun_out/bench_sft4.json')); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['launches_per_step'], d['roofline']['gemm_ms_per_step'], d['roofline']['attention_ms_per_step'])"; tail -3 gpurun_out/bench_sft4.err
