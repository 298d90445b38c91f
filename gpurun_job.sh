timeout 900 python -X faulthandler -m pytest tests/test_generate_gpu.py tests/test_ops_gpu.py tests/test_backward_gpu.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -6
timeout 600 python bench.py --workload decode --steps 3 --warmup 3 > gpurun_out/bench_decode2.json 2> gpurun_out/bench_decode2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_decode2.json')); print('decode', d['value'], d['config']['ms_per_token'], d['config']['prefill_ms'], d['roofline']['frac'])"; tail -3 gpurun_out/bench_decode2.err
timeout 1500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sft6.json 2> gpurun_out/bench_sft6.err
python -c "
import json; d=json.load(open('gpurun_out/bench_sft6.json')); print('sft', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['launches_per_step'], d['roofline']['gemm_ms_per_step'], d['roofline']['attention_ms_per_step'])"; tail -3 gpurun_out/bench_sft6.err
