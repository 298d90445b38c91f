timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -8 > gpurun_out/all_gpu1.log
tail -8 gpurun_out/all_gpu1.log
timeout 1500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sft3.json 2> gpurun_out/bench_sft3.err
tail -c 1300 gpurun_out/bench_sft3.json; tail -5 gpurun_out/bench_sft3.err
timeout 900 python bench.py --workload prefill --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_prefill2.json 2> gpurun_out/bench_prefill2.err
tail -c 1300 gpurun_out/bench_prefill2.json; tail -5 gpurun_out/bench_prefill2.err
