timeout 900 python -X faulthandler -m pytest tests/test_generate_gpu.py -m gpu -q --no-header -p no:cacheprovider -s 2>&1 | grep -v "^  File\|^$" | tail -60 > gpurun_out/gen1.log
tail -60 gpurun_out/gen1.log
