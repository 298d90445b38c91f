timeout 900 python -X faulthandler -m pytest tests/test_backward_gpu.py -m gpu -q --no-header -p no:cacheprovider -s 2>&1 | grep -v "pooler layers\|^dq\|^dk\|^dv\|lora model.layers.[1-9]" | tail -80 > gpurun_out/bwd3.log
tail -80 gpurun_out/bwd3.log
