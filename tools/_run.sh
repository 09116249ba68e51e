python -m pytest tests -m gpu -x -q -k "sketch_sparse" 2>&1 | tail -8 > gpurun_out/t2_tests.log
python tools/exp_sksp.py > gpurun_out/t2_exp.log 2>&1
tail -3 gpurun_out/t2_tests.log; cat gpurun_out/t2_exp.log
