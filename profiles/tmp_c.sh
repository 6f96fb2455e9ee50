timeout -s KILL 600 python -m pytest tests/test_epoch_gpu.py -q -m gpu -x 2>&1 | tail -5
