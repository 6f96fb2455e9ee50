timeout -s KILL 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -2
python profiles/host_overhead.py 2>&1 | tail -40
