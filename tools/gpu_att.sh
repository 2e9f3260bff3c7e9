timeout 200 python -m pytest tests/test_gpu_attention.py -m gpu -q --timeout 60 2>&1 | grep -E "AssertionError: \(|AssertionError: tensor|passed|failed|FAILED|Error" | head -30
timeout 200 python tools/bench_attention.py 2>&1 | tail -7
