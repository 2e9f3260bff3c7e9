timeout 200 python -m pytest tests/test_gpu_attention.py -m gpu -q --timeout 60 2>&1 | grep -E "AssertionError: \(|passed|failed|FAILED|Error" | head -40
