python -m pytest tests/test_gpu_models.py -x -q -m gpu -k "conv1 or stem or repeat or conv64f" 2>&1 | tail -2
python tools/run_backbone_bf16.py 2>&1 | grep -E "stem|Conv64F"
