python -m pytest tests/test_gpu_models.py -x -q -m gpu -k "bf16 or resnet12 or add_bias" 2>&1 | tail -3
python bench_configs.py --only C2,C4 --steps 10 2>&1 | tail -2
python bench_configs.py --only C1,C2,C4 --bf16 --steps 10 2>&1 | tail -3
