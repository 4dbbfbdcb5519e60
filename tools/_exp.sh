python -m pytest tests/test_gpu_models.py -x -q -m gpu -k "conv3x3 or conv64f or bf16 or repeat" 2>&1 | tail -3
echo "conv3 default (bf16: two epilogue groups)"; python tools/run_backbone_bf16.py 2>&1 | grep -E "block|Conv64F"
echo "conv3 EPI2=0"; AFS_CONV3_EPI2=0 python tools/run_backbone_bf16.py 2>&1 | grep -E "block|Conv64F"
