timeout 600 python -m pytest tests/test_gpu_rasterization.py -m gpu -q -p no:cacheprovider -x -k "chunked" 2>&1 | grep -v "^$" | tail -30
