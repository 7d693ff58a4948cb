#!/bin/bash
# Round 2, call AA: vectorised fused Adam + step phases; stage tests, then the 1-GPU bench line
T=${1:-r2aa}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stages.py -m gpu -q -x --timeout 600 --timeout-method=thread -p no:cacheprovider -rf > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
timeout 900 python bench.py --no-call-pattern --no-cpu-baseline > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/${T}_bench.err
python scripts/show_bench.py gpurun_out/${T}_bench.log 2>/dev/null | head -8 | cut -c1-300
python - <<PY
import time, torch, sys
sys.path.insert(0, '.')
from easy_gaussian_splatting_b200.optim import FusedAdam
N = 3_000_000
ps = [torch.randn(N, *s, device='cuda').requires_grad_(True) for s in ((3,), (3,), (4,), (1, 3), (15, 3), ())]
opt = FusedAdam([{"params": [p], "lr": 1e-3} for p in ps])
for p in ps: p.grad = torch.randn_like(p)
for _ in range(3): opt.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): opt.step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
b = sum(p.numel() for p in ps) * 28
print(f"fused Adam, 3 M Gaussians (177 M parameters): {ms:.3f} ms per step = {b / ms / 1e9:.2f} TB/s")
PY
