"""The shim package must expose our function under the name the reference imports (model/gaussian.py:8)."""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_import_line_resolves_to_our_rasterizer():
    code = ("from gsplat.rendering import rasterization; import easy_gaussian_splatting_b200 as e; "
            "assert rasterization is e.rasterization; print('ok')")
    env = {"PYTHONPATH": str(ROOT / "shim"), "PATH": "/usr/bin:/bin:/usr/local/cuda/bin"}
    import os
    env["PATH"] = os.environ.get("PATH", env["PATH"])
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp")
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr
