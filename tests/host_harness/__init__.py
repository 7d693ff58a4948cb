"""Builds and binds tests/host_harness/harness.cpp (test infrastructure, see the .cpp header)."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "libegs_host_harness.so"
SRC = HERE / "harness.cpp"
HDR = HERE.parent.parent / "easy_gaussian_splatting_b200" / "csrc" / "egs_math.cuh"


def load():
    LIB.parent.mkdir(exist_ok=True)
    if not LIB.exists() or LIB.stat().st_mtime < max(SRC.stat().st_mtime, HDR.stat().st_mtime):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                        "-x", "c++", str(SRC), "-o", str(LIB)], check=True)
    return ctypes.CDLL(str(LIB))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def projection_fwd(sc, sh_degree, eps2d=0.3, near=0.01, far=1e10, clip=0.0, tile=16):
    """sc: a synthetic.Scene (CPU tensors) -> dict of numpy arrays."""
    lib = load()
    C, N = sc.viewmats.shape[0], sc.means.shape[0]
    K = sc.colors.shape[1]
    tw, th = -(-sc.width // tile), -(-sc.height // tile)
    f = lambda t: np.ascontiguousarray(t.detach().numpy(), dtype=np.float32)
    out = dict(radii=np.zeros((C, N), np.int32), means2d=np.zeros((C, N, 2), np.float32),
               depths=np.zeros((C, N), np.float32), conics=np.zeros((C, N, 3), np.float32),
               colors=np.zeros((C, N, 3), np.float32), tiles_per_gauss=np.zeros((C, N), np.int32),
               compensations=np.zeros((C, N), np.float32))
    arrs = [f(sc.means), f(sc.quats), f(sc.scales), f(sc.colors), f(sc.viewmats), f(sc.Ks)]
    lib.hh_projection_fwd(C, N, _p(arrs[0]), _p(arrs[1]), _p(arrs[2]), _p(arrs[3]), K, sh_degree, _p(arrs[4]),
                          _p(arrs[5]), sc.width, sc.height, ctypes.c_float(eps2d), ctypes.c_float(near),
                          ctypes.c_float(far), ctypes.c_float(clip), tile, tw, th, _p(out["radii"]),
                          _p(out["means2d"]), _p(out["depths"]), _p(out["conics"]), _p(out["colors"]),
                          _p(out["tiles_per_gauss"]), _p(out["compensations"]))
    return out


def projection_bwd(sc, sh_degree, radii, colors, v_means2d, v_conics, v_colors, eps2d=0.3, v_comps=None):
    lib = load()
    C, N = sc.viewmats.shape[0], sc.means.shape[0]
    K = sc.colors.shape[1]
    f = lambda t: np.ascontiguousarray(t.detach().numpy() if hasattr(t, "detach") else t, dtype=np.float32)
    arrs = [f(sc.means), f(sc.quats), f(sc.scales), f(sc.colors), f(sc.viewmats), f(sc.Ks), f(colors), f(v_means2d),
            f(v_conics), f(v_colors)]
    radii = np.ascontiguousarray(radii, dtype=np.int32)
    vc = None if v_comps is None else f(v_comps)
    out = dict(v_means=np.zeros((N, 3), np.float32), v_quats=np.zeros((N, 4), np.float32),
               v_scales=np.zeros((N, 3), np.float32), v_sh=np.zeros((N, K, 3), np.float32))
    lib.hh_projection_bwd(C, N, _p(arrs[0]), _p(arrs[1]), _p(arrs[2]), _p(arrs[3]), K, sh_degree, _p(arrs[4]),
                          _p(arrs[5]), sc.width, sc.height, ctypes.c_float(eps2d), _p(radii), _p(arrs[6]),
                          _p(arrs[7]), _p(arrs[8]), _p(arrs[9]), _p(out["v_means"]), _p(out["v_quats"]),
                          _p(out["v_scales"]), _p(out["v_sh"]), None if v_comps is None else _p(vc))
    return out


def tight_rects(means2d, radii, conics, opacities, tw, th, tile=16):
    """-> (classic [n,4] = x0 y0 x1 y1 in tiles, packed [n,2] = the 8 bytes the projection kernel hands to the binning)"""
    lib = load()
    n = radii.shape[0]
    m2 = np.ascontiguousarray(means2d, dtype=np.float32)
    rd = np.ascontiguousarray(radii, dtype=np.int32)
    cn = np.ascontiguousarray(conics, dtype=np.float32)
    op = np.ascontiguousarray(opacities, dtype=np.float32)
    classic = np.zeros((n, 4), dtype=np.int32)
    packed = np.zeros((n, 2), dtype=np.int32)
    lib.hh_tight_rects(n, _p(m2), _p(rd), _p(cn), _p(op), tile, tw, th, _p(classic), _p(packed))
    return classic, packed
