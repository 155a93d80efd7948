"""Loader of the UNMODIFIED reference render path (TEST / BASELINE INFRASTRUCTURE ONLY).

Imports the reference's own files -- nerf/render.py, nerf/nerf.py, pi_GAN/render.py, pi_GAN/modules.py, pi_GAN/utils.py -- from
``oracle/_ref/`` (a git-ignored verbatim copy made by tools/install_ref.sh, which travels to the GPU box) or, in the build
container, straight from /root/reference.  Used by bench.py's CPU arm (``--impl reference`` / ``cpu_baseline``, kind
"reference") and by the golden-fixture generators; the product never imports this module.

The reference directories are not packages and both hold a ``render.py`` / flat imports (``from render import *``), so every
file is loaded under a distinct module name with its own directory temporarily at the head of sys.path; matplotlib, imageio,
plyfile and skimage (imported by pi_GAN/modules.py:4 and utils.py:3-7, never used on the path) are stubbed when absent.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_CACHE: dict = {}


def reference_root() -> str | None:
    for root in (os.path.join(_HERE, "_ref"), "/root/reference"):
        if os.path.exists(os.path.join(root, "nerf", "render.py")) and os.path.exists(os.path.join(root, "pi_GAN", "modules.py")):
            return root
    return None


def _load(name: str, path: str, flat_dir: str):
    saved_path, saved_mods = list(sys.path), {k: sys.modules.get(k) for k in ("render", "modules", "utils", "nerf")}
    for k in saved_mods:
        sys.modules.pop(k, None)
    sys.path.insert(0, flat_dir)
    try:
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path[:] = saved_path
        for k, v in saved_mods.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
    return mod


def load_reference() -> dict | None:
    """{'root', 'nerf_render', 'nerf_nerf', 'pigan_render', 'pigan_modules', 'pigan_utils'} or None when no copy of the
    reference is reachable.  Importing nerf/nerf.py switches autograd anomaly mode on (nerf/nerf.py:2); it is left as shipped."""
    if _CACHE:
        return _CACHE
    root = reference_root()
    if root is None:
        return None
    for n in ("matplotlib", "matplotlib.pyplot", "imageio", "plyfile", "skimage", "skimage.measure"):
        try:
            __import__(n)
        except Exception:
            sys.modules.setdefault(n, types.ModuleType(n))
    nd, pd = os.path.join(root, "nerf"), os.path.join(root, "pi_GAN")
    out = {"root": root}
    out["nerf_render"] = _load("b2r_ref_nerf_render", os.path.join(nd, "render.py"), nd)
    out["nerf_nerf"] = _load("b2r_ref_nerf_nerf", os.path.join(nd, "nerf.py"), nd)
    out["pigan_render"] = _load("b2r_ref_pigan_render", os.path.join(pd, "render.py"), pd)
    out["pigan_modules"] = _load("b2r_ref_pigan_modules", os.path.join(pd, "modules.py"), pd)
    try:
        out["pigan_utils"] = _load("b2r_ref_pigan_utils", os.path.join(pd, "utils.py"), pd)
    except Exception:                                            # utils.py is only needed for create_mesh's loop
        out["pigan_utils"] = None
    _CACHE.update(out)
    return _CACHE
