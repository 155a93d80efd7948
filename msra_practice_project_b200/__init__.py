"""B200-native volumetric ray-march render path behind the MSRA-practice-project render API.

Sub-modules
-----------
nerf_render   drop-in for the reference's ``nerf/render.py``
pigan_render  drop-in for the reference's ``pi_GAN/render.py``
models        NeRF / FilmSirenNeRF parameter containers (reference state-dict keys)
ops           autograd-aware wrappers over the C-ABI CUDA library (``libb2r.so``)
_lib          ctypes binding of ``include/b2r.h``
"""
__all__ = ["nerf_render", "pigan_render", "models", "ops"]
