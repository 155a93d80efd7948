"""B200-native volumetric ray-march render path behind the MSRA-practice-project render API.

Sub-modules
-----------
nerf_render   drop-in for the reference's ``nerf/render.py``
pigan_render  drop-in for the reference's ``pi_GAN/render.py``
models        NeRF / SirenNeRF / FilmSirenNeRF and the pi-GAN Generator / MappingNetwork / Renderer (reference state-dict keys)
train_step    NerfTrainStep (fused CUDA-graph training step) and RayBatcher (GPU-resident ray buffer, start-up crop sampler)
dist          ray / latent sharding helpers over torch.distributed (image gather, gradient all-reduce)
ops           autograd-aware wrappers over the C-ABI CUDA library (``libb2r.so``)
_lib          ctypes binding of ``include/b2r.h``
"""
__all__ = ["nerf_render", "pigan_render", "models", "ops"]
