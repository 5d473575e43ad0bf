"""zest_nerf_b200 -- B200-native per-ray rendering hot path behind ZeST-NeRF's API.

Public surface (mirrors the reference modules for this path only):
  renderer.rendering(...)          drop-in for `renderer.py:579-626`
  networks.{Embedding,Renderer,MVSNeRF}   drop-in for `networks.py:29-65,73-221,321-353`
  driver.FrameRenderer             ray-sharded full-frame / multi-GPU driver
"""
__version__ = "0.1.0"
