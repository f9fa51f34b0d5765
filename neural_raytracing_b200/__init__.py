"""neural_raytracing_b200 -- B200 (sm_100a) native per-ray hot path of
prashantraina/neural_raytracing behind the reference's `pytorch3d.pathtracer` call surface.

  neural_raytracing_b200.pathtracer   drop-in mirror of pytorch3d.pathtracer (classes, pathtrace)
  neural_raytracing_b200.ops          functional ops on CUDA tensors -> libnrt_b200.so (C ABI)
"""
__version__ = "0.1.0"
