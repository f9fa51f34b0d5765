from .integrators import (BasisBRDF, Debug, Depth, Direct, Illumination, Integrator, Luminance, Mask, NeRFIntegrator,
                          NeRFReproduce, Path, Silhouette)
