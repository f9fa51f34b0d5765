from .integrators import Debug, Direct, Integrator, Mask, NeRFIntegrator, NeRFReproduce, Path, Silhouette
from .vis import BasisBRDF, Depth, Illumination, Luminance
