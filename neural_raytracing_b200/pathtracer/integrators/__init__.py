from .integrators import Debug, Direct, Integrator, Mask, NeRFIntegrator, NeRFReproduce, Path, Silhouette
