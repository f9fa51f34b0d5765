from .bsdfs import BSDF, BSDFSample, ComposeSpatialVarying, Conductor, Diffuse, NeuralBSDF, fresnel_conductor
