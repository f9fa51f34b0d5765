from .bsdfs import BSDF, ComposeSpatialVarying, Conductor, Diffuse, NeuralBSDF, fresnel_conductor
