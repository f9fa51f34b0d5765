from .sdfs import SDF, SPHERE_SDF, SphereSDF
from .nerf import NeRFLE, PlainNeRF
from .shapes import Sphere
