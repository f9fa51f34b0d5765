"""Mirror of the pieces of `pytorch3d.renderer` that the path tracer's scripts import (cameras, and the fork's
path-tracing point light)."""
from .cameras import (FoVPerspectiveCameras, OpenGLPerspectiveCameras, camera_position_from_spherical_angles,  # noqa: F401
                      look_at_rotation, look_at_view_transform)
from .lighting import PointLights  # noqa: F401
