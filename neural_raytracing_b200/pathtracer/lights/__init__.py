from .lights import Light, LightField, PointLights
