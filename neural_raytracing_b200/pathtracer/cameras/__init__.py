from .cameras import Camera, DTUCamera, NeRFCamera, lift
