"""CPU test (-m "not gpu"): every learned component of the path has the reference's parameter names and shapes, so
`load_state_dict` of a reference checkpoint (dtu.py:93-108 saves these modules) works on the mirror (SURVEY.md
section 8f rank 3).  Fixture: tests/golden/plain_nerf.npz (`sd_*`), written from the unmodified reference."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

G = np.load(os.path.join(HERE, "golden", "plain_nerf.npz"))


def _build(name):
    from neural_raytracing_b200.pathtracer.shapes.sdfs import SphereSDF
    from neural_raytracing_b200.pathtracer.shapes.nerf import NeRFLE, PlainNeRF
    from neural_raytracing_b200.pathtracer.bsdf import ComposeSpatialVarying, Diffuse, NeuralBSDF
    from neural_raytracing_b200.pathtracer.lights import LightField
    return {"SphereSDF": lambda: SphereSDF(n=64, device="cpu"), "NeRFLE": lambda: NeRFLE(device="cpu"),
            "NeRFLE_envmap": lambda: NeRFLE(envmap=True, device="cpu"), "PlainNeRF": lambda: PlainNeRF(device="cpu"),
            "NeuralBSDF": lambda: NeuralBSDF(device="cpu"), "LightField": lambda: LightField(device="cpu"),
            "ComposeSpatialVarying": lambda: ComposeSpatialVarying([NeuralBSDF(device="cpu"), Diffuse(device="cpu")],
                                                                   device="cpu")}[name]()


@pytest.mark.parametrize("name", ["SphereSDF", "NeRFLE", "NeRFLE_envmap", "PlainNeRF", "NeuralBSDF", "LightField",
                                  "ComposeSpatialVarying"])
def test_state_dict_layout_is_the_reference(name):
    ours = ["%s:%s" % (k, "x".join(str(d) for d in v.shape)) for k, v in _build(name).state_dict().items()]
    assert ours == [str(s) for s in G["sd_" + name]]
