def load_model_from_path(path):
    raise RuntimeError("mujoco_py shim: MuJoCo 1.50 is not available in this image")


class MjSim(object):
    def __init__(self, *a, **k):
        raise RuntimeError("mujoco_py shim: MuJoCo 1.50 is not available in this image")
