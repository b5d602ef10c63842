class TextureModder(object):
    """mujoco_py.modder.TextureModder: recolours textures for the viewer (fetch_env.py:305-315); a no-op here."""

    def __init__(self, sim):
        self.sim = sim

    def set_rgb(self, name, rgb):
        return None
