"""Host camera, mirroring ``camera::Camera`` of the reference (src/camera.rs:5-79) through the C ABI."""
import ctypes as C

from . import _types as T
from ._lib import lib


class Camera:
    """``Camera::new((w, h), fov, position)`` (camera.rs:31-52).  ``update()`` is camera.rs:54-79, including
    its one-call lag: the basis is derived from the previous look-at point, then the look-at is refreshed
    from ``theta_x`` / ``theta_y``."""

    def __init__(self, size, fov, position):
        self.c = T.HostCamera()
        pos = (C.c_float * 3)(*[float(p) for p in position])
        lib().ptc_camera_new(C.byref(self.c), int(size[0]), int(size[1]), float(fov), pos)

    def update(self):
        lib().ptc_camera_update(C.byref(self.c))

    def __getattr__(self, name):
        c = object.__getattribute__(self, "c")
        v = getattr(c, name)
        return list(v) if hasattr(v, "__len__") else v

    def set(self, **kw):
        for k, v in kw.items():
            if isinstance(v, (list, tuple)):
                arr = getattr(self.c, k)
                for i, x in enumerate(v):
                    arr[i] = float(x)
            else:
                setattr(self.c, k, v)
        return self

    def look_at(self, target):
        """Convenience (not in the reference, which only steers with the mouse): choose theta so that the
        steady-state basis looks at ``target``; runs update() twice to flush the one-call lag."""
        import math
        d = [float(target[i]) - self.c.position[i] for i in range(3)]
        n = math.sqrt(sum(x * x for x in d))
        d = [x / n for x in d]
        self.c.theta_y = math.asin(max(-1.0, min(1.0, d[1])))
        self.c.theta_x = math.atan2(d[2], d[0])
        self.update()
        self.update()
        return self

    def uniforms(self, frame=None, is_grabbed=0):
        """``uniforms::Uniforms::new(&camera)`` (uniforms.rs:64-71); ``frame`` is the RNG stream id
        (raytracer.rs:180 increments it before every path-traced frame, so the first frame is 1)."""
        if frame is not None:
            self.c.frame = int(frame)
        u = T.Uniforms()
        lib().ptc_camera_uniforms(C.byref(self.c), C.byref(u))
        u.is_grabbed = int(is_grabbed)
        return u
