"""Scene container, mirroring ``scene::Scene`` of the reference (src/scene.rs:65-394).

The arrays are exactly the ones the reference uploads (``vertices``, global ``indices``, ``meshes``,
``materials``, ``textures``); OBJ/MTL ingest is done by the C++ loader in csrc/pt_host.cpp through the C ABI,
texture *files* are decoded here with OpenCV (the reference uses the ``image`` crate, texture.rs:232-243).
"""
import ctypes as C
import os

import numpy as np

from . import _types as T
from ._lib import check, lib


class Texture:
    """Decoded texels (H x W x 4, uint8 or float32).  ``texture::Texture`` after ``load()`` minus the BC7
    step (texture.rs:255-273), which this core deliberately does not reproduce."""

    def __init__(self, texels, path=None):
        a = np.ascontiguousarray(texels)
        if a.ndim != 3 or a.shape[2] != 4 or a.dtype not in (np.uint8, np.float32):
            raise ValueError("texture texels must be HxWx4 uint8 or float32")
        self.texels = a
        self.path = path

    @staticmethod
    def from_color(rgba):
        return Texture(np.asarray(rgba, dtype=np.float32).reshape(1, 1, 4))

    @staticmethod
    def from_scalar(v):
        return Texture(np.asarray([v, 0.0, 0.0, 1.0], dtype=np.float32).reshape(1, 1, 4))

    @property
    def format(self):
        return T.TEXEL_RGBA8 if self.texels.dtype == np.uint8 else T.TEXEL_RGBA32F

    @property
    def size(self):
        return self.texels.shape[1], self.texels.shape[0]


def decode_image(path):
    """``image::open(path).to_rgba32f()`` semantics (texture.rs:239-243): 8-bit -> kept as RGBA8 (the kernel
    divides by 255, no sRGB decode), 16-bit -> value/65535 as float, float (EXR/HDR) -> as is; grey -> rgb,
    missing alpha -> 1."""
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    import cv2
    im = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if im is None:
        raise IOError(f"cannot decode texture {path}")
    if im.ndim == 2:
        im = np.repeat(im[:, :, None], 3, axis=2)
    elif im.shape[2] >= 3:
        im = im[:, :, [2, 1, 0] + list(range(3, im.shape[2]))]   # BGR(A) -> RGB(A)
    if im.shape[2] == 3:
        one = {np.dtype(np.uint8): 255, np.dtype(np.uint16): 65535}.get(im.dtype, 1.0)
        im = np.concatenate([im, np.full(im.shape[:2] + (1,), one, dtype=im.dtype)], axis=2)
    if im.dtype == np.uint16:
        im = (im.astype(np.float32) / np.float32(65535.0)).astype(np.float32)
    elif im.dtype != np.uint8:
        im = im.astype(np.float32)
    return Texture(im, path=path)


class Scene:
    """``Scene::new`` (scene.rs:77-122): textures 0..8 are the default constants, texture 9 is the environment
    map, material 0 is the default material."""

    def __init__(self, envmap):
        """envmap: a ``Texture`` (or H x W x 4 array) for texture slot 9 — the reference hard-codes
        ``res/knob/envmap.exr`` (scene.rs:93, sky.wgsl:10) — or a path to decode."""
        self.vertices = np.zeros(0, dtype=T.VERTEX_DT)
        self.indices = np.zeros(0, dtype=np.uint32)
        self.meshes = np.zeros(0, dtype=T.MESH_DT)
        env = envmap if isinstance(envmap, Texture) else (decode_image(envmap) if isinstance(envmap, str) else Texture(envmap))
        self.textures = [Texture.from_color([1, 1, 1, 1]), Texture.from_scalar(0.99), Texture.from_color([.5, .5, .5, 1]),
                         Texture.from_scalar(0.0), Texture.from_color([0, 0, 0, 1]), Texture.from_scalar(0.0),
                         Texture.from_color([.5, .5, 1, 1]), Texture.from_scalar(0.0), Texture.from_scalar(1.0), env]
        m = np.zeros(1, dtype=T.MATERIAL_DT)
        m["albedo_texture_idx"], m["roughness_texture_idx"], m["specular_texture_idx"], m["metallic_texture_idx"] = 0, 1, 2, 3
        m["emissive_texture_idx"], m["sheen_texture_idx"], m["normal_texture_idx"] = 4, 5, 6
        m["transmission_texture_idx"], m["ior_texture_idx"] = 7, 8
        self.materials = m
        self._keep = None

    # -- ingest ------------------------------------------------------------------------------------
    def load_obj(self, path=None, obj_text=None, mtl_text=None, obj_dir="", fix_normals=False, decode=decode_image, decode_threads=0, cuda_device=0):
        """``Scene::load_obj`` (scene.rs:124-257) through the C++ loader; appends to this scene.  Only valid on
        a scene that has not been modified since construction or a previous load_obj (the loader replays the
        same texture/material numbering).  decode: a callable path -> Texture (default: OpenCV), or "library" =
        ``ptc_scene_decode_textures`` (PNG / PFM on the library's host threads, JPEG by nvJPEG on ``cuda_device``)."""
        L = lib()
        h = C.c_void_p()
        check(L.ptc_scene_new(b"<envmap>", C.byref(h)))
        try:
            check(L.ptc_scene_set_fix_normals(h, 1 if fix_normals else 0))
            # replay what this scene already holds so indices continue correctly
            if len(self.textures) != 10 or len(self.materials) != 1 or len(self.indices):
                raise NotImplementedError("load_obj on a non-fresh Scene: load all OBJ files through one loader handle")
            if path is not None:
                check(L.ptc_scene_load_obj(h, os.fsencode(path)))
            else:
                check(L.ptc_scene_load_obj_text(h, obj_text.encode(), None if mtl_text is None else mtl_text.encode(), obj_dir.encode()))
            n = L.ptc_scene_num_textures(h)
            texs = list(self.textures)
            files = []
            for i in range(10, n):
                info = T.SceneTextureInfo()
                check(L.ptc_scene_texture_info_at(h, i, C.byref(info)))
                if info.source == T.TEX_FILE:
                    files.append((len(texs), info.path.decode()))
                    texs.append(None)
                else:
                    texs.append(Texture(np.array(list(info.color), dtype=np.float32).reshape(1, 1, 4)))
            # decode the files on a pool of threads, as the reference does (32 loader threads per chunk, scene.rs:239-254);
            # OpenCV releases the GIL while decoding.  Results are stored by index, so the order of completion does not matter.
            if decode == "library":
                for i, t in enumerate(texs[:10]):                 # the environment map (slot 9) is this scene's, not a file of the loader's
                    w, hh = t.size
                    check(L.ptc_scene_set_texels(h, i, w, hh, t.format, t.texels.ctypes.data_as(C.c_void_p)))
                check(L.ptc_scene_decode_textures(h, cuda_device, decode_threads))
            elif len(files) > 1 and decode_threads != 1:
                from concurrent.futures import ThreadPoolExecutor
                with ThreadPoolExecutor(max_workers=min(decode_threads or 32, len(files))) as pool:
                    for (slot, _), tex in zip(files, pool.map(decode, [p for _, p in files])):
                        texs[slot] = tex
            else:
                for slot, path in files:
                    texs[slot] = decode(path)
            # give the loader texels so get_view succeeds
            for i, t in enumerate(texs):
                if t is None:
                    continue                      # decoded by the library
                w, hh = t.size
                check(L.ptc_scene_set_texels(h, i, w, hh, t.format, t.texels.ctypes.data_as(C.c_void_p)))
            v = T.SceneView()
            check(L.ptc_scene_get_view(h, C.byref(v)))
            descs = C.cast(v.textures, C.POINTER(T.TextureDesc))
            for slot, fpath in files:
                if texs[slot] is None:
                    d = descs[slot]
                    px = d.width * d.height
                    raw = C.string_at(d.texels, px * (4 if d.format == T.TEXEL_RGBA8 else 16))
                    arr = np.frombuffer(raw, dtype=np.uint8 if d.format == T.TEXEL_RGBA8 else np.float32).reshape(d.height, d.width, 4).copy()
                    texs[slot] = Texture(arr, path=fpath)
            self.vertices = np.frombuffer(C.string_at(v.vertices, v.num_vertices * 48), dtype=T.VERTEX_DT).copy()
            self.indices = np.frombuffer(C.string_at(v.indices, v.num_indices * 4), dtype=np.uint32).copy()
            self.meshes = np.frombuffer(C.string_at(v.meshes, v.num_meshes * 20), dtype=T.MESH_DT).copy()
            self.materials = np.frombuffer(C.string_at(v.materials, v.num_materials * 52), dtype=T.MATERIAL_DT).copy()
            self.textures = texs
        finally:
            L.ptc_scene_free(h)
        return self

    # -- direct construction (procedural stand-ins) ---------------------------------------------------
    def add_texture(self, tex):
        self.textures.append(tex if isinstance(tex, Texture) else Texture(tex))
        return len(self.textures) - 1

    def add_material(self, **slots):
        m = self.materials[:1].copy()
        for k, v in slots.items():
            m[k] = v
        self.materials = np.concatenate([self.materials, m])
        return len(self.materials) - 1

    def add_mesh(self, positions, normals, uvs, indices, material_index):
        """Append one mesh the way scene.rs:169-234 does (global index rebasing, one ``Mesh`` record)."""
        positions = np.asarray(positions, dtype=np.float32).reshape(-1, 3)
        nv = len(positions)
        v = np.zeros(nv, dtype=T.VERTEX_DT)
        v["position"] = positions
        if normals is not None:
            v["normal"] = np.asarray(normals, dtype=np.float32).reshape(-1, 3)
        if uvs is not None:
            v["uv"] = np.asarray(uvs, dtype=np.float32).reshape(-1, 2)
        idx = np.asarray(indices, dtype=np.uint32).reshape(-1)
        mesh = np.zeros(1, dtype=T.MESH_DT)
        mesh["vertex_offset"], mesh["vertex_count"] = len(self.vertices), nv
        mesh["index_offset"], mesh["index_count"], mesh["material_index"] = len(self.indices), len(idx), material_index
        self.indices = np.concatenate([self.indices, idx + np.uint32(len(self.vertices))])
        self.vertices = np.concatenate([self.vertices, v])
        self.meshes = np.concatenate([self.meshes, mesh])
        return len(self.meshes) - 1

    def add_meshes(self, parts):
        """Append many meshes at once: parts = iterable of (positions, normals, uvs, indices, material_index).
        Same records as repeated add_mesh, one concatenation (stand-ins with thousands of meshes)."""
        V, I, M = [self.vertices], [self.indices], [self.meshes]
        nv0, ni0 = len(self.vertices), len(self.indices)
        for positions, normals, uvs, indices, material_index in parts:
            positions = np.asarray(positions, dtype=np.float32).reshape(-1, 3)
            v = np.zeros(len(positions), dtype=T.VERTEX_DT)
            v["position"] = positions
            if normals is not None:
                v["normal"] = np.asarray(normals, dtype=np.float32).reshape(-1, 3)
            if uvs is not None:
                v["uv"] = np.asarray(uvs, dtype=np.float32).reshape(-1, 2)
            idx = np.asarray(indices, dtype=np.uint32).reshape(-1)
            mesh = np.zeros(1, dtype=T.MESH_DT)
            mesh["vertex_offset"], mesh["vertex_count"] = nv0, len(v)
            mesh["index_offset"], mesh["index_count"], mesh["material_index"] = ni0, len(idx), material_index
            V.append(v); I.append(idx + np.uint32(nv0)); M.append(mesh)
            nv0 += len(v); ni0 += len(idx)
        self.vertices, self.indices, self.meshes = np.concatenate(V), np.concatenate(I), np.concatenate(M)

    @property
    def num_triangles(self):
        return len(self.indices) // 3

    # -- C view ----------------------------------------------------------------------------------------
    def view(self):
        """Borrowed ``ptc_scene_view`` over the numpy arrays (kept alive by this object)."""
        self.vertices = np.ascontiguousarray(self.vertices)
        self.indices = np.ascontiguousarray(self.indices, dtype=np.uint32)
        self.meshes = np.ascontiguousarray(self.meshes)
        self.materials = np.ascontiguousarray(self.materials)
        descs = (T.TextureDesc * len(self.textures))()
        for i, t in enumerate(self.textures):
            w, h = t.size
            descs[i].width, descs[i].height, descs[i].format = w, h, t.format
            descs[i].texels = t.texels.ctypes.data
        v = T.SceneView()
        v.vertices, v.num_vertices = self.vertices.ctypes.data, len(self.vertices)
        v.indices, v.num_indices = self.indices.ctypes.data, len(self.indices)
        v.meshes, v.num_meshes = self.meshes.ctypes.data, len(self.meshes)
        v.materials, v.num_materials = self.materials.ctypes.data, len(self.materials)
        v.textures, v.num_textures = descs, len(self.textures)
        self._keep = descs
        return v
