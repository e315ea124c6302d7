"""Asset / PLY I/O in the formats of volprim/io.py.

* PLY: the 3DGS-compatible ellipsoid file the reference writes with `ellipsoid_dict_to_ply`
  (io.py:362-411) and Mitsuba's `ellipsoidsmesh` loader reads back.  The reader here is the inverse of that
  writer; `plyfile` is not a dependency -- the little PLY codec below handles the
  binary_little_endian / ascii scalar-property files involved.
* Python assets: a directory with `__init__.py` defining OBJECTS / SENSORS / EMITTERS dictionaries plus
  `data/*.ply`, `data/*.npy` (io.py:43-70, 87-272).  Assets written by the reference import `mitsuba` and
  `drjit`; when those are absent a tiny stand-in providing `ScalarTransform4f` is injected while the asset
  module executes.  Mesh / bitmap / envmap content is outside the hot path and rejected.
"""
from __future__ import annotations

import os
import shutil
import sys
import types
from os.path import basename, exists, join, splitext

import numpy as np

from .transforms import Transform4f

_PLY_TYPES = {
    'char': 'i1', 'int8': 'i1', 'uchar': 'u1', 'uint8': 'u1', 'short': 'i2', 'int16': 'i2', 'ushort': 'u2',
    'uint16': 'u2', 'int': 'i4', 'int32': 'i4', 'uint': 'u4', 'uint32': 'u4', 'float': 'f4', 'float32': 'f4',
    'double': 'f8', 'float64': 'f8',
}


def read_ply_vertices(filename) -> np.ndarray:
    """Structured array of the `vertex` element (scalar properties only)."""
    with open(filename, 'rb') as f:
        if f.readline().strip() != b'ply':
            raise Exception(f"{filename}: not a PLY file")
        fmt, elements, cur = None, [], None
        while True:
            line = f.readline()
            if not line:
                raise Exception(f"{filename}: truncated PLY header")
            tok = line.decode('ascii', 'replace').split()
            if not tok or tok[0] == 'comment' or tok[0] == 'obj_info':
                continue
            if tok[0] == 'format':
                fmt = tok[1]
            elif tok[0] == 'element':
                cur = {'name': tok[1], 'count': int(tok[2]), 'props': []}
                elements.append(cur)
            elif tok[0] == 'property':
                if tok[1] == 'list':
                    cur['props'].append(('list', tok[2], tok[3], tok[4]))
                else:
                    cur['props'].append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == 'end_header':
                break
        out = None
        for el in elements:
            if any(p[0] == 'list' for p in el['props']):
                if el['name'] == 'vertex':
                    raise Exception("list properties on the vertex element are not supported")
                break  # faces etc. after the vertices: not needed
            if fmt == 'ascii':
                rows = [f.readline().split() for _ in range(el['count'])]
                dt = np.dtype([(n, t) for n, t in el['props']])
                arr = np.empty(el['count'], dt)
                for i, (n, t) in enumerate(el['props']):
                    arr[n] = np.array([r[i] for r in rows], dtype=t)
            else:
                end = '<' if fmt == 'binary_little_endian' else '>'
                dt = np.dtype([(n, end + t) for n, t in el['props']])
                arr = np.frombuffer(f.read(dt.itemsize * el['count']), dtype=dt, count=el['count'])
            if el['name'] == 'vertex':
                out = arr
                break
        if out is None:
            raise Exception(f"{filename}: no vertex element")
        return out


def write_ply_vertices(filename, names, columns: np.ndarray):
    """binary_little_endian PLY with float32 scalar properties `names` (the layout plyfile produces for the
    reference's writer, io.py:406-411)."""
    columns = np.ascontiguousarray(columns, dtype='<f4')
    assert columns.ndim == 2 and columns.shape[1] == len(names)
    with open(filename, 'wb') as f:
        f.write(b'ply\nformat binary_little_endian 1.0\n')
        f.write(f'element vertex {columns.shape[0]}\n'.encode())
        for n in names:
            f.write(f'property float {n}\n'.encode())
        f.write(b'end_header\n')
        f.write(columns.tobytes())


def _sh_rest_permutation(sh_n: int):
    """Column order the writer applies to f_rest (io.py:383-386): channel-major planes."""
    col_mapping = sum([[(j * 3 + 0 - 3, j - 1 + 3 - 3), (j * 3 + 1 - 3, j - 1 + sh_n + 2 - 3),
                        (j * 3 + 2 - 3, j - 1 + 2 * sh_n + 1 - 3)] for j in range(1, sh_n)], [])
    return [a for a, b in sorted(col_mapping, key=lambda x: x[1])]


def load_ellipsoids_ply(filename, normalize_quaternions: bool = False) -> dict:
    """Inverse of `ellipsoid_dict_to_ply`: scales = exp(scale_k) (io.py:372), quaternion (rot_0..3) =
    (r,i,j,k) -> (i,j,k,r) (io.py:373), opacities = sigmoid(opacity) (io.py:388-389), sh_coeffs from
    f_dc / f_rest with the writer's permutation undone (io.py:381-386), generic attributes name_k -> name[:,k]
    (io.py:401-403); normals are ignored.  Quaternions are NOT normalised (reference quirk Q6) unless asked."""
    v = read_ply_vertices(filename)
    names = v.dtype.names
    col = lambda n: np.asarray(v[n], np.float32)
    d = {
        'centers': np.stack([col('x'), col('y'), col('z')], 1),
        'scales': np.exp(np.stack([col('scale_0'), col('scale_1'), col('scale_2')], 1)),
    }
    q = np.stack([col('rot_0'), col('rot_1'), col('rot_2'), col('rot_3')], 1)[:, [1, 2, 3, 0]]
    if normalize_quaternions:
        q = q / np.linalg.norm(q, axis=1, keepdims=True)
    d['quaternions'] = q.astype(np.float32)
    skip = {'x', 'y', 'z', 'nx', 'ny', 'nz', 'scale_0', 'scale_1', 'scale_2', 'rot_0', 'rot_1', 'rot_2', 'rot_3'}
    if 'f_dc_0' in names and 'opacity' in names:
        f_dc = np.stack([col(f'f_dc_{i}') for i in range(3)], 1)
        n_rest = sum(1 for n in names if n.startswith('f_rest_'))
        if n_rest:
            f_rest_file = np.stack([col(f'f_rest_{i}') for i in range(n_rest)], 1)
            perm = _sh_rest_permutation(n_rest // 3 + 1)
            f_rest = np.empty_like(f_rest_file)
            f_rest[:, perm] = f_rest_file
            d['sh_coeffs'] = np.concatenate([f_dc, f_rest], 1)
        else:
            d['sh_coeffs'] = f_dc
        d['opacities'] = (1.0 / (1.0 + np.exp(-col('opacity').astype(np.float64)))).astype(np.float32)[:, None]
        skip |= {'opacity'} | {n for n in names if n.startswith('f_dc_') or n.startswith('f_rest_')}
    groups = {}
    for n in names:
        if n in skip:
            continue
        base, _, k = n.rpartition('_')
        if base and k.isdigit():
            groups.setdefault(base, []).append((int(k), n))
    for base, cols in groups.items():
        d[base] = np.stack([col(n) for _, n in sorted(cols)], 1)
    return d


def ellipsoid_dict_to_ply(d, extras_keys, filename):
    '''
    Export a dictionary representing an ellipsoid shape into a PLY file (same property order, encodings
    and clamps as the reference writer, io.py:362-411).
    '''
    arr = lambda x: np.asarray(x.detach().cpu() if hasattr(x, 'detach') else x, dtype=np.float64)
    is_3dg = 'sh_coeffs' in extras_keys and 'opacities' in extras_keys
    extras = {k: arr(d[k]).reshape(arr(d['centers']).shape[0], -1).shape[1] for k in extras_keys}

    centers = arr(d['centers'])
    n = centers.shape[0]
    scales = np.log(np.maximum(arr(d['scales']), 1e-6))
    quaternions = arr(d['quaternions'])[:, [3, 0, 1, 2]]  # Reorder (i, j, k, r -> r, i, j, k)
    normals = np.zeros_like(centers)

    if is_3dg:
        sh_coeffs = arr(d['sh_coeffs']).reshape(n, -1)
        f_dc, f_rest = sh_coeffs[:, :3], sh_coeffs[:, 3:]
        if f_rest.shape[1] > 0:
            f_rest = f_rest[:, _sh_rest_permutation((f_rest.shape[1] // 3) + 1)]
        opacities = np.clip(arr(d['opacities']).reshape(n, -1), 1e-8, 1.0 - 1e-8)
        opacities = np.log(opacities) - np.log(1.0 - opacities)
        extras_data = [f_dc, f_rest, opacities]
    else:
        extras_data = [arr(d[k]).reshape(n, -1) for k in extras]

    attributes = ['x', 'y', 'z', 'nx', 'ny', 'nz']
    if is_3dg:
        attributes += ['f_dc_0', 'f_dc_1', 'f_dc_2']
        attributes += [f'f_rest_{i}' for i in range(extras['sh_coeffs'] - 3)]
        attributes += ['opacity']
    else:
        for k, dim in extras.items():
            attributes += [f'{k}_{i}' for i in range(dim)]
    attributes += ['scale_0', 'scale_1', 'scale_2', 'rot_0', 'rot_1', 'rot_2', 'rot_3']
    table = np.concatenate((centers, normals, *extras_data, scales, quaternions), axis=1)
    d['filename'] = filename
    write_ply_vertices(filename, attributes, table)


# ---------------------------------------------------------------------------------------------------
# Python assets
# ---------------------------------------------------------------------------------------------------
class _AssetShim:
    """Makes `import mitsuba as mi`, `import drjit as dr`, `from mitsuba.scalar_rgb import ScalarTransform4f as T`
    resolvable while an asset's __init__.py executes on a machine without Mitsuba."""

    def __enter__(self):
        self.added = []
        try:
            import mitsuba  # noqa: F401
            import drjit  # noqa: F401
            return self
        except Exception:
            pass
        mi = types.ModuleType('mitsuba')
        mi.ScalarTransform4f = Transform4f
        mi.ScalarTransform3f = lambda m: np.asarray(m, np.float64)
        srgb = types.ModuleType('mitsuba.scalar_rgb')
        srgb.ScalarTransform4f = Transform4f
        mi.scalar_rgb = srgb
        dr = types.ModuleType('drjit')
        for name, mod in (('mitsuba', mi), ('mitsuba.scalar_rgb', srgb), ('drjit', dr)):
            if name not in sys.modules:
                sys.modules[name] = mod
                self.added.append(name)
        return self

    def __exit__(self, *exc):
        for name in self.added:
            sys.modules.pop(name, None)


def asset_to_dict(asset, objects=True, emitters=True, sensors=True, integrator=True) -> dict:
    '''
    Assemble a scene Python dictionary for a given asset (reference io.py:43-70).

    Parameter:
        asset: (str or module) path to asset or Python module
    '''
    if isinstance(asset, str):
        from importlib.machinery import SourceFileLoader
        init_path = join(asset, '__init__.py')
        if not exists(init_path):
            raise Exception(f'Invalid asset path: {init_path}')
        with _AssetShim():
            module = types.ModuleType('asset')
            module.__file__ = init_path
            loader = SourceFileLoader('asset', init_path)
            loader.exec_module(module)
        asset = module

    d = {'type': 'scene'}
    if objects:
        d.update(getattr(asset, 'OBJECTS', {}))
    if emitters:
        d.update(getattr(asset, 'EMITTERS', {}))
    if sensors:
        d.update(getattr(asset, 'SENSORS', {}))
    if integrator and hasattr(asset, 'INTEGRATOR'):
        d['integrator'] = asset.INTEGRATOR
    root = os.path.dirname(getattr(asset, '__file__', '') or '')
    _resolve_filenames(d, root)
    return d


def _resolve_filenames(d, root):
    """The reference resolves relative filenames through the 'resources' entry (a Mitsuba file resolver)."""
    for k, v in list(d.items()):
        if isinstance(v, dict):
            if v.get('type') == 'resources':
                continue
            _resolve_filenames(v, root)
        elif k == 'filename' and isinstance(v, str) and not os.path.isabs(v) and root:
            d[k] = join(root, v)


def scale_films(d: dict, scale: float = 1.0) -> dict:
    '''
    Scale the films resolution in the given scene dictionary (reference io.py:72-85)
    '''
    def walk(d):
        for k, v in d.items():
            if k == 'film':
                v['width'] = int(scale * v['width'])
                v['height'] = int(scale * v['height'])
            elif isinstance(v, dict):
                walk(v)
    walk(d)
    return d


_ASSET_HEADER = ('import os', 'from os.path import join, dirname', 'import numpy as np', 'import drjit as dr',
                 'import mitsuba as mi', 'from mitsuba.scalar_rgb import ScalarTransform4f as T')
_ASSET_SENSORS = ('perspective', 'orthographic', 'thinlens')
_ASSET_EMITTERS = ('envmap', 'constant', 'point', 'distant', 'spot', 'directional')
_ASSET_SUBFOLDER = {'.json': 'data', '.obj': 'meshes', '.jpg': 'textures', '.png': 'textures', '.exr': 'textures'}
_ASSET_FOREIGN = ('meshholder', 'obj', 'ply', 'bitmap', 'envmap')      # scene content outside the volprim hot path


class _AssetWriter:
    """Emits the text of an asset's `__init__.py` (the file format of reference io.py:87-272: three dictionaries
    OBJECTS / SENSORS / EMITTERS written as Python literals, arrays and primitive clouds moved to files next to it).
    One `emit_*` method per kind of value; `emit` dispatches on the value."""

    def __init__(self, folder: str):
        self.folder = folder

    # -- files that travel with the asset ----------------------------------------------------------------
    def _subdir(self, name: str) -> str:
        os.makedirs(join(self.folder, name), exist_ok=True)
        return name

    def store_ellipsoids(self, node: dict, path: str) -> dict:
        """Tensor-defined ellipsoids shape -> data/<path>.ply; returns the node with a `filename` instead of arrays."""
        is_array = lambda v: isinstance(v, np.ndarray) or hasattr(v, 'detach')
        geometry = ('centers', 'scales', 'quaternions')
        attributes = [k for k, v in node.items() if is_array(v) and k not in geometry]
        target = join(self.folder, self._subdir('data'), f'{path}.ply')
        ellipsoid_dict_to_ply(node, attributes, target)
        slim = {k: v for k, v in node.items() if k not in geometry and k not in attributes}
        slim['filename'] = target
        return slim

    def store_array(self, value, path: str, key: str) -> str:
        rel = f"{self._subdir('data')}/{path}.{key}.npy"
        np.save(join(self.folder, rel), np.asarray(value.detach().cpu() if hasattr(value, 'detach') else value))
        return rel

    def store_file(self, src: str, ellipsoids: bool) -> str:
        stem, ext = splitext(basename(src))
        sub = 'data' if (ext == '.ply' and ellipsoids) else ('meshes' if ext == '.ply' else _ASSET_SUBFOLDER[ext])
        rel = join(self._subdir(sub), stem + ext)
        if not exists(join(self.folder, rel)):
            shutil.copy(src, join(self.folder, rel))
        return rel

    # -- literals --------------------------------------------------------------------------------------
    @staticmethod
    def emit_transform(value, pad: str, as_look_at: bool) -> str:
        T = Transform4f(value)
        if as_look_at:      # sensors are written as look_at(origin, target, up), like the reference does
            origin, target = T @ np.zeros(3), T @ np.array([0.0, 0.0, 1.0])
            up = T.transform_vector([0.0, 1.0, 0.0])
            rows = [f"origin={origin.tolist()}", f"target={target.tolist()}", f"up={up.tolist()}"]
            return "T().look_at(\n" + ''.join(f"{pad}         {r},\n" for r in rows) + f"{pad}     )"
        m = T.matrix.tolist()
        return f"T([{m[0]}, {m[1]}, {m[2]}, {m[3]}])"

    def emit(self, key: str, value, node_type, ellipsoids: bool, path: str, pad: str, depth: int) -> str:
        if isinstance(value, dict):
            return self.emit_dict(value, f'{path}.{key}', depth + 1)
        if isinstance(value, str):
            if key == 'filename':
                return "r'" + self.store_file(value, ellipsoids) + "'"
            return "'" + (value.replace('.', '_') if key == 'id' else value) + "'"
        if key == 'to_world' or isinstance(value, Transform4f):
            return self.emit_transform(value, pad, as_look_at=key == 'to_world' and node_type in _ASSET_SENSORS)
        if isinstance(value, np.ndarray) or hasattr(value, 'detach'):
            return f"np.load(join(dirname(__file__), '{self.store_array(value, path, key)}'))"
        return str(value)

    def emit_dict(self, node: dict, path: str, depth: int = 0, with_resources: bool = False) -> str:
        pad = ' ' * (4 * depth)
        node_type = node.get('type')
        if node_type in _ASSET_FOREIGN:
            raise Exception(f"dict_to_asset: '{node_type}' objects are outside the volprim hot path")
        ellipsoids = isinstance(node_type, str) and 'ellipsoid' in node_type
        if ellipsoids and 'filename' not in node:
            node = self.store_ellipsoids(dict(node), path)
        lines = []
        if with_resources:
            lines.append("'resources': { 'type': 'resources', 'path': dirname(__file__) }")
        for key, value in node.items():
            if isinstance(value, dict) and value.get('type') == 'resources':
                continue
            lines.append(f"'{key.replace('.', '_')}': " + self.emit(key, value, node_type, ellipsoids, path, pad, depth))
        text = '{\n' + ''.join(f"{pad}    {line},\n" for line in lines) + pad + '}'
        return text.replace('\\', '/')


def dict_to_asset(scene_dict: dict, output_folder: str, verbose=False):
    '''
    Generate a Python asset that contains a dictionary that represents a scene (file format of reference
    io.py:87-272): `__init__.py` with OBJECTS / SENSORS / EMITTERS, ellipsoid shapes in data/<path>.ply, arrays in
    data/<path>.<key>.npy.  The header imports are the reference's, so the asset also loads under Mitsuba.
    '''
    assert scene_dict['type'] == 'scene', 'can only process scene dictionary!'
    print(f'Writing asset to {output_folder} ...')
    groups = {'OBJECTS': {}, 'SENSORS': {}, 'EMITTERS': {}}
    for key, value in scene_dict.items():
        if key == 'type':
            continue
        kind = value['type']
        groups['SENSORS' if kind in _ASSET_SENSORS else 'EMITTERS' if kind in _ASSET_EMITTERS else 'OBJECTS'][key] = value
    os.makedirs(output_folder, exist_ok=True)
    writer = _AssetWriter(output_folder)
    blocks = [f'{title} = ' + writer.emit_dict(group, 'root', with_resources=True) for title, group in groups.items()]
    with open(join(output_folder, '__init__.py'), 'w') as f:
        f.write('\n'.join(_ASSET_HEADER) + '\n\n' + '\n\n'.join(blocks) + '\n')


def object_to_dict(root) -> dict:
    '''
    Convert a scene of this package back into the dictionary `load_dict()` accepts (reference io.py:275-360;
    ellipsoid shapes split `data` N x 10 into centers / scales / quaternions and reshape attributes to N x k,
    io.py:322-331).
    '''
    from .scene import EllipsoidsShape, PerspectiveSensor, Scene
    if isinstance(root, EllipsoidsShape):
        n = root.count
        data = root.data.detach().cpu().numpy().reshape(n, 10)
        d = {'type': root.plugin, 'centers': data[:, 0:3].copy(), 'scales': data[:, 3:6].copy(),
             'quaternions': data[:, 6:10].copy(), 'extent': root.extent}
        for k, v in root.attributes.items():
            d[k] = v.detach().cpu().numpy().reshape(n, -1)
        return d
    if isinstance(root, PerspectiveSensor):
        return {'type': 'perspective', 'fov_axis': 'x', 'fov': root.fov, 'to_world': root.to_world,
                'near_clip': root.near_clip, 'far_clip': root.far_clip,
                'principal_point_offset_x': root.cx, 'principal_point_offset_y': root.cy,
                'film': {'type': 'hdrfilm', 'width': root.width, 'height': root.height,
                         'rfilter': {'type': root.rfilter}}}
    if isinstance(root, Scene):
        d = {'type': 'scene'}
        integ = root.integrator
        if integ is not None:
            name = {'VolumetricPrimitiveRadianceFieldIntegrator': 'volprim_rf',
                    'VolumetricPrimitiveTomographyIntegrator': 'volprim_tomography'}[type(integ).__name__]
            di = {'type': name, 'max_depth': -1 if integ.max_depth == 0xFFFFFFFF else integ.max_depth,
                  'kernel_type': integ.kernel.type}
            if name == 'volprim_rf':
                di['srgb_primitives'] = integ.srgb_primitives
            d['integrator'] = di
        for sh in root.shapes():
            d[sh.id] = object_to_dict(sh)
        for i, s in enumerate(root.sensors()):
            if isinstance(s, PerspectiveSensor):
                d[f'sensor_{i:04d}'] = object_to_dict(s)
        if root._env is not None:
            d['environment'] = {'type': 'constant', 'radiance': list(root._env)}
        return d
    raise Exception(f"object_to_dict: unsupported object {type(root)}")
