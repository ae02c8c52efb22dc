"""Host-side camera path of `GaussianRasterizerWrapper`.

The reference builds every frame's settings (DGR/__init__.py:349-402) as: numpy/scipy w2c on the
host -> H2D copy -> `w2c @ P^T` as a device matmul -> `w2c.inverse()[3,:3]` as a device LU
(a cuSOLVER call for a 4x4 matrix) -> a fresh `bg` tensor (another H2D copy).  At GaussianCity's own
scale (P <= 16 384 points, ~0.25 ms of kernels per frame) those launches, copies and the solver's
synchronisation rival the rasterizer itself.

`PoseSettingsCache` computes the same quantities on the host and ships them in ONE packed upload:

  * view_matrix : the SAME numpy expression as the reference (bit-identical float32 values);
  * proj_matrix : `view_matrix @ P^T` accumulated in float64, rounded once to float32 (within 1 ulp
                  of any float32 evaluation order of the reference's matmul);
  * campos      : the camera position itself (what `inverse(view)[3,:3]` equals for a rigid w2c,
                  without the LU's rounding);
  * bg          : zeros, in the same buffer.

Poses are cached (LRU): an orbit that revisits a pose, or a trainer that renders several batches
from one view, pays nothing.  Opt-in through `GaussianRasterizerWrapper(..., fast_camera=True)`:
proj_matrix may differ from the reference's device matmul in the last bit, so the default wrapper
path stays the reference's own sequence of operations.
"""
from collections import OrderedDict

import numpy as np
import torch

# packed layout (float32 words): 16-word sections keep every tensor 16 B-aligned
_VIEW, _PROJ, _CAMPOS, _BG, _WORDS = 0, 16, 32, 36, 40


def quat_xyzw_to_matrix(q):
    """Rotation matrix of a (qx,qy,qz,qw) quaternion in float64.  Uses scipy when present (the
    reference does, DGR/__init__.py:355) so view matrices round to identical float32 values."""
    q = np.asarray(q, dtype=np.float64)
    try:
        import scipy.spatial.transform
        return scipy.spatial.transform.Rotation.from_quat(q).as_matrix()
    except ImportError:  # same formula, normalised quaternion
        x, y, z, w = q / np.linalg.norm(q)
        return np.array([
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _to_numpy(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


def w2c_host(cam_position, cam_quaternion):
    """float32 world-to-camera matrix exactly as DGR/__init__.py:349-368 forms it on the host:
    R = from_quat(q)[:, [1,2,0]] ([F|R|U] -> [R|U|F]);  Rt = [[R^T, -R^T p], [0, 1]]."""
    pos = _to_numpy(cam_position)
    R = quat_xyzw_to_matrix(_to_numpy(cam_quaternion))[:, [1, 2, 0]]
    Rt = np.zeros((4, 4), dtype=np.float32)
    Rt[:3, :3] = R.transpose()
    Rt[:3, [3]] = -R.transpose() @ pos[:, None]
    Rt[3, 3] = 1.0
    return Rt


class PoseSettingsCache:
    """pose -> (view_matrix, proj_matrix, campos, bg) device tensors, one upload per new pose."""

    def __init__(self, P_host, device, capacity=64):
        self.PT64 = np.asarray(P_host, dtype=np.float32).astype(np.float64).T.copy()
        self.device = device
        self.capacity = int(capacity)
        self._lru = OrderedDict()
        self.uploads = 0          # number of host->device copies issued (one per cache miss)

    @staticmethod
    def _key(pos, quat):
        return (pos.dtype.str, pos.tobytes(), quat.dtype.str, quat.tobytes())

    def pack(self, cam_position, cam_quaternion):
        """The 40-word host buffer for one pose (exposed for tests)."""
        w2c_t = np.ascontiguousarray(w2c_host(cam_position, cam_quaternion).T)   # = view_matrix
        buf = np.zeros(_WORDS, dtype=np.float32)
        buf[_VIEW:_VIEW + 16] = w2c_t.reshape(-1)
        buf[_PROJ:_PROJ + 16] = (w2c_t.astype(np.float64) @ self.PT64).astype(np.float32).reshape(-1)
        buf[_CAMPOS:_CAMPOS + 3] = _to_numpy(cam_position).astype(np.float32).reshape(3)
        return buf

    def get(self, cam_position, cam_quaternion):
        pos = np.ascontiguousarray(_to_numpy(cam_position))
        quat = np.ascontiguousarray(_to_numpy(cam_quaternion))
        key = self._key(pos, quat)
        hit = self._lru.get(key)
        if hit is not None:
            self._lru.move_to_end(key)
            return hit
        dev = torch.from_numpy(self.pack(pos, quat)).to(self.device)
        self.uploads += 1
        out = (dev[_VIEW:_VIEW + 16].view(4, 4), dev[_PROJ:_PROJ + 16].view(4, 4),
               dev[_CAMPOS:_CAMPOS + 3], dev[_BG:_BG + 3])
        self._lru[key] = out
        while len(self._lru) > self.capacity:
            self._lru.popitem(last=False)
        return out
