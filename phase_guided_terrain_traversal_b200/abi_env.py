"""Thin object wrapper over the C ABI handle: owns a `pgtt_env*`, exposes its device buffers as
zero-copy array views and forwards the calls. Two array back-ends:

* "torch"  - CUDA tensors (the product path; buffers are wrapped through `__cuda_array_interface__`);
* "numpy"  - host arrays; ONLY meaningful with the host-emulated test library
             (tests/simt_emu/libpgtt_emu.so), where "device" memory is host memory.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nat


class _CudaView:
    """Minimal `__cuda_array_interface__` provider for a raw device pointer."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2, "strides": None}


_TYPESTR = {"f": "<f4", "i": "<i4", "u": "<u4"}


class AbiEnv:
    def __init__(self, model, cfg, num_envs: int, device: int = 0, backend: str = "torch", lib=None, rng_partitionable: bool = True,
                 variant: int = 0):
        self.lib = lib if lib is not None else nat.load_library()
        self.backend = backend
        self.N = int(num_envs)
        self.device = device
        self.model, self.cfg = model, cfg
        self._md = nat.model_desc(model)
        self._td = nat.task_desc(cfg, model, rng_partitionable, variant)
        self.variant = int(variant)
        h = C.c_void_p()
        nat.check(self.lib, self.lib.pgtt_create(C.byref(self._md), C.byref(self._td), device, self.N, C.byref(h)))
        self.h = h
        self._keep = []
        if backend == "torch":
            import torch
            self.torch = torch
            self.torch_device = torch.device("cuda", device)
        b = nat.Buffers()
        nat.check(self.lib, self.lib.pgtt_get_buffers(self.h, C.byref(b)))
        self.buf = {}
        no, npv = C.c_int(), C.c_int()
        nat.check(self.lib, self.lib.pgtt_obs_dims(self.h, C.byref(no), C.byref(npv)))
        self.nobs, self.npriv = no.value, npv.value
        obs_dim = {"obs_state": self.nobs, "obs_privileged": self.npriv, "first_obs_state": self.nobs, "first_obs_privileged": self.npriv}
        for name, kind, dim in nat.BUFFER_FIELDS:
            dim = obs_dim.get(name, dim)
            ptr = C.cast(getattr(b, name), C.c_void_p).value
            self.buf[name] = self._wrap(ptr, kind, (self.N, dim))
        self.n_terrains = 0

    # -- array plumbing ---------------------------------------------------------------
    def _wrap(self, ptr, kind, shape):
        if self.backend == "numpy":
            ct = {"f": C.c_float, "i": C.c_int32, "u": C.c_uint32}[kind]
            n = int(np.prod(shape))
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).reshape(shape)
        t = self.torch.as_tensor(_CudaView(ptr, shape, _TYPESTR[kind]), device=self.torch_device)
        return t

    def _dev(self, arr, dtype):
        """Returns (object to keep alive, raw pointer) of `arr` as a contiguous device array."""
        if self.backend == "numpy":
            a = np.ascontiguousarray(arr, dtype=dtype)
            return a, a.ctypes.data
        torch = self.torch
        tdt = {np.float32: torch.float32, np.uint32: torch.uint32, np.int32: torch.int32}[dtype]
        if isinstance(arr, torch.Tensor):
            t = arr.to(device=self.torch_device, dtype=tdt).contiguous()
        else:
            t = torch.from_numpy(np.ascontiguousarray(arr, dtype=dtype)).to(self.torch_device)
        return t, t.data_ptr()

    def _stream(self):
        if self.backend == "numpy":
            return None
        return C.c_void_p(self.torch.cuda.current_stream(self.torch_device).cuda_stream)

    def _empty(self, shape):
        if self.backend == "numpy":
            a = np.zeros(shape, dtype=np.float32)
            return a, a.ctypes.data
        t = self.torch.zeros(shape, dtype=self.torch.float32, device=self.torch_device)
        return t, t.data_ptr()

    # -- ABI calls ----------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.pgtt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        nat.check(self.lib, self.lib.pgtt_sync(self.h, self._stream()))

    def set_terrain(self, table):
        from .terrain import validate_terrain
        t = np.ascontiguousarray(table, dtype=np.float32)
        validate_terrain(t)      # shape, finiteness, yaw-only boxes, positive half-sizes: unsupported tables fail loudly
        nat.check(self.lib, self.lib.pgtt_set_terrain_table(self.h, t.ctypes.data, t.shape[0]))
        self.n_terrains = t.shape[0]

    def randomize(self, keys, dynamics: bool = True):
        k, p = self._dev(keys, np.uint32)
        assert tuple(k.shape) == (self.N, 2)
        nat.check(self.lib, self.lib.pgtt_randomize(self.h, p, int(dynamics), self._stream()))
        self._keep = [k]

    def reset(self, keys):
        k, p = self._dev(keys, np.uint32)
        assert tuple(k.shape) == (self.N, 2)
        nat.check(self.lib, self.lib.pgtt_reset(self.h, p, self._stream()))
        self._keep = [k]

    def step(self, action, wrapped: bool = True):
        a, p = self._dev(action, np.float32)
        assert tuple(a.shape) == (self.N, 12)
        nat.check(self.lib, self.lib.pgtt_step(self.h, p, int(wrapped), self._stream()))
        self._keep = [a]

    def step_ptr(self, action_ptr: int, wrapped: bool = True, stream=None):
        """Hot-loop variant: caller guarantees a contiguous float32 [N,12] device buffer."""
        nat.check(self.lib, self.lib.pgtt_step(self.h, action_ptr, int(wrapped), stream if stream is not None else self._stream()))

    def forward(self):
        nat.check(self.lib, self.lib.pgtt_forward(self.h, self._stream()))

    def heightscan(self, center, yaw):
        c, pc = self._dev(center, np.float32)
        y, py = self._dev(yaw, np.float32)
        out, po = self._empty((self.N, 117, 3))
        nat.check(self.lib, self.lib.pgtt_heightscan(self.h, pc, py, po, self._stream()))
        self._keep = [c, y]
        return out.reshape(self.N, 13, 9, 3)

    def debug_forward(self):
        out, po = self._empty((self.N, nat.DEBUG_FLOATS))
        nat.check(self.lib, self.lib.pgtt_debug_forward(self.h, po, self._stream()))
        self.sync()
        return out if self.backend == "numpy" else out.cpu().numpy()

    def launch_count(self) -> int:
        return int(self.lib.pgtt_launch_count(self.h))

    def step_kernel(self) -> str:
        if int(self.lib.pgtt_step_kernel_generation(self.h)) == 1:
            return "pgtt_quad_kernel<OP_STEP>"
        return "pgtt_env_kernel<OP_STEP_TASK>" if self.step_launches() == 1 else "pgtt_env_kernel<OP_STEP>"

    def step_launches(self) -> int:
        return int(self.lib.pgtt_step_launches(self.h))

    # -- host-side convenience (tests) ----------------------------------------------------
    def get(self, name) -> np.ndarray:
        self.sync()
        b = self.buf[name]
        return np.array(b) if self.backend == "numpy" else b.cpu().numpy()

    def set(self, name, value):
        b = self.buf[name]
        if self.backend == "numpy":
            b[...] = np.asarray(value).reshape(b.shape)
        else:
            self.sync()
            v = np.asarray(value).reshape(tuple(b.shape))
            b.copy_(self.torch.from_numpy(np.ascontiguousarray(v)).to(b.dtype))
