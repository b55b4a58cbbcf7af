"""Batch handle: thin Python owner of an `agarcl_batch*` (include/agarcl_b200.h).

This is the host-side mirror of the batched hot path.  Device buffers are owned by the C library;
`obs_tensor()/rewards_tensor()/dones_tensor()` expose them to PyTorch without a copy through
`__cuda_array_interface__` (torch is plumbing here: device memory views and streams only).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._abi import OBS_I16, Layout, StateView

_vp = C.c_void_p


class _CudaView:
    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2}
        self._owner = owner


class _ObsListsStruct(C.Structure):
    """struct agarcl_obs_lists"""
    _fields_ = [(n, C.c_int32) for n in ("n_images", "n_chunks", "images_per_chunk", "frames", "channels", "grid", "obs_dtype",
                                         "mask_words", "rec_words", "entries_per_image")] + \
               [(n, C.c_uint32) for n in ("off_rec", "off_entries", "chunk_words")] + [("chunks", _vp), ("slot_of", _vp)]


class ObsLists:
    """numpy view of the lists of the last Batch.step_lists (layout: include/agarcl_b200.h, agarcl_obs_lists)"""

    def __init__(self, batch, L):
        self._b, self.L = batch, L
        words = L.n_chunks * L.chunk_words
        self.words = np.ctypeslib.as_array(C.cast(L.chunks, C.POINTER(C.c_uint32)), shape=(words,))
        self.slot_of = np.ctypeslib.as_array(C.cast(L.slot_of, C.POINTER(C.c_uint32)), shape=(L.n_images,))

    def record(self, image):
        L = self.L
        s = int(self.slot_of[image])
        c, li = divmod(s, L.images_per_chunk)
        base = c * L.chunk_words
        rec = self.words[base + L.off_rec + li * L.rec_words: base + L.off_rec + (li + 1) * L.rec_words]
        return c, rec

    def entries(self, image):
        """(op, element offset, operand) arrays of the image, in application order; None if it overflowed its slot"""
        L = self.L
        c, rec = self.record(image)
        if int(rec[0]) == 0xFFFFFFFF:
            return None
        e0 = c * L.chunk_words + L.off_entries + 2 * int(rec[1])
        e = self.words[e0: e0 + 2 * int(rec[0])].reshape(-1, 2)
        return e[:, 0] >> 29, e[:, 0] & 0x1FFFFFFF, e[:, 1].astype(np.int64)

    def decode(self, image):
        """pure-numpy decoder of one image from the documented layout (what a consumer writes; the C decoder is expand())"""
        L = self.L
        G, CH = L.grid, L.frames * L.channels
        dt = np.int16 if L.obs_dtype == OBS_I16 else np.int32
        top = 32767 if dt == np.int16 else 2147483647
        out = np.zeros((CH, G, G), np.int64)
        _, rec = self.record(image)
        ent = self.entries(image)
        if ent is None:
            return None
        for f in range(L.frames):
            m = rec[3 + 2 * f * L.mask_words: 3 + 2 * (f + 1) * L.mask_words]
            bits = lambda w: np.unpackbits(w.view(np.uint8), bitorder="little")[:G].astype(bool)  # noqa: E731
            rows, cols = bits(m[:L.mask_words].copy()), bits(m[L.mask_words:].copy())
            out[f * L.channels][rows[:, None] | cols[None, :]] = -1
        flat = out.reshape(-1)
        for op, off, v in zip(*ent):
            x = flat[off]
            v = min(int(v), top)
            flat[off] = v if op == 0 else min(x + v, top) if op == 1 else (x if (x != 0 and x < v) else v) if op == 2 else max(x, v)
        return out.astype(dt)

    def expand(self, image):
        """dense [frames*C, G, G] frame of one image (C decoder, agarcl_batch_lists_expand)"""
        L = self.L
        out = np.empty((L.frames * L.channels, L.grid, L.grid), np.int16 if L.obs_dtype == OBS_I16 else np.int32)
        _lib.check(_lib.lib().agarcl_batch_lists_expand(self._b._h, int(image), out.ctypes.data_as(_vp)))
        return out

    def rewards_dones(self):
        """(rewards f64 [n_images], dones u8 [n_images]) read from the records"""
        L = self.L
        rew, done = np.zeros(L.n_images, np.float64), np.zeros(L.n_images, np.uint8)
        for i in range(L.n_images):
            _, rec = self.record(i)
            rew[i] = rec[L.rec_words - 3: L.rec_words - 1].copy().view(np.float64)[0]
            done[i] = rec[L.rec_words - 1]
        return rew, done


class Batch:
    def __init__(self, cfg):
        self.cfg = cfg
        self._h = _vp()
        _lib.check(_lib.lib().agarcl_batch_create(C.byref(cfg), C.byref(self._h)))
        self.layout = Layout()
        _lib.check(_lib.lib().agarcl_batch_get_layout(self._h, C.byref(self.layout)))
        self.N = cfg.n_instances
        self.A = self.layout.A
        shape = (C.c_int64 * 4)()
        ptr = _vp()
        dt = C.c_int32()
        _lib.check(_lib.lib().agarcl_batch_obs(self._h, C.byref(ptr), C.byref(shape), C.byref(dt)))
        self.obs_shape = tuple(int(x) for x in shape)
        self.obs_dtype = np.int16 if dt.value == OBS_I16 else np.int32
        self._obs_ptr = ptr.value
        p = _vp()
        _lib.check(_lib.lib().agarcl_batch_rewards(self._h, C.byref(p)))
        self._rew_ptr = p.value
        p = _vp()
        _lib.check(_lib.lib().agarcl_batch_dones(self._h, C.byref(p)))
        self._done_ptr = p.value

    def close(self):
        if self._h:
            self._mirror = None
            _lib.lib().agarcl_batch_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reference-shaped calls
    def seed(self, seeds):
        s = np.ascontiguousarray(np.broadcast_to(np.asarray(seeds, dtype=np.uint64), (self.N,)) if np.ndim(seeds) else
                                 np.arange(self.N, dtype=np.uint64) + np.uint64(seeds))
        _lib.check(_lib.lib().agarcl_batch_seed(self._h, s.ctypes.data_as(_vp)))

    def reset(self, mask=None, stream=0):
        m = None
        if mask is not None:
            m = np.ascontiguousarray(mask, dtype=np.uint8)
            assert m.size == self.N
        _lib.check(_lib.lib().agarcl_batch_reset(self._h, m.ctypes.data_as(_vp) if m is not None else None, _vp(stream)))
        if m is None:  # strict_reference: the player order of the new episode (quirk Q3)
            _lib.check(_lib.lib().agarcl_batch_get_layout(self._h, C.byref(self.layout)))

    def set_actions(self, dxdy, act, stream=0):
        dxdy = np.ascontiguousarray(dxdy, dtype=np.float32)
        act = np.ascontiguousarray(act, dtype=np.int32)
        assert dxdy.size == self.N * self.A * 2 and act.size == self.N * self.A, "Number of actions does not match number of agents"
        _lib.check(_lib.lib().agarcl_batch_set_actions(self._h, dxdy.ctypes.data_as(_vp), act.ctypes.data_as(_vp), 0, _vp(stream)))

    def set_actions_device(self, dxdy_ptr, act_ptr, stream=0):
        _lib.check(_lib.lib().agarcl_batch_set_actions(self._h, _vp(dxdy_ptr), _vp(act_ptr), 1, _vp(stream)))

    def step(self, stream=0):
        _lib.check(_lib.lib().agarcl_batch_step(self._h, _vp(stream)))

    def render(self, stream=0):
        _lib.check(_lib.lib().agarcl_batch_render(self._h, _vp(stream)))

    def render_ram(self, stream=0):
        _lib.check(_lib.lib().agarcl_batch_render_ram(self._h, _vp(stream)))

    def ram_tensor(self):
        """structured observation records [N, P, RAM_RECORD] float32 on the device (cfg.ram_obs)"""
        import torch
        p, shape = _vp(), (C.c_int64 * 3)()
        _lib.check(_lib.lib().agarcl_batch_ram(self._h, C.byref(p), C.byref(shape)))
        return torch.as_tensor(_CudaView(p.value, tuple(shape), "<f4", self), device=f"cuda:{self.cfg.device}")

    def _host_actions(self, dxdy, act):
        """the C ABI reads raw float32 / int32 buffers: coerce (and check the sizes of) whatever the caller hands in"""
        dxdy = np.ascontiguousarray(dxdy, dtype=np.float32)
        act = np.ascontiguousarray(act, dtype=np.int32)
        if dxdy.size != self.N * self.A * 2 or act.size != self.N * self.A:
            raise _lib.AgarclError("Number of actions does not match number of agents")
        return dxdy, act

    @staticmethod
    def _host_out(a, dtype, size, what):
        if a is None:
            return None
        if not (isinstance(a, np.ndarray) and a.dtype == dtype and a.flags.c_contiguous and a.size == size):
            raise _lib.AgarclError(f"{what} must be a C-contiguous {np.dtype(dtype).name} array of {size} elements")
        return a.ctypes.data_as(_vp)

    def step_host(self, dxdy, act, obs_out=None, rewards_out=None, dones_out=None):
        dxdy, act = self._host_actions(dxdy, act)
        NA = self.N * self.A
        _lib.check(_lib.lib().agarcl_batch_step_host(
            self._h, dxdy.ctypes.data_as(_vp), act.ctypes.data_as(_vp),
            self._host_out(obs_out, self.obs_dtype, int(np.prod(self.obs_shape)), "obs_out"),
            self._host_out(rewards_out, np.float64, NA, "rewards_out"), self._host_out(dones_out, np.uint8, NA, "dones_out")))

    # ---- host-resident observation mirror (include/agarcl_b200.h, mirror.cu)
    def mirror(self):
        """numpy view [N*A, C*frames, G, G] of the library-owned pinned host mirror of the observation (read-only)"""
        if getattr(self, "_mirror", None) is None:
            p, shape, dt = _vp(), (C.c_int64 * 4)(), C.c_int32()
            _lib.check(_lib.lib().agarcl_batch_mirror(self._h, C.byref(p), C.byref(shape), C.byref(dt)))
            n = int(np.prod(list(shape)))
            ct = C.c_int16 if dt.value == OBS_I16 else C.c_int32
            a = np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(n,)).reshape(tuple(int(x) for x in shape))
            a.flags.writeable = False
            self._mirror = a
        return self._mirror

    def sync_mirror(self, stream=0):
        self.mirror()
        _lib.check(_lib.lib().agarcl_batch_sync_mirror(self._h, _vp(stream)))
        return self._mirror

    def step_mirror(self, dxdy, act, rewards_out=None, dones_out=None):
        """take_actions (host arrays) + step + mirror sync; returns the mirror view"""
        dxdy, act = self._host_actions(dxdy, act)
        NA = self.N * self.A
        self.mirror()
        _lib.check(_lib.lib().agarcl_batch_step_mirror(
            self._h, dxdy.ctypes.data_as(_vp), act.ctypes.data_as(_vp),
            self._host_out(rewards_out, np.float64, NA, "rewards_out"), self._host_out(dones_out, np.uint8, NA, "dones_out")))
        return self._mirror

    # ---- the observation as lists (include/agarcl_b200.h, agarcl_obs_lists): no dense host tensor, nothing to patch
    def step_lists(self, dxdy, act, rewards_out=None, dones_out=None):
        """take_actions (host arrays) + step; rewards / dones to host; returns an ObsLists view of the library-owned pinned
        lists (valid until the next step call)"""
        dxdy, act = self._host_actions(dxdy, act)
        NA = self.N * self.A
        L = _ObsListsStruct()
        _lib.check(_lib.lib().agarcl_batch_step_lists(
            self._h, dxdy.ctypes.data_as(_vp), act.ctypes.data_as(_vp),
            self._host_out(rewards_out, np.float64, NA, "rewards_out"), self._host_out(dones_out, np.uint8, NA, "dones_out"), C.byref(L)))
        return ObsLists(self, L)

    def mirror_stats(self):
        out = (C.c_uint64 * 4)()
        _lib.check(_lib.lib().agarcl_batch_mirror_stats(self._h, C.byref(out)))
        t = (C.c_uint64 * 4)()
        _lib.check(_lib.lib().agarcl_batch_mirror_timing(self._h, C.byref(t)))
        return dict(entries=int(out[0]), dense_images=int(out[1]), d2h_bytes=int(out[2]), host_threads=int(out[3]),
                    wait_us=int(t[0]), total_us=int(t[1]), launch_us=int(t[2]), call_us=int(t[3]))

    def set_timing(self, enable=True):
        _lib.check(_lib.lib().agarcl_batch_set_timing(self._h, int(enable)))

    def get_timing(self):
        """(engine-tick kernel ms, observation kernel ms, steps) accumulated since the last call"""
        a, o, n = C.c_double(), C.c_double(), C.c_int32()
        _lib.check(_lib.lib().agarcl_batch_get_timing(self._h, C.byref(a), C.byref(o), C.byref(n)))
        return a.value, o.value, n.value

    def launches_per_step(self):
        return _lib.lib().agarcl_batch_launches_per_step(self._h)

    def flags(self, stream=0):
        """(OR of hdr.flags over all instances, {flag name: instances with it set}); reduced on the device"""
        from ._abi import FLAG_NAMES
        o, counts = C.c_uint32(), (C.c_uint32 * 32)()
        _lib.check(_lib.lib().agarcl_batch_flags(self._h, _vp(stream), C.byref(o), C.byref(counts)))
        return int(o.value), {FLAG_NAMES.get(1 << i, f"bit{i}"): int(counts[i]) for i in range(32) if counts[i]}

    # ---- parity / snapshot transport
    def download_state(self, i):
        sv = StateView(self.layout)
        _lib.check(_lib.lib().agarcl_batch_download_state(self._h, i, sv.ptr))
        return sv

    def upload_state(self, i, sv):
        _lib.check(_lib.lib().agarcl_batch_upload_state(self._h, i, sv.ptr))

    def save_env_state(self, i, path):
        _lib.check(_lib.lib().agarcl_batch_save_env_state(self._h, i, str(path).encode()))

    def load_env_state(self, i, path, lossless=False):
        _lib.check(_lib.lib().agarcl_batch_load_env_state(self._h, i, str(path).encode(), int(lossless)))

    def set_replay(self, i, draws):
        d = np.ascontiguousarray(draws, dtype=np.float32)
        _lib.check(_lib.lib().agarcl_batch_set_replay(self._h, i, d.ctypes.data_as(_vp), d.size))

    # ---- zero-copy device views
    def obs_view(self):
        return _CudaView(self._obs_ptr, self.obs_shape, "<i2" if self.obs_dtype == np.int16 else "<i4", self)

    def rewards_view(self):
        return _CudaView(self._rew_ptr, (self.N * self.A,), "<f8", self)

    def dones_view(self):
        return _CudaView(self._done_ptr, (self.N * self.A,), "|u1", self)

    def obs_tensor(self):
        import torch
        return torch.as_tensor(self.obs_view(), device=f"cuda:{self.cfg.device}")

    def rewards_tensor(self):
        import torch
        return torch.as_tensor(self.rewards_view(), device=f"cuda:{self.cfg.device}")

    def dones_tensor(self):
        import torch
        return torch.as_tensor(self.dones_view(), device=f"cuda:{self.cfg.device}")
