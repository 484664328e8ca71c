"""Host-side handle around one libsmartcal_b200 calibration (conv-SAC) agent: torch-owned HBM arena, zero-copy views of
every tensor the C ABI exposes, thin call wrappers.  PyTorch is plumbing (device memory, streams)."""
import ctypes as C

import numpy as np
import torch

from . import _lib as L

_DT = {(4, 0): torch.float32, (4, 1): torch.int32, (8, 1): torch.int64, (8, 0): torch.float64}


class CalibDevice:
    def __init__(self, batch_size, mem_size, n_actions, M, gamma, tau, lr_actor, lr_critic, alpha, use_hint=False, hint_threshold=0.1,
                 admm_rho=1.0, img=128, rng_seed=0, device=None, variant=0, update_actor_interval=2):
        self.lib = L.load()
        if not torch.cuda.is_available():
            raise L.SmartcalError("no CUDA device visible: smart-calibration_b200 has no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.B, self.C, self.A, self.M, self.img = int(batch_size), int(mem_size), int(n_actions), int(M), int(img)
        cfg = L.ScCalibConfig(abi_version=L.SC_ABI_VERSION, batch_size=self.B, mem_size=self.C, n_actions=self.A, M=self.M, img=self.img,
                              use_hint=int(bool(use_hint)), variant=int(variant), gamma=float(gamma), tau=float(tau), lr_actor=float(lr_actor),
                              lr_critic=float(lr_critic), alpha=float(alpha), hint_threshold=float(hint_threshold),
                              admm_rho=float(admm_rho), rng_seed=int(rng_seed), update_actor_interval=int(update_actor_interval))
        self.cfg = cfg
        nbytes = self.lib.sc_calib_arena_bytes(C.byref(cfg))
        if nbytes == 0:
            raise L.SmartcalError(self.lib.sc_last_error().decode())
        with torch.cuda.device(self.device):
            self.arena = torch.zeros(nbytes + 256, dtype=torch.uint8, device=self.device)
            self._pad = (-self.arena.data_ptr()) % 256
            self.stream = torch.cuda.current_stream(self.device)
            h = C.c_void_p()
            L.check(self.lib.sc_calib_create(C.byref(cfg), C.c_void_p(self.arena.data_ptr() + self._pad), nbytes,
                                             C.c_void_p(self.stream.cuda_stream), C.byref(h)))
        self.h = h
        self._noise = L.ScNoise()
        self._stats = L.ScLearnStats()

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.sc_calib_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _view(self, info):
        dt = _DT[(info.elem_bytes, info.is_int)]
        off = info.dev_ptr - self.arena.data_ptr()
        nbytes = ((info.rows - 1) * info.pitch + info.cols) * info.elem_bytes
        return self.arena[off:off + nbytes].view(dt).as_strided((info.rows, info.cols), (info.pitch, 1))

    def tensor(self, net, name, slot=L.SLOT_PARAM):
        info = L.ScTensorInfo()
        L.check(self.lib.sc_calib_tensor(self.h, net, slot, name.encode(), C.byref(info)))
        return self._view(info)

    def buffer(self, name):
        info = L.ScTensorInfo()
        L.check(self.lib.sc_calib_buffer(self.h, name.encode(), C.byref(info)))
        return self._view(info)

    def counters(self):
        m, l = C.c_int64(), C.c_int64()
        L.check(self.lib.sc_calib_get_counters(self.h, C.byref(m), C.byref(l)))
        return m.value, l.value

    def set_counters(self, mem_cntr, learn_counter):
        L.check(self.lib.sc_calib_set_counters(self.h, int(mem_cntr), int(learn_counter)))

    def set_scalars(self, rho, admm_rho):
        L.check(self.lib.sc_calib_set_scalars(self.h, float(rho), float(admm_rho)))

    @staticmethod
    def _f32(x):
        return np.ascontiguousarray(np.asarray(x, dtype=np.float32).reshape(-1))

    def store(self, img, sky, action, reward, img2, sky2, done, hint):
        arrs = [None if v is None else self._f32(v) for v in (img, sky, img2, sky2, action, hint)]     # (img / img2 None: MLP family)
        L.check(self.lib.sc_calib_store_host(self.h, *[L.ptr(v) for v in arrs], float(reward), int(bool(done))))

    def set_batch(self, idx):
        idx = np.ascontiguousarray(np.asarray(idx, dtype=np.int32))
        L.check(self.lib.sc_calib_set_batch_host(self.h, L.ptr(idx), idx.size))

    def act(self, img, sky, eps=None):
        img, sky = (None if img is None else self._f32(img)), self._f32(sky)
        e = None if eps is None else self._f32(eps)
        out = np.empty(self.A, np.float32)
        L.check(self.lib.sc_calib_act(self.h, L.ptr(img), L.ptr(sky), L.ptr(e), L.ptr(out)))
        return out

    def learn(self, eps_next=None, eps_cur=None, want_stats=False, eps_dual=None):
        n = self._noise
        keep = []
        for f, v in (("eps_next_host", eps_next), ("eps_cur_host", eps_cur), ("eps_dual_host", eps_dual)):
            if v is None:
                setattr(n, f, None)
            else:
                arr = self._f32(v)
                keep.append(arr)
                setattr(n, f, arr.ctypes.data)
        if want_stats:
            L.check(self.lib.sc_calib_learn(self.h, C.byref(n), C.byref(self._stats)))
            s = self._stats
            return {f: getattr(s, f) for f, _ in s._fields_}
        L.check(self.lib.sc_calib_learn(self.h, C.byref(n), None))
        return None

    def learn_td3(self, per_u01=None, target_noise=None, want_stats=False):
        """SC_CALIB_TD3 agents: one demix_td3 learn().  per_u01 (B float64 uniforms of PER.sample_buffer) and target_noise (the
        np.random.normal(scale=0.2) scalar) reproduce the caller's numpy stream; None: device Philox"""
        n = self._noise
        n.eps_next_host = n.eps_cur_host = n.eps_dual_host = n.batch_idx_host = None
        keep = None
        if per_u01 is None:
            n.per_u01_host, n.has_td3_target_noise = None, 0
        else:
            keep = np.ascontiguousarray(np.asarray(per_u01, dtype=np.float64).reshape(-1))
            if keep.size != self.B:
                raise ValueError(f"need {self.B} uniforms")
            n.per_u01_host, n.td3_target_noise, n.has_td3_target_noise = keep.ctypes.data, float(target_noise), 1
        try:
            if want_stats:
                L.check(self.lib.sc_calib_learn(self.h, C.byref(n), C.byref(self._stats)))
                s = self._stats
                return {f: getattr(s, f) for f, _ in s._fields_}
            L.check(self.lib.sc_calib_learn(self.h, C.byref(n), None))
            return None
        finally:
            n.per_u01_host, n.has_td3_target_noise = None, 0

    def per_beta(self):
        b = C.c_double()
        L.check(self.lib.sc_calib_get_per_beta(self.h, C.byref(b)))
        return b.value

    def set_per_beta(self, beta):
        L.check(self.lib.sc_calib_set_per_beta(self.h, float(beta)))

    def soft_update(self, tau):
        L.check(self.lib.sc_calib_soft_update(self.h, float(tau)))

    def sync(self):
        L.check(self.lib.sc_calib_sync(self.h))

    def launch_count(self):
        return self.lib.sc_calib_launch_count(self.h)
