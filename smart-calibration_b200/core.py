"""Host-side handle around one libsmartcal_b200 agent: HBM arena owned by a torch tensor, zero-copy torch views
of every tensor the C ABI exposes, and thin call wrappers.  PyTorch is plumbing here (device memory, streams,
torch.distributed); all arithmetic of the update path runs in the CUDA library."""
import ctypes as C

import numpy as np
import torch

from . import _lib as L

_TORCH_DT = {4: torch.float32, 8: torch.float64, 1: torch.uint8}


class DeviceAgent:
    def __init__(self, kind, state_dim, n_actions, batch_size, mem_size, gamma, tau, lr_actor, lr_critic,
                 reward_scale=1.0, alpha=0.0, prioritized=False, use_hint=False, update_actor_interval=2,
                 per_clip=100.0, hint_threshold=0.1, admm_rho=0.01, max_action=1.0, rng_seed=0, device=None, flags=0):
        self.lib = L.load()
        if not torch.cuda.is_available():
            raise L.SmartcalError("no CUDA device visible: smart-calibration_b200 has no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.kind, self.S, self.A, self.B, self.C = kind, int(state_dim), int(n_actions), int(batch_size), int(mem_size)
        self.T = 2 * self.S + 2 * self.A + 2
        self.prioritized = bool(prioritized)
        cfg = L.ScConfig(abi_version=L.SC_ABI_VERSION, kind=kind, state_dim=self.S, n_actions=self.A,
                         batch_size=self.B, mem_size=self.C, prioritized=int(bool(prioritized)),
                         use_hint=int(bool(use_hint)), update_actor_interval=int(update_actor_interval), flags=int(flags),
                         gamma=float(gamma), tau=float(tau), lr_actor=float(lr_actor), lr_critic=float(lr_critic),
                         reward_scale=float(reward_scale), alpha=float(alpha), max_action=float(max_action),
                         per_clip=float(per_clip), hint_threshold=float(hint_threshold), admm_rho=float(admm_rho),
                         rng_seed=int(rng_seed))
        self.cfg = cfg
        nbytes = self.lib.sc_agent_arena_bytes(C.byref(cfg))
        if nbytes == 0:
            raise L.SmartcalError(self.lib.sc_last_error().decode())
        with torch.cuda.device(self.device):
            self.arena = torch.zeros(nbytes + 256, dtype=torch.uint8, device=self.device)
            self._pad = (-self.arena.data_ptr()) % 256
            self.stream = torch.cuda.current_stream(self.device)
            h = C.c_void_p()
            L.check(self.lib.sc_agent_create(C.byref(cfg), C.c_void_p(self.arena.data_ptr() + self._pad), nbytes,
                                             C.c_void_p(self.stream.cuda_stream), C.byref(h)))
        self.h = h
        self._noise = L.ScNoise()
        self._stats = L.ScLearnStats()
        self._pending = None          # a packed transition whose store is fused into the next learn() (sc_agent_store_learn)
        self.n_stored = 0             # host mirror of mem_cntr (the facade's "buffer filled yet?" test costs no ABI call)
        self.n_learned = 0            # host mirror of learn_counter

    def flush(self):
        """issue a deferred store_transition now (anything that looks at the replay state calls this first)"""
        if self._pending is not None:
            packed, mode, prio = self._pending
            self._pending = None
            L.check(self.lib.sc_replay_store_host(self.h, L.ptr(packed), int(mode), float(prio)))

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.sc_agent_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- views ------------------------------------------------------------------------------------------
    def _view(self, info):
        dt = torch.int32 if info.is_int else _TORCH_DT[info.elem_bytes]
        off = info.dev_ptr - self.arena.data_ptr()
        nbytes = ((info.rows - 1) * info.pitch + info.cols) * info.elem_bytes
        flat = self.arena[off:off + nbytes].view(dt)
        return flat.as_strided((info.rows, info.cols), (info.pitch, 1))

    def tensor(self, net, name, slot=L.SLOT_PARAM):
        info = L.ScTensorInfo()
        L.check(self.lib.sc_agent_tensor(self.h, net, slot, name.encode(), C.byref(info)))
        return self._view(info)

    def buffer(self, name):
        self.flush()
        info = L.ScTensorInfo()
        L.check(self.lib.sc_agent_buffer(self.h, name.encode(), C.byref(info)))
        return self._view(info)

    # ---- counters ---------------------------------------------------------------------------------------
    def counters(self):
        self.flush()
        m, l, d, b = C.c_int64(), C.c_int64(), C.c_int64(), C.c_double()
        L.check(self.lib.sc_agent_get_counters(self.h, C.byref(m), C.byref(l), C.byref(d), C.byref(b)))
        return m.value, l.value, d.value, b.value

    def set_counters(self, mem_cntr, learn_counter, data_pointer, per_beta):
        self.flush()
        self.n_stored, self.n_learned = int(mem_cntr), int(learn_counter)
        L.check(self.lib.sc_agent_set_counters(self.h, int(mem_cntr), int(learn_counter), int(data_pointer), float(per_beta)))

    def set_scalars(self, alpha, rho, admm_rho):
        L.check(self.lib.sc_agent_set_scalars(self.h, float(alpha), float(rho), float(admm_rho)))

    def rho(self):
        """current hint multiplier / adaptive ADMM penalty (device scalar)"""
        self.flush()
        return float(self.buffer("rho")[0, 0].item())

    def sync(self):
        self.flush()
        L.check(self.lib.sc_agent_sync(self.h))

    # ---- replay -----------------------------------------------------------------------------------------
    def pack(self, s, a, r, s2, done, hint):
        t = np.empty(self.T, np.float32)
        S, A = self.S, self.A
        t[:S] = s
        t[S:2 * S] = s2
        t[2 * S:2 * S + A] = a
        t[2 * S + A:2 * S + 2 * A] = hint
        t[2 * S + 2 * A] = r
        t[2 * S + 2 * A + 1] = 1.0 if done else 0.0
        return t

    def store(self, packed, priority_mode=0, priority=0.0, defer=False):
        """store one packed transition.  defer=True (the facades' store_transition): the store is held back and fused
        into the next learn() -- one pinned block, one host->device copy, one graph launch for store + update"""
        self.flush()
        packed = np.ascontiguousarray(packed, np.float32)
        self.n_stored += 1
        if defer:
            self._pending = (packed, int(priority_mode), float(priority))
            return
        L.check(self.lib.sc_replay_store_host(self.h, L.ptr(packed), int(priority_mode), float(priority)))

    def store_dev(self, block, index, priority_mode=0, priority=0.0):
        self.flush()
        self.n_stored += 1
        L.check(self.lib.sc_replay_store_dev(self.h, C.c_void_p(block.data_ptr()), int(index), int(priority_mode), float(priority)))

    def fill(self, packed, priorities=None):
        packed = np.ascontiguousarray(packed, np.float32)
        pr = None if priorities is None else np.ascontiguousarray(priorities, np.float64)
        self.flush()
        self.n_stored = int(packed.shape[0])
        L.check(self.lib.sc_replay_fill_host(self.h, L.ptr(packed), packed.shape[0], L.ptr(pr)))

    def set_batch(self, idx):
        self.flush()
        idx = np.ascontiguousarray(idx, np.int32)
        L.check(self.lib.sc_replay_set_batch_host(self.h, L.ptr(idx), idx.shape[0]))

    def per_sample(self, u01=None, want_outputs=False):
        self.flush()
        u = None if u01 is None else np.ascontiguousarray(u01, np.float64)
        if want_outputs:
            ti, w = np.empty(self.B, np.int32), np.empty(self.B, np.float32)
            L.check(self.lib.sc_per_sample(self.h, L.ptr(u), L.ptr(ti), L.ptr(w)))
            return ti, w
        L.check(self.lib.sc_per_sample(self.h, L.ptr(u), None, None))
        return None

    def per_update(self, tree_idx, abs_err):
        self.flush()
        ti = np.ascontiguousarray(tree_idx, np.int32)
        er = np.ascontiguousarray(np.asarray(abs_err, np.float32).reshape(-1))
        ps = np.empty(self.B, np.float32)
        L.check(self.lib.sc_per_update_host(self.h, L.ptr(ti), L.ptr(er), ti.shape[0], L.ptr(ps)))
        return ps

    # ---- compute ----------------------------------------------------------------------------------------
    def act(self, obs, eps=None):
        obs = np.ascontiguousarray(obs, np.float32)
        e = None if eps is None else np.ascontiguousarray(eps, np.float32)
        out = np.empty(self.A, np.float32)
        L.check(self.lib.sc_agent_act(self.h, L.ptr(obs), L.ptr(e), L.ptr(out)))
        return out

    def _fill_noise(self, n, keep, eps_next, eps_cur, eps_dual, td3_noise, per_u01, batch_idx):
        def hp(x, dt):
            if x is None:
                return None
            arr = np.ascontiguousarray(x, dt)
            keep.append(arr)
            return arr.ctypes.data

        n.eps_next_host, n.eps_cur_host, n.eps_dual_host = hp(eps_next, np.float32), hp(eps_cur, np.float32), hp(eps_dual, np.float32)
        n.has_td3_target_noise = 0 if td3_noise is None else 1
        n.td3_target_noise = 0.0 if td3_noise is None else float(td3_noise)
        n.per_u01_host, n.batch_idx_host = hp(per_u01, np.float64), hp(batch_idx, np.int32)

    def learn(self, eps_next=None, eps_cur=None, eps_dual=None, td3_noise=None, want_stats=False, per_u01=None, batch_idx=None):
        """one Agent.learn().  Host arrays (noise, PER uniforms / batch indices of the caller's numpy stream) and a deferred
        store_transition all ride in the step's single host->device copy."""
        keep = []
        n = self._noise
        self._fill_noise(n, keep, eps_next, eps_cur, eps_dual, td3_noise, per_u01, batch_idx)
        st = C.byref(self._stats) if want_stats else None
        if self._pending is not None:
            packed, mode, prio = self._pending
            self._pending = None
            L.check(self.lib.sc_agent_store_learn(self.h, L.ptr(packed), mode, prio, C.byref(n), st))
        else:
            L.check(self.lib.sc_agent_learn(self.h, C.byref(n), st))
        if self.n_stored >= self.B:
            self.n_learned += 1
        if want_stats:
            s = self._stats
            return {f: getattr(s, f) for f, _ in s._fields_}
        return None

    def set_stats_mode(self, deferred):
        """deferred=True: learn(want_stats=True) returns the PREVIOUS step's losses without synchronising (sc_agent_set_stats_mode)"""
        L.check(self.lib.sc_agent_set_stats_mode(self.h, int(bool(deferred))))

    def soft_update(self, tau):
        L.check(self.lib.sc_agent_soft_update(self.h, float(tau)))

    def params_changed(self):
        """call after writing parameters through a view obtained EARLIER (views obtained by tensor() right before the write
        need nothing): the tf32 operand copies of the wide weight matrices are re-derived by the next learn()"""
        L.check(self.lib.sc_agent_params_changed(self.h))

    def launch_count(self):
        return self.lib.sc_agent_launch_count(self.h)

    def enable_timing(self, on=True):
        L.check(self.lib.sc_agent_enable_kernel_timing(self.h, int(on)))

    def kernel_time(self, which, reset=False):
        ms, n = C.c_double(), C.c_int64()
        L.check(self.lib.sc_agent_kernel_time(self.h, which, int(reset), C.byref(ms), C.byref(n)))
        return ms.value, n.value


class DeviceEnsemble:
    """E agents of identical configuration stepped in lock-step by ONE launch sequence (sc_ensemble_*): every kernel of
    learn() runs once with blockIdx.y = agent.  The reference's counterpart is one python process per seed
    (elasticnet/do.sh:1-6).  Each agent's results are bit-identical to the same agent stepped alone."""

    def __init__(self, devs):
        self.devs = list(devs)
        self.lib = self.devs[0].lib
        E = len(self.devs)
        arr = (C.c_void_p * E)(*[d.h.value for d in self.devs])
        h = C.c_void_p()
        L.check(self.lib.sc_ensemble_create(arr, E, C.byref(h)))
        self.h = h
        self._noise = (L.ScNoise * E)()
        self._stats = (L.ScLearnStats * E)()

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.sc_ensemble_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def learn(self, noise=None, want_stats=False):
        """noise: None (device Philox for every agent) or a list of E dicts with the keyword arguments of DeviceAgent.learn"""
        keep = []
        for d in self.devs:
            d.flush()
        na = None
        if noise is not None:
            for i, kw in enumerate(noise):
                kw = kw or {}
                self.devs[i]._fill_noise(self._noise[i], keep, kw.get("eps_next"), kw.get("eps_cur"), kw.get("eps_dual"),
                                         kw.get("td3_noise"), kw.get("per_u01"), kw.get("batch_idx"))
            na = self._noise
        L.check(self.lib.sc_ensemble_learn(self.h, na, self._stats if want_stats else None))
        for d in self.devs:
            if d.n_stored >= d.B:
                d.n_learned += 1
        if want_stats:
            return [{f: getattr(s, f) for f, _ in s._fields_} for s in self._stats]
        return None

    def launch_count(self):
        return self.lib.sc_ensemble_launch_count(self.h)
