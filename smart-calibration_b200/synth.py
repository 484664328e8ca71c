"""Synthetic transitions of the elastic-net shape: state, new_state ~ N(0,1) (C,S) f32,
action, hint ~ U(-1,1) (C,A), reward ~ N(0,1) (SAC) or U(0,1) (TD3/PER), terminal False (optionally a fraction True).
Shared by tests/, bench.py and the golden-vector generator so that every side sees identical inputs."""
import numpy as np


def make_transitions(seed, n, S, A, reward="normal", done_frac=0.0):
    rs = np.random.RandomState(seed)
    state = rs.standard_normal((n, S)).astype(np.float32)
    new_state = rs.standard_normal((n, S)).astype(np.float32)
    action = rs.uniform(-1, 1, (n, A)).astype(np.float32)
    hint = rs.uniform(-1, 1, (n, A)).astype(np.float32)
    if reward == "normal":
        rew = rs.standard_normal(n).astype(np.float32)
    else:
        rew = rs.uniform(0, 1, n).astype(np.float32)
    done = rs.uniform(0, 1, n) < done_frac
    return dict(state=state, new_state=new_state, action=action, hint=hint, reward=rew, done=done)


def pack(tr):
    """(n, 2S+2A+2) float32 rows [s | s2 | a | hint | r | done] -- the C ABI's packed transition"""
    return np.concatenate([tr["state"], tr["new_state"], tr["action"], tr["hint"], tr["reward"][:, None],
                           tr["done"].astype(np.float32)[:, None]], axis=1).astype(np.float32)


def make_calib_transitions(seed, n, M, A, img=128, done_frac=0.0):
    """synthetic calibration transitions (calib_sac.ReplayBuffer layout :26-42): influence-map image (1, img, img), sky
    table (M+1, 7), action / hint in [-1, 1]^A"""
    rs = np.random.RandomState(seed)
    return dict(img=rs.standard_normal((n, 1, img, img)).astype(np.float32), sky=rs.standard_normal((n, M + 1, 7)).astype(np.float32),
                img2=rs.standard_normal((n, 1, img, img)).astype(np.float32), sky2=rs.standard_normal((n, M + 1, 7)).astype(np.float32),
                action=rs.uniform(-1, 1, (n, A)).astype(np.float32), hint=rs.uniform(-1, 1, (n, A)).astype(np.float32),
                reward=rs.standard_normal(n).astype(np.float32), done=rs.uniform(0, 1, n) < done_frac)


def make_demix_transitions(seed, n, M, A, img=128, done_frac=0.0):
    """synthetic demixing transitions (demix_sac.ReplayBuffer layout): influence map (1, img, img) (img = 0: none, the MLP
    family), flat metadata (M,), action / hint in [-1, 1]^A"""
    rs = np.random.RandomState(seed)
    shp = (n, 1, img, img) if img else (n, 1, 1, 1)
    return dict(img=rs.standard_normal(shp).astype(np.float32), meta=rs.standard_normal((n, M)).astype(np.float32),
                img2=rs.standard_normal(shp).astype(np.float32), meta2=rs.standard_normal((n, M)).astype(np.float32),
                action=rs.uniform(-1, 1, (n, A)).astype(np.float32), hint=rs.uniform(-1, 1, (n, A)).astype(np.float32),
                reward=rs.standard_normal(n).astype(np.float32), done=rs.uniform(0, 1, n) < done_frac)
