"""Minimal stand-in for the `gymnasium` package (absent from this image) -- just enough for the reference's environments to be
IMPORTED and stepped by tools/record_main_sac_transcript.py: `gym.Env` base class and `spaces.Box` / `spaces.Dict` containers.
Test infrastructure only."""
from . import spaces  # noqa: F401


class Env(object):
    metadata = {}

    def __init__(self, *a, **k):
        pass
