"""Minimal ``gymnasium`` stand-in (TEST INFRASTRUCTURE, used only by oracle/make_golden.py).

gymnasium is a third-party dependency of the reference (unpinned, pyproject.toml:21) that
is absent from this image and cannot be installed (no network).  This stub provides just
the surface the reference imports so that the reference's *own* env / wrapper /
runtime modules run unmodified while golden vectors are generated.  ``SyncVectorEnv`` here
is our restatement of SAME_STEP autoreset -- parity is unpinned at that boundary.
"""
from gymnasium import spaces, utils, vector  # noqa: F401
from gymnasium.core import ActionWrapper, Env, ObservationWrapper, Wrapper  # noqa: F401
