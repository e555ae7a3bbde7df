"""TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for `gym` 0.12 so the unmodified
reference package can be imported in this container (reference use: pypownet/environment.py:9
imports MultiBinary, Box, Dict, Discrete and nothing else).  Never imported by pypownet_b200."""
from . import spaces  # noqa: F401
