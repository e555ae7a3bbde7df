"""TEST INFRASTRUCTURE ONLY (oracle/): re-statement of the part of PYPOWER 5.1.4 the reference calls
(requirements.txt:9; call sites pypownet/grid.py:62-65, 226-231, 595).  PYPOWER is not installed in this
image and cannot be fetched (no network); see api.py for what is restated and how it is pinned."""
