"""pypownet_b200: the per-timestep hot path of pypownet (topology -> load-flow -> overflows -> cascading failures ->
game over -> observation/reward) as CUDA kernels for B200 behind a C ABI, with a host-side mirror of the reference's
RunEnv / Runner / Agent interface.  See DESIGN.md and INTEGRATION.md."""
ARTIFICIAL_NODE_STARTING_STRING = '666'        # pypownet/__init__.py:10
__all__ = ['ARTIFICIAL_NODE_STARTING_STRING']
