"""cabana_b200 -- B200-native particle neighbour lists behind Cabana's API.

Host-side Python mirror of the reference interface for the hot path
(LinkedCellList -> VerletList -> neighbor_parallel_for, plus slab Halo/Distributor).
All compute goes through the C-ABI shared library built from ``csrc/`` (hand-written
sm_100a CUDA); there is no CPU fallback -- importing the compute modules without the
built library raises.
"""
__version__ = "0.1.0"
