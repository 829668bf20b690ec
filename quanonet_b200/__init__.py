"""quanonet_b200 — B200-native batched statevector simulator for QuanONet's HEA hot path."""
__version__ = "0.1.0"
