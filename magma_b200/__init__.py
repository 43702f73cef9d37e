"""magma_b200 — B200-native (sm_100a) re-backing of the MAGMA forward/backward hot path."""
__version__ = "0.1.0"
