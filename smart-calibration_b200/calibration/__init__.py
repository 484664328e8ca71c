"""Drop-in modules for the reference's calibration/ directory (conv-SAC agent on libsmartcal_b200.so)."""
