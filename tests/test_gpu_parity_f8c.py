"""test_gpu_parity.py once more on precision "f16f8c" (TWR_PREC_F16_F8C: the terms of f16x2w16 with the two correction
products issued as fp8 MMAs) -- same 1e-3 bar, same near-tie rule as the w16 suite."""
from suite_loader import clone_suite

globals().update(clone_suite("test_gpu_parity", "f16f8c"))
