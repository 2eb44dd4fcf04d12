"""test_gpu_az.py once more on precision "f16f8c" (TWR_PREC_F16_F8C, the mode bench.py's headline is measured in): main
products in fp16, the two correction products as fp8 MMAs -- near-tie rule at the 1e-3-grade margin."""
from suite_loader import clone_suite

globals().update(clone_suite("test_gpu_az", "f16f8c"))
