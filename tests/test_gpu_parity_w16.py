"""test_gpu_parity.py once more on the tensor-core forward's 1e-3-grade mode (precision "f16x2w16", TWR_PREC_F16X2_W16:
the common layer's weight as one fp16 term) -- the mode bench.py's headline `value` is measured in."""
from suite_loader import clone_suite

globals().update(clone_suite("test_gpu_parity", "f16x2w16"))
