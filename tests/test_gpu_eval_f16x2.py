"""test_gpu_eval.py once more, in this process, on the tensor-core forward with every operand split (precision "f16x2",
TWR_PREC_F16X2): the fused persistent pair kernel is the path bench.py measures."""
from suite_loader import clone_suite

globals().update(clone_suite("test_gpu_eval", "f16x2"))
