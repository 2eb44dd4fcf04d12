"""test_cpp_host.py once more on precision "f16f8c" (TWR_PREC_F16_F8C): the C++17 host drives the same C ABI in the mode
bench.py's headline is measured in."""
from suite_loader import clone_suite

globals().update(clone_suite("test_cpp_host", "f16f8c"))
