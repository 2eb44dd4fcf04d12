// twr_forward_tc.cu -- K2, tensor-core variant (tcgen05 / TMEM / TMA).  Placeholder until the
// kernel lands: reports "unsupported" so engine creation with TWR_PREC_F16X2 fails loudly.
#include "twr_kernels.cuh"

int forward_tc_supported(const PolicyDev&, const EnvParams&, const char** why) {
    static const char* msg = "TWR_PREC_F16X2 (tcgen05) forward is not built into this library yet";
    *why = msg;
    return 0;
}
size_t forward_tc_pack_bytes(const PolicyDev&) { return 0; }
void launch_forward_tc_pack(cudaStream_t, const PolicyDev&, void*) {}
void launch_forward_tc(cudaStream_t, const PolicyDev&, const ForwardArgs&) {}
