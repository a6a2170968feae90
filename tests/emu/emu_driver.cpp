// Host-emulation build of libsrcv_b200's C ABI — TEST INFRASTRUCTURE (see emu_cuda.h).
//
// tests/emu/__init__.py compiles the product's own sources (srcv_api.cu, srcv_prep.cu,
// srcv_dot.cu, srcv_dot_bwd.cu, srcv_mlp_generic.cu, srcv_mlp_bwd.cu) as host C++ with
// -DSRCV_HOST_EMU and links them with this file into tests/emu/_build/libsrcv_emu.so, which
// therefore exports the SAME entry points as include/srcv_b200.h, taking HOST pointers.
// The tcgen05 kernel (srcv_mlp_tc.cu) cannot be emulated; the stubs below make the library
// report it as unsupported, so srcv_mlp_forward_f32 takes the fp32 SIMT variant here.
#define SRCV_HOST_EMU 1
#include "../../simplerecon_b200/csrc/srcv_kernels.h"

namespace srcv {
bool mlp_tc_supported(const srcv_shape&, const srcv_mlp_weights&) { return false; }
size_t mlp_tc_extra_bytes() { return 0; }
cudaError_t launch_mlp_tc(const srcv_shape&, const float*, const Workspace&, const float*, bool,
                          const srcv_mlp_weights&, float*, float*, uint8_t*, cudaStream_t) {
  return cudaErrorInvalidValue;
}
cudaError_t launch_tc_selftest(const float*, const float*, int, float*, void*, cudaStream_t) {
  return cudaErrorInvalidValue;
}
}  // namespace srcv

extern "C" {
// "SM count" seen by the launch heuristics (plane-loop split of the dot sweep, persistent
// grid of the MLP backward): lets a test drive those code paths with tiny inputs.
void emu_set_sms(int n) { ::emu::g_sms = n > 0 ? n : 1; }
}
