// rsx_scatter_2.cu -- scatter-pass instantiations for 2-byte records (see rsx_scatter.cuh).
#include "rsx_scatter_dispatch.cuh"

namespace rsx {
cudaError_t launch_scatter_2(const ScatterParams &sp, int payload_bytes, bool is_float, bool wide,
                          int num_sms, cudaStream_t st) {
	return launch_scatter_es<2>(sp, payload_bytes, is_float, wide, num_sms, st);
}
} // namespace rsx
